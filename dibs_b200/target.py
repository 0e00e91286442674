"""Synthetic ground-truth Bayes nets for benchmarks and examples (host side, NumPy) -- the reference's data factory
(dibs/target.py:12-321) with the same function names, keyword-only signatures, defaults and return values
``(data, graph_model, likelihood_model)``.

Random numbers come from ``numpy.random.default_rng`` seeded with the two words of the JAX-style ``key`` (bit parity
of the DATA with the reference is not needed: the same ``x`` is fed to both sides of every comparison);
graphs are sampled by ``dibs_b200.synthetic`` (ER: graph.py:44-53; scale-free: graph.py:132-151).
"""
from typing import Any, NamedTuple

import numpy as np

from . import synthetic as syn
from .models import (BGe, DenseNonlinearGaussian, ErdosReniDAGDistribution, LinearGaussian, ScaleFreeDAGDistribution)


class Data(NamedTuple):
    """Ground-truth network and observations (dibs/target.py:12-40)."""
    passed_key: Any
    n_vars: int
    n_observations: int
    n_ho_observations: int
    g: Any
    theta: Any
    x: Any
    x_ho: Any
    x_interv: Any


def _rng(key):
    k = np.asarray(key, dtype=np.uint32).reshape(-1)
    return np.random.default_rng([int(v) for v in k])


def make_graph_model(*, n_vars, graph_prior_str, edges_per_node=2):
    """``er`` -> ErdosReniDAGDistribution, ``sf`` -> ScaleFreeDAGDistribution (dibs/target.py:122-149)."""
    if graph_prior_str == 'er':
        return ErdosReniDAGDistribution(n_vars=n_vars, n_edges_per_node=edges_per_node)
    if graph_prior_str == 'sf':
        return ScaleFreeDAGDistribution(n_vars=n_vars, n_edges_per_node=edges_per_node)
    raise ValueError(f"Invalid value `{graph_prior_str}` for `graph_prior_str`; choices: `er`, `sf`")


def _sample_g(rng, graph_model):
    per_node = getattr(graph_model, "n_edges_per_node", None)
    if isinstance(graph_model, ScaleFreeDAGDistribution):
        return syn.sample_sf_dag(rng, graph_model.n_vars, per_node or 2)
    if isinstance(graph_model, ErdosReniDAGDistribution):
        return syn.sample_er_dag(rng, graph_model.n_vars, graph_model.n_edges / graph_model.n_vars)
    raise NotImplementedError(f"no ground-truth sampler for {type(graph_model).__name__}")


def _sample_parameters(rng, model, n_vars):
    if isinstance(model, LinearGaussian):
        return syn.sample_linear_gaussian_parameters(rng, n_vars, model.mean_edge, model.sig_edge, model.min_edge)
    if isinstance(model, DenseNonlinearGaussian):
        model.check_native()
        return syn.sample_dense_nn_parameters(rng, n_vars, model.hidden, model.sig_param)
    raise NotImplementedError(f"no parameter sampler for {type(model).__name__}")


def _sample_obs(rng, model, g, theta, n_samples, interv=None):
    if isinstance(model, LinearGaussian):
        return syn.sample_obs_linear_gaussian(rng, g, theta, n_samples, model.obs_noise, interv)
    return syn.sample_obs_dense_nn(rng, g, theta, n_samples, model.obs_noise, interv)


def make_synthetic_bayes_net(*, key, n_vars, graph_model, generative_model, n_observations=100, n_ho_observations=100,
                             n_intervention_sets=10, perc_intervened=0.1):
    """Ground-truth DAG, parameters, observations, held-out observations and ``n_intervention_sets`` data sets under
    random 0-clamp interventions on ``ceil(n_vars * perc_intervened)`` nodes (dibs/target.py:43-119)."""
    rng = _rng(key)
    g = _sample_g(rng, graph_model)
    theta = _sample_parameters(rng, generative_model, n_vars)
    x = _sample_obs(rng, generative_model, g, theta, n_observations)
    x_ho = _sample_obs(rng, generative_model, g, theta, n_ho_observations)
    x_interv = []
    n_interv = int(np.ceil(n_vars * perc_intervened))
    for _ in range(n_intervention_sets):
        targets = rng.choice(n_vars, size=n_interv, replace=False)
        interv = {int(k): 0.0 for k in targets}
        x_interv.append((interv, _sample_obs(rng, generative_model, g, theta, n_observations, interv)))
    return Data(passed_key=np.array(key, copy=True), n_vars=n_vars, n_observations=n_observations,
                n_ho_observations=n_ho_observations, g=g, theta=theta, x=x, x_ho=x_ho, x_interv=x_interv)


def make_linear_gaussian_equivalent_model(*, key, n_vars=20, graph_prior_str='sf', bge_mean_obs=None, bge_alpha_mu=None,
                                          bge_alpha_lambd=None, obs_noise=0.1, mean_edge=0.0, sig_edge=1.0, min_edge=0.5,
                                          n_observations=100, n_ho_observations=100):
    """Linear-Gaussian ground truth with the BGe marginal likelihood as inference model (dibs/target.py:152-212)."""
    graph_model = make_graph_model(n_vars=n_vars, graph_prior_str=graph_prior_str)
    generative_model = LinearGaussian(n_vars=n_vars, obs_noise=obs_noise, mean_edge=mean_edge, sig_edge=sig_edge, min_edge=min_edge)
    likelihood_model = BGe(n_vars=n_vars, mean_obs=bge_mean_obs, alpha_mu=bge_alpha_mu, alpha_lambd=bge_alpha_lambd)
    data = make_synthetic_bayes_net(key=key, n_vars=n_vars, graph_model=graph_model, generative_model=generative_model,
                                    n_observations=n_observations, n_ho_observations=n_ho_observations)
    return data, graph_model, likelihood_model


def make_linear_gaussian_model(*, key, n_vars=20, graph_prior_str='sf', obs_noise=0.1, mean_edge=0.0, sig_edge=1.0,
                               min_edge=0.5, n_observations=100, n_ho_observations=100):
    """Linear-Gaussian ground truth and inference model (dibs/target.py:215-267)."""
    graph_model = make_graph_model(n_vars=n_vars, graph_prior_str=graph_prior_str)
    kw = dict(n_vars=n_vars, obs_noise=obs_noise, mean_edge=mean_edge, sig_edge=sig_edge, min_edge=min_edge)
    data = make_synthetic_bayes_net(key=key, n_vars=n_vars, graph_model=graph_model, generative_model=LinearGaussian(**kw),
                                    n_observations=n_observations, n_ho_observations=n_ho_observations)
    return data, graph_model, LinearGaussian(**kw)


def make_nonlinear_gaussian_model(*, key, n_vars=20, graph_prior_str='sf', obs_noise=0.1, sig_param=1.0,
                                  hidden_layers=(5,), n_observations=100, n_ho_observations=100):
    """Dense-MLP ground truth and inference model (dibs/target.py:270-321)."""
    graph_model = make_graph_model(n_vars=n_vars, graph_prior_str=graph_prior_str)
    kw = dict(n_vars=n_vars, hidden_layers=hidden_layers, obs_noise=obs_noise, sig_param=sig_param)
    data = make_synthetic_bayes_net(key=key, n_vars=n_vars, graph_model=graph_model,
                                    generative_model=DenseNonlinearGaussian(**kw), n_observations=n_observations,
                                    n_ho_observations=n_ho_observations)
    return data, graph_model, DenseNonlinearGaussian(**kw)
