"""TEST INFRASTRUCTURE ONLY -- golden-fixture generator.

Runs the UNMODIFIED reference sources (/root/reference/dibs, larslorch/dibs @ 5350d1a) on top of
oracle/jaxshim (JAX itself is not installable in this image) and records inputs/outputs of the SVGD
hot path as small ``.npz`` fixtures under tests/golden/.  Only runs in the build container (needs
/root/reference); the fixtures travel, this script's inputs do not.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Every fixture stores: the inputs (x, interv_mask, z, theta_flat, key, sf_baseline, t, hyper-parameters)
and the outputs of the reference's own methods (named after them).
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import jax  # noqa: E402,F401  (the shim)
from jax import random  # noqa: E402

from dibs.inference import JointDiBS, MarginalDiBS  # noqa: E402  (the reference)
from dibs.models import (BGe, LinearGaussian, DenseNonlinearGaussian,  # noqa: E402
                         ErdosReniDAGDistribution, ScaleFreeDAGDistribution, UniformDAGDistributionRejection)
from dibs.graph_utils import acyclic_constr_nograd  # noqa: E402

from dibs_b200.synthetic import make_linear_gaussian_data, make_nonlinear_gaussian_data  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def npy(t):
    if isinstance(t, torch.Tensor):
        return t.detach().numpy()
    return np.asarray(t)


def key_np(k):
    return k.detach().numpy().astype(np.uint32)


def flat_theta(theta, kind):
    """Reference theta pytree -> dibs_b200 flat layout [M, Dtheta]."""
    if kind == "lingauss":
        return npy(theta).reshape(theta.shape[0], -1)
    (w1, b1), _, (w2, b2) = theta
    m = w1.shape[0]
    return np.concatenate([npy(w1).reshape(m, -1), npy(b1).reshape(m, -1),
                           npy(w2).reshape(m, -1), npy(b2).reshape(m, -1)], axis=1)


def flat_theta_grad(theta, kind):
    return flat_theta(theta, kind)


def make_prior(kind, d, epn):
    if kind == "er":
        return ErdosReniDAGDistribution(n_vars=d, n_edges_per_node=epn)
    if kind == "sf":
        return ScaleFreeDAGDistribution(n_vars=d, n_edges_per_node=epn)
    return UniformDAGDistributionRejection(n_vars=d)


def make_case(name, *, lik, prior="er", epn=1, d=5, n_obs=20, m=3, s=8, a=4, t=7, hidden=4, interv=False,
              estimator=None, baseline=0.0, optimizer="rmsprop", n_dim=None, seed=0, alpha_linear=None,
              tau=1.0, beta_linear=1.0, activation="relu"):
    rng = np.random.default_rng(seed)
    if lik == "densenn":
        data = make_nonlinear_gaussian_data(seed=seed, n_vars=d, n_observations=n_obs, n_edges_per_node=1, hidden=hidden)
    else:
        data = make_linear_gaussian_data(seed=seed, n_vars=d, n_observations=n_obs, n_edges_per_node=1)
    x = torch.tensor(data["x"])
    mask_np = np.zeros((n_obs, d), np.int32)
    if interv:
        rows = rng.permutation(n_obs)[: n_obs // 3]
        mask_np[rows, rng.integers(0, d, size=rows.size)] = 1
    mask = torch.tensor(mask_np)
    gm = make_prior(prior, d, epn)
    joint = lik != "bge"
    kw = dict(x=x, graph_model=gm, interv_mask=mask, n_grad_mc_samples=s, n_acyclicity_mc_samples=a,
              score_function_baseline=baseline, optimizer=optimizer, tau=tau, beta_linear=beta_linear)
    if estimator:
        kw["grad_estimator_z"] = estimator
    if alpha_linear is not None:
        kw["alpha_linear"] = alpha_linear
    if lik == "bge":
        lm = BGe(n_vars=d)
        model = MarginalDiBS(likelihood_model=lm, **kw)
    elif lik == "lingauss":
        lm = LinearGaussian(n_vars=d)
        model = JointDiBS(likelihood_model=lm, **kw)
    else:
        lm = DenseNonlinearGaussian(n_vars=d, hidden_layers=(hidden,), activation=activation)
        model = JointDiBS(likelihood_model=lm, **kw)

    key = random.PRNGKey(seed + 11)
    key, subk = random.split(key)
    init = model._sample_initial_random_particles(key=subk, n_particles=m, n_dim=n_dim)
    z, theta = (init, None) if not joint else init
    # spread particles a little so that the fixture is not at the degenerate init scale
    z = z * 3.0
    k = z.shape[2]
    model.latent_prior_std = 1.0 / jax.numpy.sqrt(k)
    opt_init, model.opt_update, get_params = model.opt
    model.get_params = get_params
    sf = torch.tensor(rng.normal(size=m).astype(np.float32)) * (1.0 if baseline > 0 else 0.0)
    tt = torch.tensor(t, dtype=torch.int32)

    out = dict(x=data["x"], interv_mask=mask_np, z=npy(z), key=key_np(key), sf_baseline=npy(sf), t=np.int32(t),
               n_grad_mc_samples=np.int32(s), n_acyclicity_mc_samples=np.int32(a), hidden=np.int32(hidden),
               lik=np.array(lik), prior=np.array(prior), n_edges_per_node=np.int32(epn),
               estimator=np.array(model.grad_estimator_z), score_function_baseline=np.float32(baseline),
               optimizer=np.array(optimizer), alpha_linear=np.float32(model.alpha(1)), beta_linear=np.float32(beta_linear),
               tau=np.float32(tau), latent_prior_std=npy(model.latent_prior_std), activation=np.array(activation))
    if joint:
        out["theta"] = flat_theta(theta, lik)

    # --- per-function outputs ---------------------------------------------------------------
    out["edge_probs"] = npy(model.edge_probs(z, tt))
    out["g_lim"] = npy(model.particle_to_g_lim(z))
    subkeys = random.split(random.PRNGKey(seed + 5), m)
    out["mc_keys"] = key_np(subkeys)
    gs = torch.stack([model.sample_g(model.edge_probs(z[i], tt), subkeys[i], s) for i in range(m)])
    out["sample_g"] = npy(gs)
    eps = torch.stack([random.logistic(subkeys[i], shape=(s, d, d)) for i in range(m)])
    out["logistic_eps"] = npy(eps)
    soft = torch.stack([torch.stack([model.particle_to_soft_graph(z[i], eps[i, j], tt) for j in range(s)]) for i in range(m)])
    out["soft_g"] = npy(soft)
    th_i = (lambda i: None) if not joint else (lambda i: jax.tree_util.tree_map(lambda l: l[i], theta))
    out["logprob_hard"] = npy(torch.stack([model.eltwise_log_joint_prob(gs[i], th_i(i), None) for i in range(m)]))
    if joint or estimator == "reparam":
        out["logprob_soft"] = npy(torch.stack([model.eltwise_log_joint_prob(soft[i], th_i(i), None) for i in range(m)]))
    out["acyclic_h_hard"] = npy(torch.stack([acyclic_constr_nograd(gs[i, 0], d) for i in range(m)]))
    out["acyclic_h_soft"] = npy(torch.stack([acyclic_constr_nograd(soft[i, 0], d) for i in range(m)]))

    dz_lik, sf_new = model.eltwise_grad_z_likelihood(z, theta, sf, tt, subkeys)
    out["grad_z_likelihood"] = npy(dz_lik)
    out["sf_baseline_new"] = npy(sf_new)
    if joint:
        out["grad_theta_likelihood"] = flat_theta_grad(model.eltwise_grad_theta_likelihood(z, theta, tt, subkeys), lik)
    out["grad_latent_prior"] = npy(model.eltwise_grad_latent_prior(z, subkeys, tt))
    out["grad_constraint"] = npy(torch.stack([model.grad_constraint_gumbel(z[i], subkeys[i], tt) for i in range(m)]))

    # --- kernel, phi ---------------------------------------------------------------------------
    if joint:
        kxx = model._f_kernel_mat(z, theta, z, theta)
        out["kxx"] = npy(kxx)
        gz = torch.tensor(rng.normal(size=tuple(z.shape)).astype(np.float32))
        out["phi_in_grad_z"] = npy(gz)
        out["phi_z"] = npy(model._parallel_update_z(z, theta, kxx, z, theta, gz))
        gth = jax.tree_util.tree_map(lambda l: torch.tensor(rng.normal(size=tuple(l.shape)).astype(np.float32)), theta)
        out["phi_in_grad_theta"] = flat_theta(gth, lik)
        out["phi_theta"] = flat_theta(model._parallel_update_theta(z, theta, kxx, z, theta, gth), lik)
    else:
        kxx = model._f_kernel_mat(z, z)
        out["kxx"] = npy(kxx)
        gz = torch.tensor(rng.normal(size=tuple(z.shape)).astype(np.float32))
        out["phi_in_grad_z"] = npy(gz)
        out["phi_z"] = npy(model._parallel_update_z(z, kxx, z, gz))

    # --- two full steps from a warm optimizer state -----------------------------------------------
    opt_z = opt_init(z)
    if joint:
        opt_th = opt_init(theta)
        carry = (opt_z, opt_th, key, sf)
    else:
        carry = (opt_z, key, sf)
    for step in range(2):
        carry = model._svgd_step(torch.tensor(t + step, dtype=torch.int32), *carry)
        st = carry[0].tree
        out[f"step{step + 1}_z"] = npy(st[0])
        out[f"step{step + 1}_v_z"] = npy(st[1]) if len(st) > 1 else np.zeros_like(npy(st[0]))
        out[f"step{step + 1}_key"] = key_np(carry[-2])
        out[f"step{step + 1}_sf_baseline"] = npy(carry[-1])
        if joint:
            out[f"step{step + 1}_theta"] = flat_theta(get_params(carry[1]), lik)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(f"wrote {name}.npz ({len(out)} arrays)")


def make_sample_case(name, *, lik, prior, epn, d, n_obs, m, s, a, steps, callback_every=None, hidden=4, seed=0, n_dim=None):
    if lik == "densenn":
        data = make_nonlinear_gaussian_data(seed=seed, n_vars=d, n_observations=n_obs, n_edges_per_node=1, hidden=hidden)
    else:
        data = make_linear_gaussian_data(seed=seed, n_vars=d, n_observations=n_obs, n_edges_per_node=1)
    x = torch.tensor(data["x"])
    gm = make_prior(prior, d, epn)
    kw = dict(x=x, graph_model=gm, n_grad_mc_samples=s, n_acyclicity_mc_samples=a)
    if lik == "bge":
        model = MarginalDiBS(likelihood_model=BGe(n_vars=d), **kw)
    elif lik == "lingauss":
        model = JointDiBS(likelihood_model=LinearGaussian(n_vars=d), **kw)
    else:
        model = JointDiBS(likelihood_model=DenseNonlinearGaussian(n_vars=d, hidden_layers=(hidden,)), **kw)
    trace = {}

    def cb(**kwargs):
        trace[f"cb_t{kwargs['t']}_z"] = npy(kwargs["zs"])
        if "thetas" in kwargs:
            trace[f"cb_t{kwargs['t']}_theta"] = flat_theta(kwargs["thetas"], lik)

    res = model.sample(key=random.PRNGKey(seed), n_particles=m, steps=steps, n_dim_particles=n_dim,
                       callback=cb, callback_every=callback_every)
    out = dict(x=data["x"], seed=np.int32(seed), steps=np.int32(steps), n_particles=np.int32(m),
               callback_every=np.int32(callback_every or 0), n_grad_mc_samples=np.int32(s),
               n_acyclicity_mc_samples=np.int32(a), hidden=np.int32(hidden), lik=np.array(lik), prior=np.array(prior),
               n_edges_per_node=np.int32(epn), **trace)
    if lik == "bge":
        out["g_final"] = npy(res)
    else:
        out["g_final"] = npy(res[0])
        out["theta_final"] = flat_theta(res[1], lik)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(f"wrote {name}.npz ({len(out)} arrays)")


def make_prng_case():
    out = {}
    for seed in (0, 123, 2**33 + 7):
        k = random.PRNGKey(seed)
        out[f"key_{seed}"] = key_np(k)
        out[f"split5_{seed}"] = key_np(random.split(k, 5))
        out[f"uniform_{seed}"] = npy(random.uniform(k, (3, 7)))
        out[f"logistic_{seed}"] = npy(random.logistic(k, shape=(2, 3, 3)))
        out[f"normal_{seed}"] = npy(random.normal(k, shape=(11,)))
        out[f"bernoulli_{seed}"] = npy(random.bernoulli(k, p=torch.tensor(0.3), shape=(4, 5, 5)))
    np.savez_compressed(os.path.join(OUT, "prng.npz"), **out)
    print("wrote prng.npz")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    make_prng_case()
    make_case("step_marginal_bge_sf", lik="bge", prior="sf", d=5, m=3, s=8, a=4, t=3, alpha_linear=0.2)
    make_case("step_marginal_bge_er_interv_baseline", lik="bge", prior="er", epn=1, d=6, m=4, s=10, a=3, t=2,
              interv=True, baseline=0.1, seed=1, alpha_linear=0.3)
    make_case("step_joint_lingauss_er", lik="lingauss", prior="er", epn=1, d=5, m=3, s=8, a=4, t=7)
    make_case("step_joint_lingauss_sf_interv_gd", lik="lingauss", prior="sf", d=6, m=4, s=9, a=5, t=12, interv=True,
              optimizer="gd", seed=2, n_dim=4, tau=0.7, beta_linear=0.5)
    make_case("step_joint_lingauss_score", lik="lingauss", prior="uniform", d=4, m=3, s=8, a=4, t=5,
              estimator="score", baseline=0.05, seed=3)
    # BGe with the reparameterisation estimator (soft graphs, real-valued parent counts)
    make_case("step_marginal_bge_reparam", lik="bge", prior="er", epn=1, d=6, m=3, s=8, a=4, t=4, estimator="reparam",
              seed=6, alpha_linear=0.25)
    make_case("step_marginal_bge_reparam_interv", lik="bge", prior="sf", d=5, m=3, s=6, a=4, t=3, estimator="reparam",
              interv=True, seed=7, alpha_linear=0.3)
    make_case("step_joint_densenn_er", lik="densenn", prior="er", epn=1, d=5, m=3, s=8, a=4, t=6, hidden=4, seed=4)
    make_case("step_joint_densenn_sf_interv", lik="densenn", prior="sf", d=6, m=3, s=7, a=3, t=9, hidden=5, interv=True, seed=5)
    # the other activations of the reference's MLP (nonlinearGaussian.py:52-61)
    make_case("step_joint_densenn_tanh", lik="densenn", prior="er", epn=1, d=5, m=3, s=8, a=4, t=6, hidden=4, seed=8, activation="tanh")
    make_case("step_joint_densenn_sigmoid", lik="densenn", prior="sf", d=4, m=3, s=6, a=4, t=5, hidden=3, seed=9, activation="sigmoid")
    make_case("step_joint_densenn_leakyrelu", lik="densenn", prior="er", epn=1, d=5, m=3, s=8, a=4, t=8, hidden=5, interv=True, seed=10,
              activation="leakyrelu")
    # BASELINE.json configs[0]: MarginalDiBS BGe n_vars=5 n_particles=4 steps=50 (sf prior because ER p>=1 at d=5)
    make_sample_case("sample_c1_marginal_bge", lik="bge", prior="sf", epn=2, d=5, n_obs=100, m=4, s=128, a=32, steps=50,
                     callback_every=10)
    make_sample_case("sample_joint_lingauss", lik="lingauss", prior="er", epn=1, d=5, n_obs=50, m=4, s=16, a=8, steps=12,
                     callback_every=5, seed=1)
    make_sample_case("sample_joint_densenn", lik="densenn", prior="sf", epn=2, d=5, n_obs=40, m=3, s=8, a=4, steps=6,
                     callback_every=3, hidden=5, seed=2)
