"""Shared helpers for the parity tests: load golden fixtures, build oracle configs."""
import os

import numpy as np

from oracle import dibs_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STEP_CASES = [
    "step_marginal_bge_sf",
    "step_marginal_bge_er_interv_baseline",
    "step_marginal_bge_reparam",
    "step_marginal_bge_reparam_interv",
    "step_joint_lingauss_er",
    "step_joint_lingauss_sf_interv_gd",
    "step_joint_lingauss_score",
    "step_joint_densenn_er",
    "step_joint_densenn_sf_interv",
    "step_joint_densenn_tanh",
    "step_joint_densenn_sigmoid",
    "step_joint_densenn_leakyrelu",
]
SAMPLE_CASES = ["sample_c1_marginal_bge", "sample_joint_lingauss", "sample_joint_densenn"]


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def oracle_config(g, sample_case=False):
    """Fixture dict -> oracle Config (constructor defaults of the reference classes where not stored)."""
    lik_kind = str(g["lik"])
    d = g["x"].shape[1]
    lik = orc.Likelihood(kind=lik_kind, n_vars=d, hidden=int(g["hidden"]),
                         activation=str(g["activation"]) if "activation" in g else "relu")
    prior = orc.GraphPrior(kind=str(g["prior"]), n_vars=d, n_edges_per_node=int(g["n_edges_per_node"]))
    joint = lik_kind != "bge"
    cfg = orc.Config(lik=lik, prior=prior, joint=joint,
                     n_grad_mc_samples=int(g["n_grad_mc_samples"]),
                     n_acyclicity_mc_samples=int(g["n_acyclicity_mc_samples"]))
    if sample_case:
        cfg.alpha_linear = 0.05 if joint else 1.0
        cfg.grad_estimator_z = "reparam" if joint else "score"
        return cfg
    cfg.alpha_linear = float(g["alpha_linear"])
    cfg.beta_linear = float(g["beta_linear"])
    cfg.tau = float(g["tau"])
    cfg.grad_estimator_z = str(g["estimator"])
    cfg.score_function_baseline = float(g["score_function_baseline"])
    cfg.optimizer = str(g["optimizer"])
    cfg.latent_prior_std = float(g["latent_prior_std"])
    return cfg


def state_from(g, dt):
    z = g["z"].astype(dt)
    theta = g["theta"].astype(dt) if "theta" in g else None
    return orc.State(z=z, v_z=np.zeros_like(z), key=g["key"].astype(np.uint32),
                     sf_baseline=g["sf_baseline"].astype(dt), theta=theta,
                     v_theta=None if theta is None else np.zeros_like(theta),
                     latent_prior_std=float(g["latent_prior_std"]))


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: {bad.sum()}/{bad.size} outside tol; worst at {i}: got {a[i]!r} want {b[i]!r} "
                             f"(|err|={err[i]:.3e}, tol={tol[i]:.3e})")
