"""Synthetic Erdos-Renyi Gaussian Bayes-net data for tests and benchmarks (host side, NumPy).

Restates the *shape* of the reference's data factory (dibs/target.py:43-119 with
dibs/models/graph.py:44-53 ER DAG sampling, dibs/models/linearGaussian.py:212-272 and
dibs/models/nonlinearGaussian.py:155-242 ancestral sampling) with ``numpy.random`` instead of the
JAX PRNG: bit parity of the *data* with the reference is not needed because the same ``x`` is fed to
both sides of every comparison.
"""
import numpy as np


def sample_er_dag(rng, n_vars, n_edges_per_node=2):
    """Lower-triangular Bernoulli(p) adjacency, randomly permuted (graph.py:44-53)."""
    p = min(1.0, (n_edges_per_node * n_vars) / ((n_vars * (n_vars - 1)) / 2))
    mat = (rng.random((n_vars, n_vars)) < p).astype(np.int32)
    dag = np.tril(mat, k=-1)
    perm = np.eye(n_vars, dtype=np.int32)[rng.permutation(n_vars)]
    return perm.T @ dag @ perm


def toporder(g):
    g = np.array(g, copy=True)
    d = g.shape[0]
    order, indeg = [], g.sum(axis=0)
    ready = [j for j in range(d) if indeg[j] == 0]
    while ready:
        j = ready.pop()
        order.append(j)
        for c in np.nonzero(g[j])[0]:
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(int(c))
    assert len(order) == d, "graph is not a DAG"
    return order


def make_linear_gaussian_data(seed=0, n_vars=20, n_observations=100, n_edges_per_node=2,
                              obs_noise=0.1, mean_edge=0.0, sig_edge=1.0, min_edge=0.5):
    """Returns dict(g, theta, x) for a linear-Gaussian SEM on an ER DAG (fp32 ``x[N, d]``)."""
    rng = np.random.default_rng(seed)
    g = sample_er_dag(rng, n_vars, n_edges_per_node)
    theta = mean_edge + sig_edge * rng.standard_normal((n_vars, n_vars))
    theta = theta + np.sign(theta) * min_edge
    z = np.sqrt(obs_noise) * rng.standard_normal((n_observations, n_vars))
    x = np.zeros((n_observations, n_vars))
    for j in toporder(g):
        pa = np.nonzero(g[:, j])[0]
        x[:, j] = (x[:, pa] @ theta[pa, j] if len(pa) else 0.0) + z[:, j]
    return dict(g=g, theta=theta.astype(np.float32), x=x.astype(np.float32))


def make_nonlinear_gaussian_data(seed=0, n_vars=20, n_observations=100, n_edges_per_node=2,
                                 obs_noise=0.1, sig_param=1.0, hidden=5):
    """Returns dict(g, x) for a one-hidden-layer ReLU-MLP SEM on an ER DAG."""
    rng = np.random.default_rng(seed)
    g = sample_er_dag(rng, n_vars, n_edges_per_node)
    w1 = sig_param * rng.standard_normal((n_vars, n_vars, hidden))
    b1 = sig_param * rng.standard_normal((n_vars, hidden))
    w2 = sig_param * rng.standard_normal((n_vars, hidden))
    b2 = sig_param * rng.standard_normal((n_vars,))
    z = np.sqrt(obs_noise) * rng.standard_normal((n_observations, n_vars))
    x = np.zeros((n_observations, n_vars))
    for j in toporder(g):
        if g[:, j].sum() > 0:
            pre = (x * g[:, j][None]) @ w1[j] + b1[j]
            x[:, j] = np.maximum(pre, 0) @ w2[j] + b2[j] + z[:, j]
        else:
            x[:, j] = z[:, j]
    return dict(g=g, x=x.astype(np.float32))
