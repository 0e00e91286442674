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


def sample_sf_dag(rng, n_vars, n_edges_per_node=2):
    """Scale-free DAG: Barabasi-Albert preferential attachment (every new node sends ``n_edges_per_node`` edges to
    existing nodes with probability proportional to their degree), then a random relabelling -- the construction the
    reference gets from ``igraph.Graph.Barabasi(directed=True)`` + permutation (graph.py:132-151)."""
    m = int(n_edges_per_node)
    g = np.zeros((n_vars, n_vars), dtype=np.int32)
    deg = np.zeros(n_vars)
    for v in range(1, n_vars):
        k = min(m, v)
        w = deg[:v] + 1.0
        targets = rng.choice(v, size=k, replace=False, p=w / w.sum())
        for u in targets:
            g[v, u] = 1          # new -> old: acyclic by construction
            deg[u] += 1
            deg[v] += 1
    perm = np.eye(n_vars, dtype=np.int32)[rng.permutation(n_vars)]
    return perm.T @ g @ perm


def sample_linear_gaussian_parameters(rng, n_vars, mean_edge=0.0, sig_edge=1.0, min_edge=0.5):
    """Theta ~ N(mean_edge, sig_edge^2), shifted away from zero by min_edge (linearGaussian.py:212-227)."""
    theta = mean_edge + sig_edge * rng.standard_normal((n_vars, n_vars))
    return (theta + np.sign(theta) * min_edge).astype(np.float32)


def sample_obs_linear_gaussian(rng, g, theta, n_samples, obs_noise=0.1, interv=None):
    """Ancestral sampling of x = x (G o Theta) + noise; ``interv`` = {node: value} clamps nodes (linearGaussian.py:230-272)."""
    interv = interv or {}
    d = g.shape[0]
    z = np.sqrt(obs_noise) * rng.standard_normal((n_samples, d))
    x = np.zeros((n_samples, d))
    for j in toporder(g):
        if j in interv:
            x[:, j] = interv[j]
            continue
        pa = np.nonzero(g[:, j])[0]
        x[:, j] = (x[:, pa] @ theta[pa, j] if len(pa) else 0.0) + z[:, j]
    return x.astype(np.float32)


def sample_dense_nn_parameters(rng, n_vars, hidden, sig_param=1.0):
    """One-hidden-layer MLP per node, all weights and biases ~ N(0, sig_param^2), in the stax pytree layout
    [(W1[d,d,H], b1[d,H]), (), (W2[d,H,1], b2[d,1])] (nonlinearGaussian.py:155-186)."""
    w1 = sig_param * rng.standard_normal((n_vars, n_vars, hidden))
    b1 = sig_param * rng.standard_normal((n_vars, hidden))
    w2 = sig_param * rng.standard_normal((n_vars, hidden, 1))
    b2 = sig_param * rng.standard_normal((n_vars, 1))
    return [(w1.astype(np.float32), b1.astype(np.float32)), (), (w2.astype(np.float32), b2.astype(np.float32))]


def sample_obs_dense_nn(rng, g, theta, n_samples, obs_noise=0.1, interv=None):
    """Ancestral sampling of x_j = MLP_j(x o G[:, j]) + noise (nonlinearGaussian.py:189-242)."""
    interv = interv or {}
    (w1, b1), _, (w2, b2) = theta
    d = g.shape[0]
    z = np.sqrt(obs_noise) * rng.standard_normal((n_samples, d))
    x = np.zeros((n_samples, d))
    for j in toporder(g):
        if j in interv:
            x[:, j] = interv[j]
        elif g[:, j].sum() > 0:
            pre = (x * g[:, j][None]) @ w1[j] + b1[j]
            x[:, j] = (np.maximum(pre, 0) @ w2[j])[:, 0] + b2[j, 0] + z[:, j]
        else:
            x[:, j] = z[:, j]
    return x.astype(np.float32)


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
