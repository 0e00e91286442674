"""Particle distributions and evaluation metrics (host side, NumPy) -- the post-processing half of the reference's
notebook flow (dibs/metrics.py:12-268).  Same names, keyword-only signatures and edge-case behaviour; inputs may be
torch tensors (CUDA or CPU) or NumPy arrays, results are Python / NumPy scalars.  Nothing here is on the SVGD hot
path: ``M <= 4096`` graphs of ``d <= 128`` nodes are a few MB, and the acyclicity filter is an integer reachability
test (exactly ``h(G) == 0`` of dibs/graph_utils.py:8-30 for 0/1 graphs, without its fp32 matrix power).
"""
from typing import Any, NamedTuple

import numpy as np


class ParticleDistribution(NamedTuple):
    """Sampled particles (G, Theta) or G and their log weights (dibs/metrics.py:12-25)."""
    logp: Any
    g: Any
    theta: Any = None


def _np(a):
    if a is None:
        return None
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def _logsumexp(a, b=None, axis=0, return_sign=False):
    """jax.scipy.special.logsumexp with weights ``b``."""
    a = np.asarray(a, dtype=np.float64)
    amax = np.max(a, axis=axis, keepdims=True)
    amax = np.where(np.isfinite(amax), amax, 0.0)
    e = np.exp(a - amax)
    if b is not None:
        e = e * np.asarray(b, dtype=np.float64)
    s = e.sum(axis=axis)
    sign = np.sign(s)
    with np.errstate(divide="ignore"):
        out = np.log(np.abs(s)) + np.squeeze(amax, axis=axis)
    return (out, sign) if return_sign else out


def elwise_acyclic(g):
    """Boolean [M]: is each 0/1 adjacency matrix a DAG (== ``elwise_acyclic_constr_nograd(g, d) == 0``,
    dibs/graph_utils.py:30, evaluated exactly: repeated squaring of the boolean reachability matrix)."""
    g = (_np(g) != 0)
    m, d, _ = g.shape
    reach = g.copy()
    steps = 1
    while steps < d:
        reach = reach | (np.einsum("mik,mkj->mij", reach.astype(np.uint8), reach.astype(np.uint8)) > 0)
        steps *= 2
    return ~np.einsum("mii->mi", reach).any(axis=1)


def pairwise_structural_hamming_distance(*, x, y):
    """[N, M] SHD between two batches of adjacency matrices; reversals count once (dibs/metrics.py:28-53)."""
    x, y = _np(x), _np(y)
    assert x.ndim == 3 and y.ndim == 3
    pw = np.abs(x[:, None].astype(np.int64) - y[None].astype(np.int64))
    pw = pw + pw.transpose(0, 1, 3, 2)
    pw = np.where(pw > 1, 1, pw)
    return pw.sum(axis=(2, 3)) / 2


def _select_dags(dist):
    g, logp = _np(dist.g), np.asarray(_np(dist.logp), dtype=np.float64)
    is_dag = elwise_acyclic(g)
    if is_dag.sum() == 0:
        return g, logp, is_dag, None
    lw = logp[is_dag] - _logsumexp(logp[is_dag])
    return g, logp, is_dag, lw


def expected_shd(*, dist, g):
    """sum_G p(G|D) SHD(G, G*) over the acyclic particles (dibs/metrics.py:56-88)."""
    g_true = _np(g)
    n_vars = g_true.shape[0]
    gs, _, is_dag, lw = _select_dags(dist)
    if lw is None:
        return n_vars * (n_vars - 1) / 2              # "wrong on every edge"
    shds = pairwise_structural_hamming_distance(x=gs[is_dag], y=g_true[None])[:, 0]
    val, sgn = _logsumexp(lw, b=shds, axis=0, return_sign=True)
    return float(sgn * np.exp(val))


def expected_edges(*, dist):
    """sum_G p(G|D) |edges(G)| (dibs/metrics.py:91-128)."""
    gs, logp, is_dag, lw = _select_dags(dist)
    if lw is None:
        val, sgn = _logsumexp(logp, b=gs.sum(axis=(-1, -2)), axis=0, return_sign=True)
        return float(sgn * np.exp(val))
    val, sgn = _logsumexp(lw, b=gs[is_dag].sum(axis=(-1, -2)), axis=0, return_sign=True)
    return float(sgn * np.exp(val))


def edge_marginals(*, dist):
    """P(G_ij = 1) under the particle distribution restricted to acyclic particles, [d, d]."""
    gs, _, is_dag, lw = _select_dags(dist)
    if lw is None:
        return None
    val, sgn = _logsumexp(lw[:, None, None], b=gs[is_dag], axis=0, return_sign=True)
    return sgn * np.exp(val)


def threshold_metrics(*, dist, g):
    """ROC / precision-recall metrics of the edge marginals against the ground truth (dibs/metrics.py:131-185)."""
    from sklearn import metrics as sklearn_metrics
    g_true = _np(g)
    n_vars = g_true.shape[0]
    p_edge = edge_marginals(dist=dist)
    if p_edge is None:
        base = float(g_true.sum() / (n_vars * (n_vars - 1)))
        return {"roc_auc": 0.5, "prc_auc": base, "ave_prec": base}
    g_flat, p_flat = g_true.reshape(-1), p_edge.reshape(-1)
    fpr_, tpr_, _ = sklearn_metrics.roc_curve(g_flat, p_flat)
    precision_, recall_, _ = sklearn_metrics.precision_recall_curve(g_flat, p_flat)
    return {
        "fpr": fpr_.tolist(), "tpr": tpr_.tolist(), "roc_auc": sklearn_metrics.auc(fpr_, tpr_),
        "precision": precision_.tolist(), "recall": recall_.tolist(),
        "prc_auc": sklearn_metrics.auc(recall_, precision_),
        "ave_prec": sklearn_metrics.average_precision_score(g_flat, p_flat),
    }


def neg_ave_log_marginal_likelihood(*, dist, eltwise_log_marginal_likelihood, x):
    """- sum_G p(G|D) log p(D_test | G); ``eltwise_log_marginal_likelihood([:,d,d], [N,d]) -> [:]`` (dibs/metrics.py:188-225)."""
    n_vars = _np(x).shape[1]
    gs, _, is_dag, lw = _select_dags(dist)
    if lw is None:
        gsel, lw = np.zeros((1, n_vars, n_vars), gs.dtype), np.zeros(1)      # score as the empty graph only
    else:
        gsel = gs[is_dag]
    ll = np.asarray(_np(eltwise_log_marginal_likelihood(gsel, x)), dtype=np.float64)
    val, sgn = _logsumexp(lw, b=ll, axis=0, return_sign=True)
    return float(-sgn * np.exp(val))


def neg_ave_log_likelihood(*, dist, eltwise_log_likelihood, x):
    """- sum_(G,Theta) p(G,Theta|D) log p(D_test | G, Theta); theta is indexed along its leading axis
    (a tensor / array, or the stax-style nested list of tensors) (dibs/metrics.py:228-268)."""
    assert dist.theta is not None
    gs, logp, is_dag, lw = _select_dags(dist)

    def pick(t, f):
        if isinstance(t, (list, tuple)):
            return type(t)(pick(u, f) for u in t)
        return f(t)

    if lw is None:
        gsel, theta, lw = gs * 0, pick(dist.theta, lambda t: t * 0.0), logp * 0.0
    else:
        idx = np.nonzero(is_dag)[0]
        gsel, theta = gs[is_dag], pick(dist.theta, lambda t: t[idx])
    ll = np.asarray(_np(eltwise_log_likelihood(gsel, theta, x)), dtype=np.float64)
    val, sgn = _logsumexp(lw, b=ll, axis=0, return_sign=True)
    return float(-sgn * np.exp(val))
