"""fp64 NumPy checks of the algebraic rewrites the CUDA kernels rely on -- each identity is what
lets a kernel do less work than the reference's formulation while producing the same numbers."""
import numpy as np
import pytest


def _spd(rng, d, n=60):
    x = rng.normal(size=(n, d))
    return x.T @ x + 0.5 * np.eye(d)


@pytest.mark.parametrize("d", [6, 20, 50])
def test_bge_certain_uncertain_split(d):
    """kernels_mc_bge.cuh: with C = parents present in every sample and U = parents present in some,
    det R[C+Us] = det R[C,C] * det W[Us]  and  schur_j(R; C+Us) = schur_j(W; Us),  W = Schur complement of C in R[C+U+j]."""
    rng = np.random.default_rng(d)
    r = _spd(rng, d)
    j = d - 1
    cand = np.arange(d - 1)
    c_set = cand[rng.random(d - 1) < 0.4]
    u_set = np.array([i for i in cand if i not in c_set and rng.random() < 0.5], dtype=int)
    rows = np.concatenate([c_set, u_set, [j]]).astype(int)
    a = r[np.ix_(rows, rows)]
    nc = len(c_set)
    if nc:
        w = a[nc:, nc:] - a[nc:, :nc] @ np.linalg.solve(a[:nc, :nc], a[:nc, nc:])
        logdet_c = np.linalg.slogdet(a[:nc, :nc])[1]
    else:
        w, logdet_c = a, 0.0
    for _ in range(8):
        pick = rng.random(len(u_set)) < 0.5
        us = u_set[pick]
        p = np.concatenate([c_set, us]).astype(int)
        pj = np.concatenate([p, [j]]).astype(int)
        ref_pp = np.linalg.slogdet(r[np.ix_(p, p)])[1] if len(p) else 0.0
        ref_pj = np.linalg.slogdet(r[np.ix_(pj, pj)])[1]
        wi = np.nonzero(pick)[0]
        got_pp = logdet_c + (np.linalg.slogdet(w[np.ix_(wi, wi)])[1] if len(wi) else 0.0)
        wj = np.concatenate([wi, [len(u_set)]]).astype(int)
        got_pj = logdet_c + np.linalg.slogdet(w[np.ix_(wj, wj)])[1]
        assert abs(got_pp - ref_pp) < 1e-9 * max(1.0, abs(ref_pp))
        assert abs(got_pj - ref_pj) < 1e-9 * max(1.0, abs(ref_pj))


@pytest.mark.parametrize("n,d", [(100, 20), (30, 36), (100, 100)])
def test_lingauss_qr_form(n, d):
    """kernels_mc_lin_qr.cuh / kernels_dense.cuh: with x = Q Rx and u_j = e_j - (G o Theta)_:j,
    sum_n R_nj^2 = |Rx u_j|^2  and  x^T R = Rx^T (Rx U)  for the residual R = x - x (G o Theta) (also when N < d)."""
    rng = np.random.default_rng(n + d)
    x = rng.normal(size=(n, d))
    g = (rng.random((d, d)) < 0.3).astype(float); np.fill_diagonal(g, 0)
    th = rng.normal(size=(d, d))
    resid = x - x @ (g * th)
    rx = np.linalg.qr(x, mode="r")                      # [min(n,d), d]
    rxp = np.zeros((d, d)); rxp[:rx.shape[0]] = rx      # rows beyond min(N, d) are zero (qr_upper_packed)
    u = np.eye(d) - g * th
    y = rxp @ u
    assert np.allclose((y ** 2).sum(0), (resid ** 2).sum(0), rtol=1e-10, atol=1e-10)
    assert np.allclose(rxp.T @ y, x.T @ resid, rtol=1e-9, atol=1e-9)


def test_soft_graph_closed_form():
    """entry_from_bits: sigmoid(log(u / (1 - u)) + a) = u / (u + (1 - u) e^-a)  (tau = 1: logistic noise and sigmoid cancel)."""
    rng = np.random.default_rng(0)
    u = rng.uniform(1e-7, 1 - 1e-7, 10000)
    a = rng.normal(scale=8.0, size=10000)
    ref = 1.0 / (1.0 + np.exp(-(np.log(u) - np.log1p(-u) + a)))
    got = u / (u + (1.0 - u) * np.exp(-a))
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-300)


def test_acyclicity_gradient_through_transposed_power():
    """kernels_acyclic.cuh / kernels_dense.cuh: d/dG tr((I + G/d)^d) = ((I + G/d)^(d-1))^T, so
    dS[a][b] = E[b][a] * tau alpha g_ab (1 - g_ab) with E = M^(d-1)."""
    rng = np.random.default_rng(1)
    d = 7
    g = rng.uniform(0, 1, (d, d)); np.fill_diagonal(g, 0)
    h = lambda gg: np.trace(np.linalg.matrix_power(np.eye(d) + gg / d, d)) - d
    e = np.linalg.matrix_power(np.eye(d) + g / d, d - 1)
    num = np.zeros((d, d))
    for a in range(d):
        for b in range(d):
            gp = g.copy(); gp[a, b] += 1e-6
            gm = g.copy(); gm[a, b] -= 1e-6
            num[a, b] = (h(gp) - h(gm)) / 2e-6
    assert np.allclose(num, e.T, rtol=1e-6, atol=1e-8)


def test_phi_slices_and_difference_form():
    """kernels_pair.cuh: phi_i = -(1/M) sum_j [K_ij g_j - (2/h) K_ij (x_j - x_i)] equals the reference's
    -(1/M)(K^T g + grad_x k) (svgd.py:194-216), and summing fixed j slices in order reproduces the full sum."""
    rng = np.random.default_rng(2)
    m, dd, hh = 12, 9, 5.0
    x = rng.normal(size=(m, dd)); gr = rng.normal(size=(m, dd))
    k = np.exp(-((x[:, None] - x[None]) ** 2).sum(-1) / hh)
    ref = -(k.T @ gr + (-(2 / hh)) * (k.T @ x - k.sum(0)[:, None] * x)) / m
    parts = []
    for j0 in range(0, m, 4):
        sl = slice(j0, j0 + 4)
        drive = k[:, sl] @ gr[sl]
        rep = np.einsum("ij,ijc->ic", k[:, sl], x[None, sl] - x[:, None])
        parts.append(drive - (2 / hh) * rep)
    got = -sum(parts) / m
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
