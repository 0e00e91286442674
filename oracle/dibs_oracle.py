"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the DiBS SVGD particle update.

This is the CPU checker for the CUDA hot path in ``dibs_b200``; it is never
imported by the product.  It restates, function by function, the reference
(larslorch/dibs @ 5350d1a, paths relative to /root/reference/):

  dibs/inference/svgd.py:125-148,226-331   MarginalDiBS init / step / sample
  dibs/inference/svgd.py:489-515,673-795   JointDiBS    init / step / sample
  dibs/inference/dibs.py:84-184            particle_to_g_lim, sample_g, soft graph, edge_probs
  dibs/inference/dibs.py:325-459           score-function and reparam estimators of grad_Z
  dibs/inference/dibs.py:488-551           grad_Theta estimator (hard graphs)
  dibs/inference/dibs.py:557-658           acyclicity-constraint gradient, latent prior score
  dibs/kernel.py:20-30,52-71               SE kernels
  dibs/graph_utils.py:8-28                 h(G) = tr((I+G/d)^d) - d
  dibs/utils/func.py:117-145               zero_diagonal, masked slogdet
  dibs/models/linearGaussian.py:63-170     BGe marginal likelihood
  dibs/models/linearGaussian.py:212-227,278-338   LinearGaussian
  dibs/models/nonlinearGaussian.py:155-186,248-326 DenseNonlinearGaussian (one hidden layer)
  dibs/models/graph.py:27-30,93-108,182-196,263-276  soft graph priors

Every ``jax.grad`` of the reference is replaced by its closed form (SURVEY.md
App. B); the closed forms are checked against the reference's autodiff through
the golden fixtures (tests/golden/, made by oracle/gen_golden.py which runs the
unmodified reference sources on oracle/jaxshim).  JAX primitives that live
outside /root/reference (PRNG, optimizers, stax init, logsumexp, norm.logpdf,
matrix_power) are restated from the published JAX sources; see oracle/threefry.py.

All functions take ``dt`` (np.float32 = the reference's precision, np.float64 =
error-bound companion).  Random bits are always the JAX fp32 stream.
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
from scipy.special import gammaln

from . import threefry as tf

# --------------------------------------------------------------------------------------
# model descriptors (plain data; mirror the reference constructors)
# --------------------------------------------------------------------------------------


@dataclass
class GraphPrior:
    kind: str = "er"          # 'er' | 'sf' | 'uniform'
    n_vars: int = 0
    n_edges_per_node: int = 2

    @property
    def p(self):  # graph.py:27-30
        return (self.n_edges_per_node * self.n_vars) / ((self.n_vars * (self.n_vars - 1)) / 2)


@dataclass
class Likelihood:
    kind: str = "lingauss"    # 'bge' | 'lingauss' | 'densenn'
    n_vars: int = 0
    # BGe (linearGaussian.py:35-48)
    alpha_mu: float = 1.0
    alpha_lambd: Optional[float] = None
    mean_obs: Optional[np.ndarray] = None
    # LinearGaussian (linearGaussian.py:190-195)
    obs_noise: float = 0.1
    mean_edge: float = 0.0
    sig_edge: float = 1.0
    min_edge: float = 0.5
    # DenseNonlinearGaussian (nonlinearGaussian.py:105-111)
    hidden: int = 5
    sig_param: float = 1.0
    activation: str = "relu"  # 'relu' | 'tanh' | 'sigmoid' | 'leakyrelu' (stax, nonlinearGaussian.py:52-61)

    def theta_dim(self):
        d = self.n_vars
        if self.kind == "lingauss":
            return d * d
        if self.kind == "densenn":
            return d * (d * self.hidden + self.hidden + self.hidden + 1)
        return 0


@dataclass
class Config:
    lik: Likelihood
    prior: GraphPrior
    joint: bool = True
    alpha_linear: float = 0.05
    beta_linear: float = 1.0
    tau: float = 1.0
    n_grad_mc_samples: int = 128
    n_acyclicity_mc_samples: int = 32
    grad_estimator_z: str = "reparam"
    score_function_baseline: float = 0.0
    latent_prior_std: Optional[float] = None
    # kernel (svgd.py:81,446; kernel.py)
    h_latent: float = 5.0
    h_theta: float = 500.0
    scale_latent: float = 1.0
    scale_theta: float = 1.0
    # optimizer (svgd.py:83,117-120)
    optimizer: str = "rmsprop"
    stepsize: float = 0.005
    partitionable: bool = False


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def zero_diagonal(g):  # utils/func.py:117-125
    g = np.array(g, copy=True)
    d = g.shape[-1]
    g[..., np.arange(d), np.arange(d)] = 0
    return g


def norm_logpdf(x, loc, scale, dt):  # jax.scipy.stats.norm.logpdf
    scale = dt(scale)
    s2 = dt(scale * scale)
    return (-(np.log(dt(2.0 * np.pi) * s2) + (x - dt(loc)) ** 2 / s2) / dt(2.0)).astype(dt)


def logsumexp(a, axis=None, b=None, return_sign=False):  # jax.scipy.special.logsumexp
    a = np.asarray(a)
    if b is not None:
        a, b = np.broadcast_arrays(a, np.asarray(b, dtype=a.dtype))
        a = np.where(b != 0, a, -np.inf)
    amax = np.max(a, axis=axis, keepdims=True)
    amax = np.where(np.isfinite(amax), amax, 0).astype(a.dtype)
    ex = np.exp(a - amax)
    if b is not None:
        ex = ex * b
    s = np.sum(ex, axis=axis, keepdims=True)
    sign = np.sign(s)
    out = np.log(np.abs(s)) + amax
    out = np.squeeze(out, axis=axis) if axis is not None else out.reshape(())
    sign = np.squeeze(sign, axis=axis) if axis is not None else sign.reshape(())
    return (out, sign) if return_sign else out


def matrix_power(a, n):  # jnp.linalg.matrix_power multiplication order
    if n == 0:
        return np.broadcast_to(np.eye(a.shape[-1], dtype=a.dtype), a.shape).copy()
    if n == 1:
        return a
    if n == 2:
        return a @ a
    if n == 3:
        return (a @ a) @ a
    z = result = None
    while n > 0:
        z = a if z is None else z @ z
        n, bit = divmod(n, 2)
        if bit:
            result = z if result is None else result @ z
    return result


def alpha_of(cfg, t, dt):  # dibs.py:70
    return dt(dt(cfg.alpha_linear) * dt(t))


def beta_of(cfg, t, dt):  # dibs.py:71
    return dt(dt(cfg.beta_linear) * dt(t))


# --------------------------------------------------------------------------------------
# generative graph model p(G | Z)
# --------------------------------------------------------------------------------------


def scores_uv(z):
    """``einsum('...ik,...jk->...ij', u, v)`` (dibs.py:94-95,179-180)."""
    return np.einsum("...ik,...jk->...ij", z[..., 0], z[..., 1])


def edge_probs(z, alpha):  # dibs.py:168-184
    return zero_diagonal(sigmoid(alpha * scores_uv(z))).astype(z.dtype)


def particle_to_g_lim(z):  # dibs.py:84-99
    return zero_diagonal((scores_uv(z) > 0).astype(np.int32))


def sample_g(p, key, n_samples, partitionable=False):  # dibs.py:102-119
    d = p.shape[-1]
    u = tf.uniform(key, (n_samples, d, d), partitionable=partitionable)
    return zero_diagonal((u < p.astype(np.float32)).astype(np.int32))


def soft_graph(z, eps, alpha, tau):  # dibs.py:121-140
    dt = z.dtype.type
    return zero_diagonal(sigmoid(dt(tau) * (eps.astype(dt) + alpha * scores_uv(z)))).astype(dt)


def ds_to_dz(ds, z):
    """Chain rule through S = U V^T:  dU = dS V, dV = dS^T U  ->  [.., d, k, 2]."""
    u, v = z[..., 0], z[..., 1]
    du = np.einsum("...ij,...jk->...ik", ds, v)
    dv = np.einsum("...ij,...ik->...jk", ds, u)
    return np.stack([du, dv], axis=-1)


# --------------------------------------------------------------------------------------
# likelihoods: value and closed-form gradients, batched over a leading sample axis
# --------------------------------------------------------------------------------------


def lingauss_logjoint(g, theta, x, mask, lik, dt, want_grads=True):
    """LinearGaussian.interventional_log_joint_prob (linearGaussian.py:278-338) for g[S,d,d].

    Returns lp[S], dlp/dG[S,d,d], dlp/dTheta[S,d,d].
    """
    g = g.astype(dt)
    theta = theta.astype(dt)
    x = x.astype(dt)
    scale = dt(np.sqrt(dt(lik.obs_noise)))
    s2 = dt(scale * scale)
    w = g * theta
    mean = x @ w                                        # [S,N,d]
    res = np.where(mask[None].astype(bool), dt(0), x[None] - mean)
    logp_all = -(np.log(dt(2.0 * np.pi) * s2) + res * res / s2) / dt(2.0)
    ll = np.where(mask[None].astype(bool), dt(0), logp_all).sum(axis=(1, 2))
    lp_theta = norm_logpdf(theta, lik.mean_edge, lik.sig_edge, dt)
    prior = (g * lp_theta).sum(axis=(1, 2))
    lp = (prior + ll).astype(dt)
    if not want_grads:
        return lp, None, None
    b = np.einsum("ni,snj->sij", x, res)               # x^T R
    dg = lp_theta[None] + theta[None] * b / s2
    sig2 = dt(dt(lik.sig_edge) * dt(lik.sig_edge))
    dtheta = g * (-(theta - dt(lik.mean_edge)) / sig2)[None] + g * b / s2
    return lp, dg.astype(dt), dtheta.astype(dt)


def nn_unpack(theta_flat, d, h):
    """Flat per-particle layout used by dibs_b200: [W1(d,d,H) | b1(d,H) | W2(d,H) | b2(d)]."""
    o = 0
    w1 = theta_flat[..., o:o + d * d * h].reshape(theta_flat.shape[:-1] + (d, d, h)); o += d * d * h
    b1 = theta_flat[..., o:o + d * h].reshape(theta_flat.shape[:-1] + (d, h)); o += d * h
    w2 = theta_flat[..., o:o + d * h].reshape(theta_flat.shape[:-1] + (d, h)); o += d * h
    b2 = theta_flat[..., o:o + d].reshape(theta_flat.shape[:-1] + (d,))
    return w1, b1, w2, b2


def nn_pack(w1, b1, w2, b2):
    lead = w1.shape[:-3]
    return np.concatenate([w1.reshape(lead + (-1,)), b1.reshape(lead + (-1,)),
                           w2.reshape(lead + (-1,)), b2.reshape(lead + (-1,))], axis=-1)


def densenn_logjoint(g, theta_flat, x, mask, lik, dt, want_grads=True):
    """DenseNonlinearGaussian.interventional_log_joint_prob (nonlinearGaussian.py:248-326), one ReLU hidden layer.

    g[S,d,d]; theta_flat[Dtheta].  Returns lp[S], dlp/dG[S,d,d], dlp/dTheta[S,Dtheta].
    W1[j,i,h] is masked by G[i,j] (``g.T[:, :, None]``, nonlinearGaussian.py:266,291).
    """
    d, h = lik.n_vars, lik.hidden
    g = g.astype(dt)
    x = x.astype(dt)
    w1, b1, w2, b2 = [a.astype(dt) for a in nn_unpack(theta_flat, d, h)]
    scale = dt(np.sqrt(dt(lik.obs_noise)))
    s2 = dt(scale * scale)
    gt = np.swapaxes(g, 1, 2)                           # gt[s,j,i] = g[s,i,j]
    w1m = gt[..., None] * w1[None]                      # [S,j,i,h]
    pre = np.einsum("ni,sjih->sjnh", x, w1m) + b1[None, :, None, :]
    kind = getattr(lik, "activation", "relu")
    if kind == "relu":
        act, dact = np.maximum(pre, dt(0)), (pre > 0).astype(dt)
    elif kind == "tanh":
        act = np.tanh(pre); dact = 1 - act * act
    elif kind == "sigmoid":
        act = sigmoid(pre); dact = act * (1 - act)
    elif kind == "leakyrelu":
        act, dact = np.where(pre >= 0, pre, dt(0.01) * pre), np.where(pre >= 0, dt(1), dt(0.01))
    else:
        raise KeyError(f"Invalid activation function `{kind}`")
    mean = np.einsum("sjnh,jh->snj", act, w2) + b2[None, None, :]
    res = np.where(mask[None].astype(bool), dt(0), x[None] - mean)
    logp_all = -(np.log(dt(2.0 * np.pi) * s2) + res * res / s2) / dt(2.0)
    ll = np.where(mask[None].astype(bool), dt(0), logp_all).sum(axis=(1, 2))
    lp_w1 = norm_logpdf(w1, 0.0, lik.sig_param, dt)
    prior = (gt[..., None] * lp_w1[None]).sum(axis=(1, 2, 3)) \
        + norm_logpdf(b1, 0.0, lik.sig_param, dt).sum() \
        + norm_logpdf(w2, 0.0, lik.sig_param, dt).sum() \
        + norm_logpdf(b2, 0.0, lik.sig_param, dt).sum()
    lp = (prior + ll).astype(dt)
    if not want_grads:
        return lp, None, None
    sp2 = dt(dt(lik.sig_param) * dt(lik.sig_param))
    delta = res / s2                                    # [S,n,j]
    dpre = np.einsum("snj,jh->sjnh", delta, w2) * dact
    dw1 = np.einsum("ni,sjnh->sjih", x, dpre) * gt[..., None] + gt[..., None] * (-w1 / sp2)[None]
    db1 = dpre.sum(axis=2) - (b1 / sp2)[None]
    dw2 = np.einsum("snj,sjnh->sjh", delta, act) - (w2 / sp2)[None]
    db2 = delta.sum(axis=1) - (b2 / sp2)[None]
    dgt = np.einsum("ni,sjnh,jih->sji", x, dpre, w1) + lp_w1.sum(axis=-1)[None]
    dg = np.swapaxes(dgt, 1, 2)
    return lp, dg.astype(dt), nn_pack(dw1, db1, dw2, db2).astype(dt)


def bge_precompute(x, mask, lik, dt):
    """Per-node R_j, N_j (linearGaussian.py:78-94).  Depends only on the data (and on j via interventions)."""
    x = x.astype(dt)
    n_obs, d = x.shape
    alpha_lambd = lik.alpha_lambd if lik.alpha_lambd else d + 2
    alpha_mu = lik.alpha_mu if lik.alpha_mu else 1.0
    mean_obs = np.zeros(d, dt) if lik.mean_obs is None else np.asarray(lik.mean_obs, dt)
    small_t = dt((alpha_mu * (alpha_lambd - d - 1)) / (alpha_mu + 1))
    r_all = np.zeros((d, d, d), dt)
    n_all = np.zeros(d, dt)
    for j in range(d):
        keep = (1 - mask[:, j]).astype(dt)
        xj = x * keep[:, None]
        n_j = keep.sum()
        x_bar = np.zeros((1, d), dt) if np.isclose(n_j, 0) else (xj.sum(axis=0, keepdims=True) / n_j)
        xc = (xj - x_bar) * keep[:, None]
        s_n = xc.T @ xc
        r_all[j] = small_t * np.eye(d, dtype=dt) + s_n + dt((n_j * alpha_mu) / (n_j + alpha_mu)) * \
            ((x_bar - mean_obs).T @ (x_bar - mean_obs))
        n_all[j] = n_j
    return r_all, n_all, small_t, dt(alpha_mu), dt(alpha_lambd)


def bge_logmarginal(g, x, mask, lik, dt, pre=None):
    """BGe.interventional_log_marginal_prob (linearGaussian.py:63-170) for hard g[S,d,d] (LU-form like the reference)."""
    s, d, _ = g.shape
    r_all, n_all, small_t, alpha_mu, alpha_lambd = pre if pre is not None else bge_precompute(x, mask, lik, dt)
    g = g.astype(dt)
    n_par = g.sum(axis=1)                               # [S,d] column sums
    eye = np.eye(d, dtype=dt)
    total = np.zeros(s, dt)
    for j in range(d):
        n_j = n_all[j]
        if np.isclose(n_j, 0):
            continue
        par = g[:, :, j]                                # [S,d]
        par_j = par + eye[:, j][None]
        l = n_par[:, j]
        log_gamma = (dt(0.5) * (np.log(alpha_mu) - np.log(n_j + alpha_mu))
                     + gammaln(dt(0.5) * (n_j + alpha_lambd - d + l + 1)).astype(dt)
                     - gammaln(dt(0.5) * (alpha_lambd - d + l + 1)).astype(dt)
                     - dt(0.5) * n_j * np.log(dt(np.pi))
                     + dt(0.5) * (alpha_lambd - d + 2 * l + 1) * np.log(small_t))

        def sub_logdet(pm):
            m = pm[:, :, None] * pm[:, None, :]
            sub = m * r_all[j][None] + (1 - m) * eye[None]
            return np.linalg.slogdet(sub)[1].astype(dt)

        log_r = (dt(0.5) * (n_j + alpha_lambd - d + l) * sub_logdet(par)
                 - dt(0.5) * (n_j + alpha_lambd - d + l + 1) * sub_logdet(par_j))
        total = total + (log_gamma + log_r).astype(dt)
    return total.astype(dt)


def bge_logmarginal_grads(g, x, mask, lik, dt, pre=None):
    """BGe score and its gradient w.r.t. a SOFT graph g[S,d,d] (the 'reparam' estimator with BGe, README.md:88-90;
    linearGaussian.py:82-118 with real-valued ``n_parents``; _slogdet_jax, utils/func.py:128-145).

    Per node j with p = g[:, j] (p_j = 0), l = sum p, A = D R D + I - D^2 (D = diag p), u_a = p_a R_aj:
      logdet B = logdet A + log s,  s = R_jj - u^T A^-1 u        (B = the same with q = p + e_j)
      d logdet A / dp_i = 2 sum_b (A^-1)_ib p_b (R_ib - delta_ib)
      d s / dp_i        = -2 R_ij w_i + 2 w_i sum_b p_b (R_ib - delta_ib) w_b,   w = A^-1 u
    and the log-gamma terms differentiate through l with digamma.  Returns (lp[S], dlp/dg[S,d,d]).
    """
    from scipy.special import digamma
    s_n, d, _ = g.shape
    r_all, n_all, small_t, alpha_mu, alpha_lambd = pre if pre is not None else bge_precompute(x, mask, lik, dt)
    g = g.astype(dt)
    eye = np.eye(d, dtype=dt)
    lp = np.zeros(s_n, dt)
    dg = np.zeros((s_n, d, d), dt)
    for j in range(d):
        n_j = n_all[j]
        if np.isclose(n_j, 0):
            continue
        r = r_all[j].astype(dt)
        for si in range(s_n):
            p = g[si, :, j].copy()
            p[j] = 0
            l = p.sum()
            a = (p[:, None] * p[None, :]) * r + (1 - p[:, None] * p[None, :]) * eye
            ainv = np.linalg.inv(a)
            la = np.linalg.slogdet(a)[1]
            u = p * r[:, j]
            w = ainv @ u
            sch = r[j, j] - u @ w
            lb = la + np.log(sch)
            c1 = dt(0.5) * (n_j + alpha_lambd - d + l)
            c2 = dt(0.5) * (n_j + alpha_lambd - d + l + 1)
            log_gamma = (dt(0.5) * (np.log(alpha_mu) - np.log(n_j + alpha_mu))
                         + gammaln(dt(0.5) * (n_j + alpha_lambd - d + l + 1)) - gammaln(dt(0.5) * (alpha_lambd - d + l + 1))
                         - dt(0.5) * n_j * np.log(dt(np.pi)) + dt(0.5) * (alpha_lambd - d + 2 * l + 1) * np.log(small_t))
            lp[si] += log_gamma + c1 * la - c2 * lb
            rm = r - eye
            dla = 2 * ((ainv * rm) @ p)
            t = rm @ (p * w)
            dsch = -2 * r[:, j] * w + 2 * w * t
            dlb = dla + dsch / sch
            dlg = (dt(0.5) * digamma(dt(0.5) * (n_j + alpha_lambd - d + l + 1)) - dt(0.5) * digamma(dt(0.5) * (alpha_lambd - d + l + 1))
                   + np.log(small_t))
            col = dlg + dt(0.5) * la - dt(0.5) * lb + c1 * dla - c2 * dlb
            col[j] = 0
            dg[si, :, j] = col
    return lp.astype(dt), dg.astype(dt)


def log_joint(cfg, g, theta, x, mask, dt, want_grads=True, pre=None):
    k = cfg.lik.kind
    if k == "lingauss":
        return lingauss_logjoint(g, theta, x, mask, cfg.lik, dt, want_grads)
    if k == "densenn":
        return densenn_logjoint(g, theta, x, mask, cfg.lik, dt, want_grads)
    if k == "bge":
        if want_grads and cfg.grad_estimator_z == "reparam":
            lp, dg = bge_logmarginal_grads(g, x, mask, cfg.lik, dt, pre)
            return lp, dg, None
        return bge_logmarginal(g, x, mask, cfg.lik, dt, pre), None, None
    raise NotImplementedError(k)


# --------------------------------------------------------------------------------------
# gradient estimators (one particle; batched over MC samples)
# --------------------------------------------------------------------------------------


def _offdiag(d, dt):
    return (1 - np.eye(d)).astype(dt)


def grad_z_score_function(cfg, z, theta, baseline, t, subk, x, mask, dt, pre=None):
    """DiBS.grad_z_likelihood_score_function (dibs.py:325-391); returns (grad[d,k,2], new_baseline, lp[S], G[S,d,d])."""
    d, k = z.shape[0:2]
    n_mc = cfg.n_grad_mc_samples
    alpha = alpha_of(cfg, t, dt)
    p = edge_probs(z, alpha)
    sk = tf.split(subk, 2, cfg.partitionable)           # subk, subk_ = split(subk)
    g = sample_g(p, sk[1], n_mc, cfg.partitionable)
    lp, _, _ = log_joint(cfg, g, theta, x, mask, dt, want_grads=False, pre=pre)
    lp_adj = lp if cfg.score_function_baseline <= 0.0 else (lp - dt(baseline)).astype(dt)
    # closed form of grad_Z log p(G|Z) (App. B-1): dS = alpha (G - P) o offdiag
    ds = alpha * (g.astype(dt) - p[None]) * _offdiag(d, dt)[None]
    gz = ds_to_dz(ds, z[None]).reshape(n_mc, d * k * 2).T       # [2dk, S]
    log_num, sign = logsumexp(lp_adj[None, :], axis=1, b=gz, return_sign=True)
    log_den = logsumexp(lp, axis=0)
    grad = sign * np.exp(log_num - np.log(dt(n_mc)) - log_den + np.log(dt(n_mc)))
    new_b = dt(cfg.score_function_baseline) * lp.mean() + dt(1 - cfg.score_function_baseline) * dt(baseline)
    return grad.reshape(d, k, 2).astype(dt), dt(new_b), lp, g


def grad_z_reparam(cfg, z, theta, baseline, t, subk, x, mask, dt):
    """DiBS.grad_z_likelihood_gumbel (dibs.py:395-459); returns (grad[d,k,2], baseline, lp[S], soft G[S,d,d])."""
    d, k = z.shape[0:2]
    n_mc = cfg.n_grad_mc_samples
    alpha = alpha_of(cfg, t, dt)
    sk = tf.split(subk, 2, cfg.partitionable)
    eps = tf.logistic(sk[1], (n_mc, d, d), cfg.partitionable)
    g = soft_graph(z[None], eps, alpha, cfg.tau)
    lp, dg, _ = log_joint(cfg, g, theta, x, mask, dt, want_grads=True)
    # App. B-4: dS = dlp/dG o tau*alpha*g(1-g) o offdiag
    ds = dg * (dt(cfg.tau) * alpha) * g * (1 - g) * _offdiag(d, dt)[None]
    gz = ds_to_dz(ds, z[None])                                    # [S,d,k,2]
    log_num, sign = logsumexp(lp[:, None, None, None], axis=0, b=gz, return_sign=True)
    log_den = logsumexp(lp, axis=0)
    grad = sign * np.exp(log_num - np.log(dt(n_mc)) - log_den + np.log(dt(n_mc)))
    return grad.astype(dt), dt(baseline), lp, g


def grad_theta(cfg, z, theta, t, subk, x, mask, dt):
    """DiBS.grad_theta_likelihood (dibs.py:488-551); returns (grad_theta, lp[S], G[S,d,d])."""
    n_mc = cfg.n_grad_mc_samples
    alpha = alpha_of(cfg, t, dt)
    p = edge_probs(z, alpha)
    g = sample_g(p, subk, n_mc, cfg.partitionable)               # key used directly (dibs.py:510)
    lp, _, dth = log_joint(cfg, g, theta, x, mask, dt, want_grads=True)
    lead = (slice(None),) + (None,) * (dth.ndim - 1)
    log_num, sign = logsumexp(lp[lead], axis=0, b=dth, return_sign=True)
    log_den = logsumexp(lp, axis=0)
    grad = sign * np.exp(log_num - np.log(dt(n_mc)) - log_den + np.log(dt(n_mc)))
    return grad.astype(dt), lp, g


def acyclic_constr(g, dt):  # graph_utils.py:8-28
    d = g.shape[-1]
    m = np.eye(d, dtype=dt) + dt(1.0 / d) * g.astype(dt)
    return np.trace(matrix_power(m, d), axis1=-2, axis2=-1) - dt(d)


def grad_constraint_gumbel(cfg, z, key, t, dt):
    """DiBS.grad_constraint_gumbel (dibs.py:576-601): mean_a grad_Z h(soft_G(Z, eps_a))."""
    d = z.shape[0]
    n_mc = cfg.n_acyclicity_mc_samples
    alpha = alpha_of(cfg, t, dt)
    eps = tf.logistic(key, (n_mc, d, d), cfg.partitionable)       # key used directly (dibs.py:595)
    g = soft_graph(z[None], eps, alpha, cfg.tau)
    m = np.eye(d, dtype=dt)[None] + dt(1.0 / d) * g
    dh_dg = np.swapaxes(matrix_power(m, d - 1), 1, 2)             # App. B-3
    ds = dh_dg * (dt(cfg.tau) * alpha) * g * (1 - g) * _offdiag(d, dt)[None]
    return ds_to_dz(ds, z[None]).mean(axis=0).astype(dt)


def grad_graph_prior(cfg, z, t, dt):
    """grad_Z of DiBS.log_graph_prior_particle (dibs.py:604-623) through P = edge_probs (App. B-5)."""
    d = z.shape[0]
    alpha = alpha_of(cfg, t, dt)
    p = edge_probs(z, alpha)
    kind = cfg.prior.kind
    if kind == "er":
        pr = dt(cfg.prior.p)
        with np.errstate(divide="ignore", invalid="ignore"):
            coef = np.log(pr) - np.log(dt(1) - pr)                # NaN/inf for d <= 5 like the reference (Q10)
        dp = np.full((d, d), coef, dt)
    elif kind == "sf":
        indeg = p.sum(axis=0)
        dp = np.broadcast_to((dt(-3.0) / (dt(1) + indeg))[None, :], (d, d)).astype(dt)
    elif kind == "uniform":
        dp = np.zeros((d, d), dt)
    else:
        raise NotImplementedError(kind)
    ds = dp * alpha * p * (1 - p) * _offdiag(d, dt)
    return ds_to_dz(ds, z).astype(dt)


def grad_latent_prior(cfg, z, key, t, latent_prior_std, dt):
    """DiBS.eltwise_grad_latent_prior for one particle (dibs.py:626-658)."""
    std = dt(latent_prior_std)
    return (- beta_of(cfg, t, dt) * grad_constraint_gumbel(cfg, z, key, t, dt)
            - z / dt(std ** dt(2.0))
            + grad_graph_prior(cfg, z, t, dt)).astype(dt)


# --------------------------------------------------------------------------------------
# kernel, phi, optimizer
# --------------------------------------------------------------------------------------


def kernel_matrix(cfg, z, theta, dt):
    """_f_kernel_mat (svgd.py:165-176,537-551) with kernel.py:30,66-71.  z[M,..], theta[M,Dtheta] or None."""
    m = z.shape[0]
    zf = z.reshape(m, -1).astype(dt)
    dz = ((zf[:, None, :] - zf[None, :, :]) ** 2).sum(-1)
    kz = dt(cfg.scale_latent) * np.exp(-dz / dt(cfg.h_latent))
    if theta is None:
        return kz.astype(dt), kz.astype(dt), None
    tf_ = theta.reshape(m, -1).astype(dt)
    dth = ((tf_[:, None, :] - tf_[None, :, :]) ** 2).sum(-1)
    kt = dt(cfg.scale_theta) * np.exp(-dth / dt(cfg.h_theta))
    return (kz + kt).astype(dt), kz.astype(dt), kt.astype(dt)


def phi_update(k_full, k_term, h, xs, grads, dt):
    """_z_update/_theta_update (svgd.py:194-224,591-670) via App. B-9.

    phi_i = -(1/M) sum_j [ K_ji grad_j + (-(2/h)) (x_j - x_i) Kterm_ji ].
    """
    m = xs.shape[0]
    xf = xs.reshape(m, -1).astype(dt)
    gf = grads.reshape(m, -1).astype(dt)
    drive = k_full.T @ gf
    rep = -(dt(2.0) / dt(h)) * (k_term.T @ xf - k_term.sum(axis=0)[:, None] * xf)
    return (-(drive + rep) / dt(m)).reshape(xs.shape).astype(dt)


def opt_update(cfg, x, v, phi, dt):
    """jax.example_libraries.optimizers sgd / rmsprop(gamma=0.9, eps=1e-8) (svgd.py:117-120,265)."""
    eta = dt(cfg.stepsize)
    if cfg.optimizer == "gd":
        return (x - eta * phi).astype(dt), v
    if cfg.optimizer == "rmsprop":
        v = (v * dt(0.9) + (phi * phi) * dt(1.0 - 0.9)).astype(dt)
        return (x - eta * phi / np.sqrt(v + dt(1e-8))).astype(dt), v
    raise ValueError()


# --------------------------------------------------------------------------------------
# full steps
# --------------------------------------------------------------------------------------


@dataclass
class State:
    z: np.ndarray
    v_z: np.ndarray
    key: np.ndarray
    sf_baseline: np.ndarray
    theta: Optional[np.ndarray] = None      # [M, Dtheta] flat
    v_theta: Optional[np.ndarray] = None
    latent_prior_std: float = 0.0
    extras: dict = field(default_factory=dict)


def theta_for_model(cfg, theta_row):
    if cfg.lik.kind == "lingauss":
        d = cfg.lik.n_vars
        return theta_row.reshape(d, d)
    return theta_row


def particle_grads(cfg, st, t, x, mask, dt, particles=None, keep=False):
    """Gradient phase of _svgd_step for particles in ``particles`` (default all): svgd.py:245-255 / 695-707.

    Returns (dz_log_prob[M,d,k,2], dtheta[M,Dtheta] or None, new_key, new_baseline[M]) with rows outside
    ``particles`` left zero -- this is what a rank computes for its shard.
    """
    m = st.z.shape[0]
    part = slice(None)
    idx = range(m) if particles is None else particles
    key = st.key
    dz = np.zeros(st.z.shape, dt)
    dth = None if st.theta is None else np.zeros(st.theta.shape, dt)
    base = st.sf_baseline.astype(dt).copy()
    pre = bge_precompute(x, mask, cfg.lik, dt) if cfg.lik.kind == "bge" else None
    extras = {}
    if cfg.joint:
        ks = tf.split(key, m + 1, cfg.partitionable); key = ks[0]
        for i in idx:
            g_th, lp, _ = grad_theta(cfg, st.z[i].astype(dt), theta_for_model(cfg, st.theta[i]), t, ks[i + 1], x, mask, dt)
            dth[i] = g_th.reshape(-1)
            if keep:
                extras.setdefault("lp_theta", {})[i] = lp
    ks = tf.split(key, m + 1, cfg.partitionable); key = ks[0]
    for i in idx:
        th = None if st.theta is None else theta_for_model(cfg, st.theta[i])
        if cfg.grad_estimator_z == "score":
            gz, nb, lp, _ = grad_z_score_function(cfg, st.z[i].astype(dt), th, base[i], t, ks[i + 1], x, mask, dt, pre)
        elif cfg.grad_estimator_z == "reparam":
            gz, nb, lp, _ = grad_z_reparam(cfg, st.z[i].astype(dt), th, base[i], t, ks[i + 1], x, mask, dt)
        else:
            raise ValueError(f"Unknown gradient estimator `{cfg.grad_estimator_z}`")
        dz[i] = gz
        base[i] = nb
        if keep:
            extras.setdefault("lp_z", {})[i] = lp
            extras.setdefault("dz_lik", {})[i] = gz
    ks = tf.split(key, m + 1, cfg.partitionable); key = ks[0]
    for i in idx:
        gp = grad_latent_prior(cfg, st.z[i].astype(dt), ks[i + 1], t, st.latent_prior_std, dt)
        if keep:
            extras.setdefault("dz_prior", {})[i] = gp
        dz[i] = gp + dz[i]
    return dz, dth, key, base, extras


def svgd_step(cfg, st, t, x, mask, dt, keep=False):
    """One full _svgd_step (svgd.py:226-267 marginal, 673-721 joint)."""
    dz, dth, key, base, extras = particle_grads(cfg, st, t, x, mask, dt, keep=keep)
    k_full, k_z, k_t = kernel_matrix(cfg, st.z, st.theta if cfg.joint else None, dt)
    phi_z = phi_update(k_full, k_z, cfg.h_latent, st.z, dz, dt)
    z_new, vz_new = opt_update(cfg, st.z.astype(dt), st.v_z.astype(dt), phi_z, dt)
    out = State(z=z_new, v_z=vz_new, key=key, sf_baseline=base, latent_prior_std=st.latent_prior_std)
    if cfg.joint:
        phi_t = phi_update(k_full, k_t, cfg.h_theta, st.theta, dth, dt)
        out.theta, out.v_theta = opt_update(cfg, st.theta.astype(dt), st.v_theta.astype(dt), phi_t, dt)
    if keep:
        extras.update(dz=dz, dth=dth, k=k_full, phi_z=phi_z)
        if cfg.joint:
            extras.update(phi_theta=phi_t)
        out.extras = extras
    return out


def init_particles(cfg, key, n_particles, n_dim, dt):
    """sample(): key split + _sample_initial_random_particles (svgd.py:125-148,294-307 / 489-515,751-766)."""
    d = cfg.lik.n_vars
    k = n_dim or d
    ks = tf.split(key, 2, cfg.partitionable); key, subk = ks[0], ks[1]
    std = np.float32(cfg.latent_prior_std) if cfg.latent_prior_std else np.float32(1.0) / np.sqrt(np.float32(k))
    kk = tf.split(subk, 2, cfg.partitionable); k2, sub2 = kk[0], kk[1]
    z = (tf.normal(sub2, (n_particles, d, k, 2), cfg.partitionable) * std).astype(np.float32)
    theta = None
    if cfg.joint:
        kk = tf.split(k2, 2, cfg.partitionable); sub3 = kk[1]
        theta = sample_parameters(cfg, sub3, n_particles)
    st = State(z=z.astype(dt), v_z=np.zeros_like(z, dt), key=key, sf_baseline=np.zeros(n_particles, dt),
               theta=None if theta is None else theta.astype(dt),
               v_theta=None if theta is None else np.zeros_like(theta, dt),
               latent_prior_std=float(std))
    return st


def sample_parameters(cfg, key, n_particles):
    """likelihood_model.sample_parameters(key=, n_particles=, n_vars=) -> flat [M, Dtheta] fp32."""
    lik, d = cfg.lik, cfg.lik.n_vars
    if lik.kind == "lingauss":  # linearGaussian.py:212-227
        th = np.float32(lik.mean_edge) + np.float32(lik.sig_edge) * tf.normal(key, (n_particles, d, d), cfg.partitionable)
        th = th + np.sign(th) * np.float32(lik.min_edge)
        return th.reshape(n_particles, d * d).astype(np.float32)
    if lik.kind == "densenn":   # nonlinearGaussian.py:155-186 + stax.serial/Dense init (App. A.4)
        h = lik.hidden
        subkeys = tf.split(key, n_particles * d, cfg.partitionable).reshape(n_particles, d, 2)
        w1 = np.zeros((n_particles, d, d, h), np.float32); b1 = np.zeros((n_particles, d, h), np.float32)
        w2 = np.zeros((n_particles, d, h), np.float32); b2 = np.zeros((n_particles, d), np.float32)
        sp = np.float32(lik.sig_param)
        for m in range(n_particles):
            for j in range(d):
                rng = subkeys[m, j]
                rng, l0 = tf.split(rng, 2, cfg.partitionable)         # Dense(H)
                k1, k2 = tf.split(l0, 2, cfg.partitionable)
                w1[m, j] = tf.normal(k1, (d, h), cfg.partitionable) * sp
                b1[m, j] = tf.normal(k2, (h,), cfg.partitionable) * sp
                rng, _ = tf.split(rng, 2, cfg.partitionable)          # activation layer consumes a split
                rng, l2 = tf.split(rng, 2, cfg.partitionable)         # Dense(1)
                k1, k2 = tf.split(l2, 2, cfg.partitionable)
                w2[m, j] = tf.normal(k1, (h, 1), cfg.partitionable)[:, 0] * sp
                b2[m, j] = tf.normal(k2, (1,), cfg.partitionable)[0] * sp
        return nn_pack(w1, b1, w2, b2).astype(np.float32)
    raise NotImplementedError("Not available for BGe score; use `LinearGaussian` model instead.")


def sample(cfg, key, n_particles, steps, x, mask=None, n_dim=None, callback_every=None, dt=np.float32):
    """MarginalDiBS.sample / JointDiBS.sample (svgd.py:274-331,730-795) -> (G int32[M,d,d], theta, final State)."""
    x = np.asarray(x)
    mask = np.zeros(x.shape, np.int32) if mask is None else np.asarray(mask)
    st = init_particles(cfg, np.asarray(key, np.uint32), n_particles, n_dim, dt)
    callback_every = callback_every or steps
    for t0 in (range(0, steps, callback_every) if steps else range(0)):
        for t in range(t0, t0 + callback_every):
            st = svgd_step(cfg, st, t, x, mask, dt)
    return particle_to_g_lim(st.z), st.theta, st
