"""TEST INFRASTRUCTURE ONLY -- ``jax.scipy.special`` subset, following jax/_src/ops/special.py."""
import torch


def gammaln(x):
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(float(x), dtype=torch.float32)
    return torch.lgamma(x if x.is_floating_point() else x.to(torch.float32))


def logsumexp(a, axis=None, b=None, keepdims=False, return_sign=False):
    """amax (finite-guarded, stop-gradient) shift; signed when ``b`` is given."""
    if b is not None:
        a, b = torch.broadcast_tensors(a, b.to(a.dtype) if b.is_floating_point() else b.to(a.dtype))
        a = torch.where(b != 0, a, torch.full_like(a, -float("inf")))
    dims = tuple(range(a.dim())) if axis is None else ((axis,) if isinstance(axis, int) else tuple(axis))
    amax = torch.amax(a, dim=dims, keepdim=True).detach()
    amax = torch.where(torch.isfinite(amax), amax, torch.zeros_like(amax))
    ex = torch.exp(a - amax)
    if b is not None:
        ex = b * ex
    s = torch.sum(ex, dim=dims, keepdim=keepdims)
    amax_out = amax if keepdims else amax.reshape(s.shape)
    sign = torch.sign(s)
    if return_sign:
        return torch.log(torch.abs(s)) + amax_out, sign
    if b is not None:
        s = torch.where(sign < 0, torch.full_like(s, float("nan")), s)
    return torch.log(s) + amax_out
