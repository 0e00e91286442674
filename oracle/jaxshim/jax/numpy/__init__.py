"""TEST INFRASTRUCTURE ONLY -- the ``jax.numpy`` names used by /root/reference/dibs, on torch CPU."""
import math

import numpy as _onp
import torch

from . import linalg  # noqa: F401

pi = math.pi
newaxis = None
float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
uint32 = torch.int64  # keys are carried in int64 lanes (values < 2**32)
bool_ = torch.bool
ndarray = torch.Tensor


def _t(x):
    """Python scalars become weak-typed fp32 / int tensors like in JAX (x64 disabled)."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, bool):
        return torch.tensor(x)
    if isinstance(x, int):
        return torch.tensor(x, dtype=torch.int32)
    if isinstance(x, float):
        return torch.tensor(x, dtype=torch.float32)
    if isinstance(x, _onp.ndarray):
        t = torch.from_numpy(_onp.ascontiguousarray(x))
        return t.to(torch.float32) if t.dtype == torch.float64 else t
    return array(x)


def _f(x):
    t = _t(x)
    return t if t.is_floating_point() else t.to(torch.float32)


def array(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        out = torch.stack([_t(v) for v in x])
    elif isinstance(x, _onp.ndarray):
        out = _t(x)
    else:
        out = torch.tensor(x)
        if out.dtype == torch.float64:
            out = out.to(torch.float32)
        elif out.dtype == torch.int64:
            out = out.to(torch.int32)
    return out if dtype is None else out.to(dtype)


asarray = array


def _shape(shape):
    if isinstance(shape, (int, torch.Tensor)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def zeros(shape, dtype=None):
    return torch.zeros(_shape(shape), dtype=dtype or torch.float32)


def ones(shape, dtype=None):
    return torch.ones(_shape(shape), dtype=dtype or torch.float32)


def zeros_like(x, dtype=None):
    return torch.zeros_like(_t(x), dtype=dtype)


def ones_like(x, dtype=None):
    return torch.ones_like(_t(x), dtype=dtype)


def eye(n, dtype=None):
    return torch.eye(int(n), dtype=dtype or torch.float32)


def arange(*a, dtype=None):
    return torch.arange(*[int(v) for v in a], dtype=dtype or torch.int64)


def log(x):
    return torch.log(_f(x))


def exp(x):
    return torch.exp(_f(x))


def sqrt(x):
    return torch.sqrt(_f(x))


def square(x):
    x = _t(x)
    return x * x


def abs(x):  # noqa: A001
    return torch.abs(_t(x))


def sign(x):
    return torch.sign(_t(x))


def ceil(x):
    return torch.ceil(_f(x))


def add(a, b):
    return _t(a) + _t(b)


def subtract(a, b):
    return _t(a) - _t(b)


def _axis_kw(axis, keepdims):
    kw = {}
    if axis is not None:
        kw["dim"] = axis
    if keepdims:
        kw["keepdim"] = True
    return kw


def sum(x, axis=None, keepdims=False):  # noqa: A001
    return torch.sum(_t(x), **_axis_kw(axis, keepdims))


def trace(x):
    return torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)


def where(cond, a, b):
    cond = _t(cond)
    if cond.dtype != torch.bool:
        cond = cond != 0
    a_t, b_t = isinstance(a, torch.Tensor), isinstance(b, torch.Tensor)
    if not a_t and not b_t:
        a = _t(a)
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.dtype != b.dtype:
        dt = torch.promote_types(a.dtype, b.dtype)
        a, b = a.to(dt), b.to(dt)
    return torch.where(cond, a, b)


def isclose(a, b, rtol=1e-05, atol=1e-08):
    a, b = _f(a), _f(b)
    return torch.abs(a - b) <= (atol + rtol * torch.abs(b))


def einsum(spec, *ops):
    ops = [_t(o) for o in ops]
    dt = ops[0].dtype
    for o in ops[1:]:
        dt = torch.promote_types(dt, o.dtype)
    return torch.einsum(spec, *[o.to(dt) for o in ops])


def expand_dims(x, axis):
    if isinstance(axis, int):
        axis = (axis,)
    for ax in sorted(axis):
        x = x.unsqueeze(ax)
    return x


def dot(a, b):
    return a @ b


def stack(xs, axis=0):
    return torch.stack([_t(v) for v in xs], dim=axis)


def concatenate(xs, axis=0):
    return torch.cat([_t(v) for v in xs], dim=axis)


def tril(x, k=0):
    return torch.tril(x, diagonal=k)


def sort(x, axis=-1):
    return torch.sort(x, dim=axis).values


def maximum(a, b):
    return torch.maximum(_t(a), _t(b))


def finfo(dtype):
    return torch.finfo(dtype)
