"""TEST INFRASTRUCTURE ONLY -- ``jax.random`` (threefry2x32, non-partitionable layout) in torch int64 lanes.

Same algorithm as oracle/threefry.py (which carries the citations and the KATs) but written with
torch ops so that it runs under ``torch.func.vmap`` with batched keys, as the reference does
(`random.split(subk)` inside vmapped per-particle functions, dibs/inference/dibs.py:350,430,517).
"""
import math

import torch

_M32 = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M32


def _block(k0, k1, x0, x1):
    ks = (k0, k1, (k0 ^ k1 ^ 0x1BD11BDA) & _M32)
    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r)
            x1 = x1 ^ x0
        x0 = (x0 + ks[(g + 1) % 3]) & _M32
        x1 = (x1 + ks[(g + 2) % 3] + (g + 1)) & _M32
    return x0, x1


def _threefry_2x32(key, n):
    """Hash counters ``arange(n)`` with ``key`` -> int64 tensor [n] of uint32 values."""
    key = key.to(torch.int64)
    m = n + (n % 2)
    h = m // 2
    cnt = torch.arange(m, dtype=torch.int64)
    if n % 2:
        cnt[-1] = 0
    y0, y1 = _block(key[0], key[1], cnt[:h], cnt[h:])
    return torch.cat([y0, y1])[:n]


def PRNGKey(seed):
    seed = int(seed)
    return torch.tensor([(seed >> 32) & _M32, seed & _M32], dtype=torch.int64)


def split(key, num=2):
    num = int(num)
    return _threefry_2x32(key, 2 * num).reshape(num, 2)


def _shape(shape):
    if isinstance(shape, (int, torch.Tensor)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def _bits(key, shape):
    shape = _shape(shape)
    return _threefry_2x32(key, int(math.prod(shape))).reshape(shape)


def uniform(key, shape=(), dtype=torch.float32, minval=0.0, maxval=1.0):
    f = (_bits(key, shape) >> 9).to(torch.float32) * (2.0 ** -23)  # == bitcast(bits>>9 | 0x3F800000) - 1
    minval = torch.tensor(minval, dtype=torch.float32)
    maxval = torch.tensor(maxval, dtype=torch.float32)
    return torch.maximum(minval, f * (maxval - minval) + minval)


def bernoulli(key, p=0.5, shape=None):
    p = p if isinstance(p, torch.Tensor) else torch.tensor(p, dtype=torch.float32)
    if shape is None:
        shape = p.shape
    return uniform(key, shape) < p


def logistic(key, shape=(), dtype=torch.float32):
    u = uniform(key, shape, minval=torch.finfo(torch.float32).eps, maxval=1.0)
    return torch.log(u) - torch.log1p(-u)


def normal(key, shape=(), dtype=torch.float32):
    lo = float(torch.nextafter(torch.tensor(-1.0), torch.tensor(0.0)))
    u = uniform(key, shape, minval=lo, maxval=1.0)
    return math.sqrt(2.0) * torch.erfinv(u)
