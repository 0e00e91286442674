"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the JAX PRNG (threefry2x32).

The reference draws every random number through ``jax.random`` (call sites:
dibs/inference/svgd.py:145-146,245,251,294,509-513,695-703,751;
dibs/inference/dibs.py:115,350,358,430-431,436,517,595;
dibs/models/linearGaussian.py:225; dibs/models/nonlinearGaussian.py:169).
JAX is a third-party dependency that is NOT vendored under /root/reference and
is only lower-bounded (requirements.txt: ``jax>=0.3.17``), so its published
algorithm is restated here:

* block function: Threefry-2x32, 20 rounds (Salmon et al., "Parallel random
  numbers: as easy as 1, 2, 3", SC'11; Random123 ``threefry2x32_20``);
* ``jax._src.prng.threefry_2x32(key, count)``: flatten ``count``, pad to even
  length, FIRST half -> lane 0, SECOND half -> lane 1, concatenate outputs;
* ``split`` / ``random_bits`` / ``uniform`` / ``bernoulli`` / ``logistic`` /
  ``normal`` as in ``jax._src.random`` with ``jax_threefry_partitionable=False``
  (the default for every JAX release < 0.5.0, i.e. the releases contemporary
  with the reference); ``partitionable=True`` selects the >= 0.5.0 layout.

Pinned by known-answer tests in tests/test_threefry.py.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & _M32


def threefry2x32_block(k0, k1, x0, x1):
    """Threefry-2x32-20 on uint32 arrays (computed in uint64 lanes, masked)."""
    k0 = np.asarray(k0, dtype=np.uint64) & _M32
    k1 = np.asarray(k1, dtype=np.uint64) & _M32
    x0 = np.asarray(x0, dtype=np.uint64) & _M32
    x1 = np.asarray(x1, dtype=np.uint64) & _M32
    ks = (k0, k1, (k0 ^ k1 ^ np.uint64(0x1BD11BDA)) & _M32)
    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r)
            x1 = x1 ^ x0
        x0 = (x0 + ks[(g + 1) % 3]) & _M32
        x1 = (x1 + ks[(g + 2) % 3] + np.uint64(g + 1)) & _M32
    return x0.astype(np.uint32), x1.astype(np.uint32)


def threefry_2x32(key, count):
    """``jax._src.prng.threefry_2x32``: hash a flat uint32 counter array."""
    key = np.asarray(key, dtype=np.uint32)
    count = np.asarray(count, dtype=np.uint32).ravel()
    n = count.size
    if n % 2:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    h = count.size // 2
    y0, y1 = threefry2x32_block(key[0], key[1], count[:h], count[h:])
    return np.concatenate([y0, y1])[:n]


def prng_key(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num=2, partitionable=False):
    if partitionable:
        hi, lo = _bits64_iota(key, num)
        return np.stack([hi, lo], axis=1)
    return threefry_2x32(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def _bits64_iota(key, n):
    key = np.asarray(key, dtype=np.uint32)
    idx = np.arange(n, dtype=np.uint64)
    return threefry2x32_block(key[0], key[1], idx >> np.uint64(32), idx & _M32)


def random_bits(key, shape, partitionable=False):
    n = int(np.prod(shape)) if len(shape) else 1
    if partitionable:
        hi, lo = _bits64_iota(key, n)
        return (hi ^ lo).reshape(shape)
    return threefry_2x32(key, np.arange(n, dtype=np.uint32)).reshape(shape)


def bits_to_unit_float(bits):
    """``bitcast((bits >> 9) | 0x3F800000) - 1.0`` -- exact in fp32."""
    m = (np.asarray(bits, dtype=np.uint32) >> np.uint32(9)).astype(np.float32)
    return m * np.float32(2.0 ** -23)


def uniform(key, shape, minval=0.0, maxval=1.0, partitionable=False):
    f = bits_to_unit_float(random_bits(key, shape, partitionable))
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    return np.maximum(minval, (f * np.float32(maxval - minval) + minval).astype(np.float32))


def bernoulli(key, p, shape, partitionable=False):
    return uniform(key, shape, partitionable=partitionable) < np.asarray(p, dtype=np.float32)


def logistic(key, shape, partitionable=False):
    u = uniform(key, shape, minval=np.finfo(np.float32).eps, maxval=1.0, partitionable=partitionable)
    return (np.log(u) - np.log1p(-u)).astype(np.float32)


def normal(key, shape, partitionable=False):
    from scipy.special import erfinv
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, minval=lo, maxval=1.0, partitionable=partitionable)
    return (np.float32(np.sqrt(2.0)) * erfinv(u.astype(np.float64)).astype(np.float32)).astype(np.float32)
