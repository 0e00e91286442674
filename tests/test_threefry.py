"""Pins for the JAX PRNG restatement (oracle/threefry.py): published known-answer vectors."""
import numpy as np

from oracle import threefry as tf
from helpers import load


def test_random123_kats():
    # Random123 kat_vectors, threefry2x32 20 rounds
    cases = [((0, 0), (0, 0), (0x6b200159, 0x99ba4efe)),
             ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
             ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]
    for key, ctr, want in cases:
        y0, y1 = tf.threefry2x32_block(key[0], key[1], [ctr[0]], [ctr[1]])
        assert (int(y0[0]), int(y1[0])) == want


def test_jax_documented_values():
    # jax.random.split(PRNGKey(0)) as printed in the JAX documentation
    assert tf.split(tf.prng_key(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    # jax.random.normal(PRNGKey(0), (10,)) from the JAX quick-start
    want = [-0.3721109, 0.26423115, -0.18252768, -0.7368197, -0.44030377,
            -0.1521442, -0.67135346, -0.5908641, 0.73168886, 0.5673026]
    np.testing.assert_allclose(tf.normal(tf.prng_key(0), (10,)), want, atol=2e-7)
    assert tf.prng_key(123).tolist() == [0, 123]
    # partitionable layout (default from JAX 0.5.0): different stream, offered as a switch
    assert tf.split(tf.prng_key(0), partitionable=True).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


def test_matches_shim_stream():
    g = load("prng")
    for seed in (0, 123, 2**33 + 7):
        k = tf.prng_key(seed)
        assert (k == g[f"key_{seed}"]).all()
        assert (tf.split(k, 5) == g[f"split5_{seed}"]).all()
        assert (tf.uniform(k, (3, 7)) == g[f"uniform_{seed}"]).all()
        np.testing.assert_allclose(tf.logistic(k, (2, 3, 3)), g[f"logistic_{seed}"], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(tf.normal(k, (11,)), g[f"normal_{seed}"], atol=5e-7)
        assert (tf.bernoulli(k, 0.3, (4, 5, 5)) == g[f"bernoulli_{seed}"]).all()


def test_odd_length_padding():
    k = tf.prng_key(7)
    a = tf.random_bits(k, (5,))
    b = tf.threefry_2x32(k, np.array([0, 1, 2, 3, 4, 0], np.uint32))[:5]
    assert (a == b).all()
