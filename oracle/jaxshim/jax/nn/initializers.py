"""TEST INFRASTRUCTURE ONLY -- ``jax.nn.initializers.normal``."""
import torch


def normal(stddev=1e-2, dtype=torch.float32):
    from .. import random

    def init(key, shape, dtype=dtype):
        return random.normal(key, shape) * stddev

    return init
