"""TEST INFRASTRUCTURE ONLY -- minimal torch-CPU stand-in for the ``jax`` package.

Implements only what /root/reference/dibs uses (see ../README.md).  The goal is
semantic fidelity of the fp32 arithmetic and of the PRNG stream (threefry2x32,
non-partitionable layout), not performance or completeness.
"""
import torch

torch.set_default_dtype(torch.float32)

from . import _patch  # noqa: F401  (adds .at / .astype / tuple-transpose to torch.Tensor)
from ._transforms import vmap, grad, jit, device_get
from . import numpy, random, lax, tree_util, nn, scipy, example_libraries  # noqa: F401

__version__ = "0.0-shim"
