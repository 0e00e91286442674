"""TEST INFRASTRUCTURE ONLY -- ``jax.nn`` subset."""
import torch

from . import initializers  # noqa: F401


def sigmoid(x):
    return torch.sigmoid(x)


def log_sigmoid(x):
    # jax.nn.log_sigmoid(x) = -softplus(-x)
    return -torch.nn.functional.softplus(-x)


def relu(x):
    return torch.relu(x)
