"""TEST INFRASTRUCTURE ONLY -- ``jax.example_libraries`` subset."""
from . import optimizers, stax  # noqa: F401
