"""TEST INFRASTRUCTURE ONLY -- ``jax.scipy`` subset."""
from . import special, stats  # noqa: F401
