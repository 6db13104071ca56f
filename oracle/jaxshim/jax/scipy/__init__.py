"""``jax.scipy`` -> SciPy."""
from . import linalg, signal  # noqa: F401
