"""tornadox_b200: the tornadox filter step (KroneckerEK0 / DiagonalEK1 ``attempt_step``) as hand-written
sm_100a CUDA kernels behind the reference's solver API.

    from tornadox_b200 import ek0, ek1, init, ivp, step

mirrors ``from tornadox import ek0, ek1, init, ivp, step`` (tornadox/__init__.py:5).  The host side is a
thin ctypes layer over the C ABI in include/tornadox_b200.h; there is no CPU fallback.
"""

from . import ek0, ek1, engine, init, ivp, iwp, odefilter, rv, sqrt, step  # noqa: F401

__version__ = "0.1.0"
