"""TEST INFRASTRUCTURE ONLY: a NumPy-backed stand-in for the handful of ``jax`` entry points that
pnkraemer/tornadox imports, so that the UNMODIFIED reference sources can be executed where JAX is not
installable (no network, no wheel).  It is NOT JAX: there is no tracing, no XLA; every ``jnp`` call is the
NumPy function of the same name (LAPACK-backed ``linalg``), ``jit`` is the identity, ``vmap`` is a Python
loop, ``jacfwd`` is complex-step differentiation, ``jet`` is a truncated Taylor-series algebra.
Arithmetic is IEEE fp64 like ``jax_enable_x64``; results agree with real JAX-CPU up to the reordering/FMA
freedom XLA takes (~1e-16 relative per operation).

Used only by oracle/make_golden.py and tests/ (put ``oracle/jaxshim`` on sys.path, then ``import tornadox``
from /root/reference).
"""

import contextlib
import types

import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, random, scipy  # noqa: F401
from ._array import JArray, wrap as _wrap
from .numpy import _is_taylor


class Array:
    """Placeholder so that third-party ``isinstance(x, jax.Array)`` probes (SciPy's array-API helpers) say no."""


class _Config:
    def update(self, *args, **kwargs):
        return None

    jax_enable_x64 = True


config = types.ModuleType("jax.config")
config.config = _Config()
config.update = config.config.update
import sys as _sys

_sys.modules["jax.config"] = config


def jit(fun=None, **kwargs):
    """Identity (no tracing).  Supports ``@jax.jit`` and ``partial(jax.jit, static_argnums=...)``."""
    if fun is None:
        return lambda f: f
    return fun


@contextlib.contextmanager
def disable_jit():
    yield


def vmap(fun, in_axes=0, out_axes=0):
    """Python-loop batching of ``fun`` over ``in_axes``; outputs (arrays or tuples of arrays) are stacked."""

    def batched(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        size = None
        for a, ax in zip(args, axes):
            if ax is not None:
                size = _np.shape(a)[ax]
                break
        outs = []
        for i in range(size):
            sliced = [a if ax is None else _np.take(a, i, axis=ax) for a, ax in zip(args, axes)]
            outs.append(fun(*sliced))
        if isinstance(outs[0], tuple):
            return tuple(_wrap(_np.stack([o[k] for o in outs], axis=out_axes)) for k in range(len(outs[0])))
        return _wrap(_np.stack(outs, axis=out_axes))

    return batched


def jacfwd(fun, argnums=0):
    """Dense Jacobian by complex-step differentiation (exact to rounding for the analytic vector fields used)."""

    def jac(*args):
        x = _np.asarray(args[argnums], dtype=float)
        h = 1e-30
        cols = []
        for j in range(x.size):
            xp = x.astype(complex).reshape(-1)
            xp[j] += 1j * h
            a = list(args)
            a[argnums] = _wrap(xp.reshape(x.shape))
            cols.append(_np.imag(_np.asarray(fun(*a))).reshape(-1) / h)
        return _wrap(_np.stack(cols, axis=1))

    return jac


from . import experimental  # noqa: E402,F401
from . import ops  # noqa: E402,F401
