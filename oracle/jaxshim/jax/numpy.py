"""``jax.numpy`` -> NumPy, returning JArray; Taylor-series arguments (see experimental/jet.py) are routed
coefficient-wise through the structural/linear functions."""

import sys
import types

import numpy as _np

from ._array import JArray, wrap

nan = _np.nan
inf = _np.inf
pi = _np.pi
ndarray = _np.ndarray
float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
newaxis = None


def _is_taylor(x):
    return type(x).__name__ == "TaylorArray"


def _any_taylor(obj):
    if _is_taylor(obj):
        return True
    if isinstance(obj, (list, tuple)):
        return any(_any_taylor(o) for o in obj)
    return False


# functions that act linearly / structurally on their first argument: apply per Taylor coefficient
_LINEAR = {"roll", "reshape", "ravel", "flip", "pad", "transpose", "convolve", "sum", "squeeze", "real"}


def _dispatch(name, fn):
    def wrapper(*args, **kwargs):
        if _any_taylor(args) or _any_taylor(list(kwargs.values())):
            from .experimental.jet import taylor_function

            return taylor_function(name, fn, args, kwargs)
        return wrap(fn(*args, **kwargs))

    wrapper.__name__ = name
    return wrapper


def isscalar(x):
    return _np.isscalar(x)


def array(obj, dtype=None, **kwargs):
    if _any_taylor(obj):
        from .experimental.jet import taylor_function

        return taylor_function("array", _np.array, (obj,), {})
    return wrap(_np.array(obj, dtype=dtype, **kwargs))


asarray = array


class _LinalgProxy(types.ModuleType):
    def __getattr__(self, name):
        return _dispatch("linalg." + name, getattr(_np.linalg, name))


linalg = _LinalgProxy("jax.numpy.linalg")
sys.modules["jax.numpy.linalg"] = linalg


def __getattr__(name):
    attr = getattr(_np, name)
    if callable(attr) and not isinstance(attr, type):
        return _dispatch(name, attr)
    return attr
