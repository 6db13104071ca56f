"""``jax.scipy.linalg`` -> scipy.linalg (results wrapped as JArray)."""
import scipy.linalg as _sl

from .._array import wrap


def __getattr__(name):
    fn = getattr(_sl, name)
    if not callable(fn):
        return fn

    def call(*args, **kwargs):
        return wrap(fn(*args, **kwargs))

    call.__name__ = name
    return call
