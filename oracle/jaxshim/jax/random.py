"""``jax.random``: threefry streams cannot be reproduced without JAX.  ``uniform``/``normal`` draw from NumPy's
PCG64 seeded with the key's seed instead -- good enough for the reference's own (value-agnostic) tests; parity
fixtures always pass y0 explicitly."""

import numpy as _np

from ._array import wrap


def PRNGKey(seed):
    return ("shim-prng-key", int(seed))


def split(key, num=2):
    return [("shim-prng-key", key[1] * 7919 + i + 1) for i in range(num)]


def uniform(key, shape=(), minval=0.0, maxval=1.0, **kwargs):
    return wrap(_np.random.default_rng(key[1]).uniform(minval, maxval, size=shape))


def normal(key, shape=(), **kwargs):
    return wrap(_np.random.default_rng(key[1]).standard_normal(size=shape))
