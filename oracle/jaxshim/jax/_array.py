"""ndarray subclass with the ``x.at[idx].set(v)`` functional-update syntax of jax arrays."""

import numpy as np


class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtRef(self.arr, idx)


class _AtRef:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        if type(value).__name__ == "TaylorArray":   # promote a constant array that receives a Taylor series
            from .experimental.jet import lift

            return lift(np.asarray(self.arr), value.order).at[self.idx].set(value)
        out = np.array(self.arr, copy=True, dtype=np.result_type(self.arr, np.asarray(value)))
        out[self.idx] = value
        return out.view(JArray)

    def add(self, value):
        out = np.array(self.arr, copy=True)
        out[self.idx] += value
        return out.view(JArray)


class JArray(np.ndarray):
    @property
    def at(self):
        return _AtIndexer(self)


def wrap(x):
    if isinstance(x, JArray):
        return x
    if isinstance(x, np.ndarray):
        return x.view(JArray)
    if isinstance(x, tuple) and not hasattr(x, "_fields"):
        return tuple(wrap(v) for v in x)
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x
