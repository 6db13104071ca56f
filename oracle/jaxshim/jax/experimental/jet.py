"""``jax.experimental.jet.jet`` as a truncated Taylor-series algebra over NumPy (TEST INFRASTRUCTURE).

Convention (same as JAX): for the path x(t) = x0 + v1 t + v2 t^2/2! + ... the call
``jet(fun, (x0,), ((v1, ..., vK),))`` returns ``(fun(x0), [y1, ..., yK])`` with y_k = d^k/dt^k fun(x(t)) at t = 0.
Internally coefficients are normalised (X_k = v_k / k!).
"""

import math

import numpy as np

from .._array import wrap


class _TaylorAt:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _TaylorAtRef(self.arr, idx)


class _TaylorAtRef:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        value = lift(value, self.arr.order)
        out = self.arr.c.copy()
        out[(slice(None),) + _as_tuple(self.idx)] = value.c
        return TaylorArray(out)


def _as_tuple(idx):
    return idx if isinstance(idx, tuple) else (idx,)


class TaylorArray:
    """sum_k c[k] t^k with array-valued coefficients; c has shape (order, *shape)."""

    __array_ufunc__ = None   # make ndarray defer to our reflected operators

    def __init__(self, c):
        self.c = np.asarray(c)

    @property
    def order(self):
        return self.c.shape[0]

    @property
    def shape(self):
        return self.c.shape[1:]

    @property
    def ndim(self):
        return self.c.ndim - 1

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def at(self):
        return _TaylorAt(self)

    @property
    def T(self):
        return self.map_linear(np.transpose)

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __getitem__(self, idx):
        return TaylorArray(self.c[(slice(None),) + _as_tuple(idx)])

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return TaylorArray(self.c.reshape((self.order,) + tuple(shape)))

    def map_linear(self, op):
        return TaylorArray(np.stack([np.asarray(op(ck)) for ck in self.c]))

    # ---- arithmetic -----------------------------------------------------------------------------
    def __neg__(self):
        return TaylorArray(-self.c)

    def __pos__(self):
        return self

    def __add__(self, other):
        o = lift(other, self.order)
        return TaylorArray(_bc(self.c, o.c, np.add))

    __radd__ = __add__

    def __sub__(self, other):
        o = lift(other, self.order)
        return TaylorArray(_bc(self.c, o.c, np.subtract))

    def __rsub__(self, other):
        o = lift(other, self.order)
        return TaylorArray(_bc(o.c, self.c, np.subtract))

    def __mul__(self, other):
        if not isinstance(other, TaylorArray):
            other = np.asarray(other)
            return TaylorArray(np.stack([ck * other for ck in self.c]))
        K = self.order
        terms = []
        for k in range(K):
            acc = 0.0
            for j in range(k + 1):
                acc = acc + self.c[j] * other.c[k - j]
            terms.append(acc)
        return TaylorArray(np.stack(np.broadcast_arrays(*terms)))

    __rmul__ = __mul__

    def __truediv__(self, other):
        if not isinstance(other, TaylorArray):
            other = np.asarray(other)
            return TaylorArray(np.stack([ck / other for ck in self.c]))
        K = self.order
        out = []
        for k in range(K):
            acc = self.c[k]
            for j in range(1, k + 1):
                acc = acc - other.c[j] * out[k - j]
            out.append(acc / other.c[0])
        return TaylorArray(np.stack(np.broadcast_arrays(*out)))

    def __rtruediv__(self, other):
        return lift(other, self.order) / self

    def __pow__(self, r):
        if isinstance(r, (int, np.integer)) and r >= 0:
            out = lift(np.ones(self.shape), self.order)
            for _ in range(int(r)):
                out = out * self
            return out
        # x y' = r x' y   for y = x^r with a real exponent (needs x0 != 0)
        K = self.order
        out = [self.c[0] ** r]
        for k in range(1, K):
            acc = 0.0
            for j in range(1, k + 1):
                acc = acc + (r * j - (k - j)) * self.c[j] * out[k - j]
            out.append(acc / (k * self.c[0]))
        return TaylorArray(np.stack(np.broadcast_arrays(*out)))


def _bc(a, b, op):
    """Broadcast the trailing (value) axes of two coefficient stacks of equal order."""
    if a.ndim < b.ndim:
        a = a.reshape((a.shape[0],) + (1,) * (b.ndim - a.ndim) + a.shape[1:])
    elif b.ndim < a.ndim:
        b = b.reshape((b.shape[0],) + (1,) * (a.ndim - b.ndim) + b.shape[1:])
    return op(a, b)


def lift(x, order):
    if isinstance(x, TaylorArray):
        return x
    x = np.asarray(x)
    c = np.zeros((order,) + x.shape, dtype=np.result_type(x.dtype, np.float64))
    c[0] = x
    return TaylorArray(c)


def _order_of(objs):
    for o in objs:
        if isinstance(o, TaylorArray):
            return o.order
        if isinstance(o, (list, tuple)):
            k = _order_of(o)
            if k:
                return k
    return 0


_SEQUENCE_FUNCS = {"concatenate", "stack", "array", "asarray", "vstack", "hstack"}
_LINEAR_FUNCS = {"roll", "reshape", "ravel", "flip", "pad", "transpose", "convolve", "sum", "squeeze", "real", "diag"}


def taylor_function(name, fn, args, kwargs):
    """Route a jnp call with Taylor arguments through the per-coefficient action of a (multi)linear function."""
    order = _order_of(list(args) + list(kwargs.values()))
    if name in _SEQUENCE_FUNCS:
        seq = args[0]
        if isinstance(seq, TaylorArray):
            return seq
        elems = [lift(e, order) for e in seq]
        rest = args[1:]
        return TaylorArray(np.stack([fn([e.c[k] for e in elems], *rest, **kwargs) for k in range(order)]))
    if name == "split":
        x = args[0]
        parts = [fn(x.c[k], *args[1:], **kwargs) for k in range(order)]
        return [TaylorArray(np.stack([parts[k][i] for k in range(order)])) for i in range(len(parts[0]))]
    if name in ("ones_like", "zeros_like"):
        return wrap(fn(np.empty(args[0].shape), *args[1:], **kwargs))
    if name in _LINEAR_FUNCS:
        x = args[0]
        if not isinstance(x, TaylorArray):
            raise NotImplementedError(f"jnp.{name}: only the first argument may carry a Taylor series")
        if name == "pad" and kwargs.get("mode", "constant") == "constant":
            zero_kw = dict(kwargs)
            zero_kw["constant_values"] = 0.0
            return TaylorArray(np.stack([fn(x.c[k], *args[1:], **(kwargs if k == 0 else zero_kw)) for k in range(order)]))
        return TaylorArray(np.stack([fn(x.c[k], *args[1:], **kwargs) for k in range(order)]))
    if name == "nan_to_num":
        # jnp.nan_to_num is select(isnan(x), 0, x) (+ the same for the infinities); jet's select rule applies the PRIMAL's
        # predicate to every Taylor coefficient, and the constant branch has a zero series
        x = args[0]
        bad = ~np.isfinite(x.c[0])
        return TaylorArray(np.stack([fn(x.c[0], *args[1:], **kwargs)] + [np.where(bad, 0.0, x.c[k]) for k in range(1, order)]))
    raise NotImplementedError(f"jnp.{name} has no Taylor-mode rule in the shim")


def jet(fun, primals, series):
    (x,) = primals
    (terms,) = series
    terms = list(terms)
    K = len(terms)
    x = np.asarray(x, dtype=float)
    coeffs = [x] + [np.asarray(terms[j - 1], dtype=float) / math.factorial(j) for j in range(1, K + 1)]
    out = fun(TaylorArray(np.stack(coeffs)))
    out = lift(out, K + 1)
    primal_out = wrap(np.array(out.c[0]))
    series_out = [wrap(math.factorial(j) * out.c[j]) for j in range(1, K + 1)]
    return primal_out, series_out
