"""Initialisation routines (one-off, not on the per-step hot path).  Mirrors tornadox/init.py.

* ``TaylorMode``  (init.py:22-103): exact derivatives y, y', ..., y^(nu) at t0 and a zero factor.  The
  reference uses ``jax.experimental.jet``; here the in-scope vector fields (all polynomial) are pushed
  through a truncated power-series algebra on device tensors.
* ``Stack``       (init.py:359-391): [y0, f(y0), 0, ...] with a diag(0, 0, 1e3, ...) factor
  (``use_df=False`` only -- the dense ``df @ f`` variant is O(d^2) and out of scope).
"""

import abc
import math

import torch

from . import _lib
from .ivp import as_device_tensor


class InitializationRoutine(abc.ABC):
    @abc.abstractmethod
    def __call__(self, f, df, y0, t0, num_derivatives):
        raise NotImplementedError


def _spec_of(f):
    spec = getattr(f, "spec", None)
    if spec is None:
        raise ValueError(
            "tornadox_b200 only integrates the vector fields of tornadox_b200.ivp (evaluated in-kernel); "
            "an arbitrary Python callable has no fused kernel and there is no CPU fallback"
        )
    return spec


class _PowerSeries:
    """sum_k c[k] h^k, coefficients stacked along axis 0 of a device tensor."""

    def __init__(self, c):
        self.c = c

    @property
    def order(self):
        return self.c.shape[0]

    def _coerce(self, other):
        if isinstance(other, _PowerSeries):
            return other
        c = torch.zeros_like(self.c)
        c[0] = other
        return _PowerSeries(c)

    def __add__(self, other):
        return _PowerSeries(self.c + self._coerce(other).c)

    __radd__ = __add__

    def __neg__(self):
        return _PowerSeries(-self.c)

    def __sub__(self, other):
        return _PowerSeries(self.c - self._coerce(other).c)

    def __rsub__(self, other):
        return _PowerSeries(self._coerce(other).c - self.c)

    def __mul__(self, other):
        if not isinstance(other, _PowerSeries):
            return _PowerSeries(self.c * other)
        out = torch.zeros_like(self.c)
        for k in range(self.order):          # Cauchy product, truncated
            for j in range(k + 1):
                out[k] += self.c[j] * other.c[k - j]
        return _PowerSeries(out)

    __rmul__ = __mul__

    def __truediv__(self, scalar):
        return _PowerSeries(self.c / scalar)

    def square(self):
        return self * self

    def cube(self):
        return self * self * self

    def map_linear(self, op):
        return _PowerSeries(torch.stack([op(ck) for ck in self.c]))

    def fields(self, nf):
        return [_PowerSeries(part) for part in self.c.view(self.order, nf, -1).unbind(1)]


def _series_rhs(spec):
    """The vector field of ``spec`` acting on a _PowerSeries (same arithmetic as the kernels)."""
    kind = spec.kind
    if kind == _lib.TDX_IVP_VANDERPOL:
        (mu,) = spec.params

        def rhs(y):
            y1, y2 = y.fields(2)
            return [y2, mu * (1.0 - y1.square()) * y2 - y1]

    elif kind == _lib.TDX_IVP_LORENZ96:
        (forcing,) = spec.params

        def rhs(y):
            nxt = y.map_linear(lambda c: torch.roll(c, -1))
            prv2 = y.map_linear(lambda c: torch.roll(c, 2))
            prv1 = y.map_linear(lambda c: torch.roll(c, 1))
            return [(nxt - prv2) * prv1 - y + forcing]

    elif kind == _lib.TDX_IVP_BRUSSELATOR:
        N = spec.grid_cols
        const = (1.0 / 50.0) * (N + 1) ** 2

        def second_difference(s, pad):
            out = torch.empty_like(s.c)
            for k in range(s.order):
                ck = s.c[k]
                edge = pad if k == 0 else 0.0     # the boundary values are constants
                left = torch.cat([ck.new_full((1,), edge), ck[:-1]])
                right = torch.cat([ck[1:], ck.new_full((1,), edge)])
                out[k] = left - 2.0 * ck + right
            return _PowerSeries(out)

        def rhs(y):
            u, v = y.fields(2)
            uuv = u.square() * v
            return [1.0 + uuv - 4.0 * u + const * second_difference(u, 1.0),
                    3.0 * u - uuv + const * second_difference(v, 3.0)]

    elif kind == _lib.TDX_IVP_FHN2D:
        dx, a, b, k_, tau = spec.params
        nx, ny = spec.grid_rows, spec.grid_cols
        inv_dx2 = 1.0 / (dx * dx)

        def laplace(s):
            def op(c):
                g = c.view(nx, ny)
                p = torch.nn.functional.pad(g[None, None], (1, 1, 1, 1), mode="replicate")[0, 0]
                lap = (p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:] - 4.0 * g) * inv_dx2
                return lap.reshape(-1)

            return s.map_linear(op)

        def rhs(y):
            u, v = y.fields(2)
            return [a * laplace(u) + u - u.cube() - v + k_, (b * laplace(v) + u - v) / tau]

    elif kind == _lib.TDX_IVP_BURGERS1D:
        dx, nu_ = spec.params

        def rhs(y):
            def shifted(s, lo, hi):
                return _PowerSeries(s.c[:, lo:hi])

            n = y.c.shape[1]
            mid, left, right = shifted(y, 1, n - 1), shifted(y, 0, n - 2), shifted(y, 2, n)
            interior = -(mid * (mid - left)) / dx + nu_ * (right - 2.0 * mid + left) / (dx * dx)
            x0, x1, xm = shifted(y, 0, 1), shifted(y, 1, 2), shifted(y, n - 2, n - 1)
            boundary = -(x0 * (x0 - xm)) / dx + nu_ * (x1 - 2.0 * x0 + xm) / (dx * dx)
            return [boundary, interior, boundary]

    elif kind == _lib.TDX_IVP_WAVE1D:
        dx, nu_ = spec.params

        def rhs(y):
            x, v = y.fields(2)
            n = x.c.shape[1]
            cut = lambda s, lo, hi: _PowerSeries(s.c[:, lo:hi])
            interior = nu_ * (cut(x, 2, n) - 2.0 * cut(x, 1, n - 1) + cut(x, 0, n - 2)) / (dx * dx)
            zero = _PowerSeries(torch.zeros_like(x.c[:, :1]))
            return [zero, cut(v, 1, n - 1), zero, cut(x, 0, 1), interior, cut(x, n - 1, n)]

    else:
        raise NotImplementedError(f"no Taylor-mode rule for {spec.name}")
    return rhs


class TaylorMode(InitializationRoutine):
    """Exact Taylor-mode initialisation (init.py:22-103); returns (derivatives (n, d), zeros (n, n))."""

    def __call__(self, f, df, y0, t0, num_derivatives):
        m0 = TaylorMode.taylor_mode(fun=f, y0=y0, t0=t0, num_derivatives=num_derivatives)
        n = num_derivatives + 1
        return m0, torch.zeros((n, n), dtype=m0.dtype, device=m0.device)

    def __repr__(self):
        return f"{self.__class__.__name__}()"

    @staticmethod
    def taylor_mode(fun, y0, t0, num_derivatives):
        """With y(t0 + h) = sum_k Y_k h^k and y' = f(y):  (k + 1) Y_{k+1} = [h^k] f(y(t0 + h))."""
        spec = _spec_of(fun)
        rhs = _series_rhs(spec)
        y0 = as_device_tensor(y0).reshape(-1)
        coeffs = [y0]
        for k in range(num_derivatives):
            series = _PowerSeries(torch.stack(coeffs))
            fk = torch.cat([part.c[k] for part in rhs(series)])
            coeffs.append(fk / (k + 1))
        return torch.stack([math.factorial(k) * c for k, c in enumerate(coeffs)])


class Stack(InitializationRoutine):
    """[y0, f(t0, y0), 0, ...] with large uncertainty on the unknown derivatives (init.py:359-391)."""

    def __init__(self, use_df=False):
        if use_df:
            raise NotImplementedError(
                "Stack(use_df=True) multiplies with the dense Jacobian (O(d^2)); use Stack(use_df=False) or TaylorMode()"
            )

    def __call__(self, f, df, y0, t0, num_derivatives):
        y0 = as_device_tensor(y0).reshape(-1)
        n = num_derivatives + 1
        fy = f(t0, y0)
        m = torch.zeros((n, y0.shape[0]), dtype=y0.dtype, device=y0.device)
        m[0], m[1] = y0, fy
        diag = torch.tensor([0.0, 0.0] + [1e3] * (n - 2), dtype=y0.dtype, device=y0.device)
        return m, torch.diag(diag)

    def __repr__(self):
        return f"{self.__class__.__name__}(use_df=False)"
