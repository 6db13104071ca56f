"""Initialisation routines (one-off, not on the per-step hot path).  Mirrors tornadox/init.py.

* ``TaylorMode``  (init.py:22-103): exact derivatives y, y', ..., y^(nu) at t0 and a zero factor.  The
  reference uses ``jax.experimental.jet``; here the in-scope vector fields (all polynomial) are pushed
  through a truncated power-series algebra on device tensors.
* ``Stack``       (init.py:359-391): [y0, f(y0), 0, ...] with a diag(0, 0, 1e3, ...) factor
  (``use_df=False`` only -- the dense ``df @ f`` variant is O(d^2) and out of scope).
* ``RungeKutta``  (init.py:109-315, ``use_df=False``): SciPy RK45 points fitted by a square-root Kalman filter + RTS smoother
  over the (n, n) Kronecker factor; the d-wide mean rows stay on the device, the n x n algebra runs in NumPy.
"""

import abc
import math

import numpy as np
import torch

from . import _lib
from .ivp import as_device_tensor


class InitializationRoutine(abc.ABC):
    @abc.abstractmethod
    def __call__(self, f, df, y0, t0, num_derivatives):
        raise NotImplementedError


def _spec_of(f, y0=None):
    """The in-kernel vector field a callable stands for, or -- for any other callable -- the EXTERNAL field: the fused
    kernels then read f (and the Jacobian diagonal) from arrays the caller's own torch code produces at the predicted state
    (``StepEngine.predict_state`` / ``set_external``), which is how the reference treats every f (odefilter.py:346-352)."""
    spec = getattr(f, "spec", None)
    if spec is not None:
        return spec
    if not callable(f):
        raise ValueError("the vector field must be callable: f(t, y) -> dy/dt on device tensors")
    if y0 is None:
        raise ValueError("a user-defined vector field needs y0 to size the state")
    from .engine import KernelSpec

    d = int(as_device_tensor(y0).reshape(-1).shape[0])
    return KernelSpec(_lib.TDX_IVP_EXTERNAL, (), 1, d, 1, getattr(f, "__name__", "user-defined f"))


def is_external(f):
    return getattr(f, "spec", None) is None


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

    def __truediv__(self, other):
        if not isinstance(other, _PowerSeries):
            return _PowerSeries(self.c / other)
        out = torch.zeros_like(self.c)
        for k in range(self.order):          # q b = a  =>  q_k = (a_k - sum_{j=1..k} b_j q_{k-j}) / b_0
            acc = self.c[k].clone()
            for j in range(1, k + 1):
                acc -= other.c[j] * out[k - j]
            out[k] = acc / other.c[0]
        return _PowerSeries(out)

    def pow(self, r):
        """Real power of a series with nonzero constant term: y = x^r solves x y' = r x' y."""
        out = torch.zeros_like(self.c)
        out[0] = self.c[0] ** r
        for k in range(1, self.order):
            acc = torch.zeros_like(self.c[0])
            for j in range(1, k + 1):
                acc += (r * j - (k - j)) * self.c[j] * out[k - j]
            out[k] = acc / (k * self.c[0])
        return _PowerSeries(out)

    def entry(self, i):
        return _PowerSeries(self.c[:, i:i + 1])

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
        mu, julia = spec.params[0], len(spec.params) > 1 and spec.params[1] != 0.0

        def rhs(y):
            y1, y2 = y.fields(2)
            if julia:                                   # ivp.py:52-73
                return [y2, mu * ((1.0 - y1.square()) * y2 - y1)]
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

    elif kind == _lib.TDX_IVP_THREEBODY:
        mu = 0.012277471
        mp = 1 - mu

        def rhs(y):
            y1, y2, v1, v2 = (y.entry(i) for i in range(4))
            D1 = ((y1 + mu).square() + y2.square()).pow(1.5)
            D2 = ((y1 - mp).square() + y2.square()).pow(1.5)
            return [v1, v2,
                    y1 + 2.0 * v2 - mp * (y1 + mu) / D1 - mu * (y1 - mp) / D2,
                    y2 - 2.0 * v1 - mp * y2 / D1 - mu * y2 / D2]

    elif kind == _lib.TDX_IVP_PLEIADES:

        def rhs(y):
            # pairwise differences as (7, 7) series; the diagonal (a body with itself: 0/0 -> 0 in the reference, ivp.py:384-385)
            # is given a unit distance and masked out of the sums
            x, yy = y.c[:, 0:7], y.c[:, 7:14]
            eye = torch.eye(7, dtype=y.c.dtype, device=y.c.device)
            dx = _PowerSeries(x[:, None, :] - x[:, :, None])          # [k, i, j] = x_j - x_i
            dy = _PowerSeries(yy[:, None, :] - yy[:, :, None])
            sq = dx.square() + dy.square()
            sq.c[0] += eye
            r3 = sq.pow(1.5)
            mass = torch.arange(1, 8, dtype=y.c.dtype, device=y.c.device)[None, :] * (1.0 - eye)
            ddx = _PowerSeries(((dx / r3).c * mass).sum(-1))
            ddy = _PowerSeries(((dy / r3).c * mass).sum(-1))
            return [_PowerSeries(y.c[:, 14:28]), ddx, ddy]

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
        if is_external(fun):
            return TaylorMode.taylor_mode_autodiff(fun, y0, t0, num_derivatives)
        spec = _spec_of(fun)
        rhs = _series_rhs(spec)
        y0 = as_device_tensor(y0).reshape(-1)
        coeffs = [y0]
        for k in range(num_derivatives):
            series = _PowerSeries(torch.stack(coeffs))
            fk = torch.cat([part.c[k] for part in rhs(series)])
            coeffs.append(fk / (k + 1))
        return torch.stack([math.factorial(k) * c for k, c in enumerate(coeffs)])


    @staticmethod
    def taylor_mode_autodiff(fun, y0, t0, num_derivatives):
        """Taylor-mode initialisation of a user-defined vector field written in differentiable torch operations (the
        reference pushes any f through ``jax.experimental.jet``, init.py:63-103).  With the autonomous extension
        z = [y, t], z' = F(z) = [f(t, y), 1]:  z^(k+1) = d/dt z^(k) = J_{z^(k)}(z) F(z), i.e. one forward-mode
        derivative (``torch.func.jvp``) per order, nested.  Cost grows like 2^nu evaluations of f: fine for nu <= 6."""
        y0 = as_device_tensor(y0).reshape(-1)
        d = y0.shape[0]
        z0 = torch.cat([y0, torch.as_tensor([float(t0)], dtype=y0.dtype, device=y0.device)])

        def F(z):
            return torch.cat([fun(z[d], z[:d]).reshape(-1), torch.ones(1, dtype=z.dtype, device=z.device)])

        def derivative(k):
            if k == 0:
                return lambda z: z
            prev = derivative(k - 1)
            return lambda z: torch.func.jvp(prev, (z,), (F(z),))[1]

        rows = [derivative(k)(z0)[:d] for k in range(num_derivatives + 1)]
        return torch.stack(rows).contiguous()


class Stack(InitializationRoutine):
    """[y0, f(t0, y0), 0, ...] with large uncertainty on the unknown derivatives (init.py:359-391).
    ``use_df=True`` also fills y'' = df(t0, y0) @ f(t0, y0) from the caller's DENSE Jacobian (init.py:377-385; O(d^2), small
    problems -- the in-kernel vector fields carry no dense Jacobian)."""

    def __init__(self, use_df=False):
        self.use_df = bool(use_df)

    def __call__(self, f, df, y0, t0, num_derivatives):
        y0 = as_device_tensor(y0).reshape(-1)
        n = num_derivatives + 1
        fy = f(t0, y0)
        m = torch.zeros((n, y0.shape[0]), dtype=y0.dtype, device=y0.device)
        m[0], m[1] = y0, fy
        known = 2
        if self.use_df:
            if df is None:
                raise NotImplementedError(
                    "Stack(use_df=True) needs the dense Jacobian df(t, y) (O(d^2)); the vector fields of tornadox_b200.ivp do not "
                    "provide one -- use Stack(use_df=False) or TaylorMode()")
            if n > 2:
                m[2] = torch.as_tensor(df(t0, y0), dtype=y0.dtype, device=y0.device) @ fy
                known = 3
        diag = torch.tensor([0.0] * min(known, n) + [1e3] * max(n - known, 0), dtype=y0.dtype, device=y0.device)
        return m, torch.diag(diag)

    def __repr__(self):
        return f"{self.__class__.__name__}(use_df={self.use_df})"


class RungeKutta(InitializationRoutine):
    """Fit the stacked initial state to a few Runge-Kutta steps (init.py:109-315; ``use_df=False`` only).

    The filter / smoother of the reference acts on the Kronecker structure: an (n, n) factor shared by all coordinates and
    an (n, d) mean.  Here the factor algebra (QR, Cholesky solves) runs in NumPy on the host exactly as in the reference,
    and every d-wide operation is a torch expression on the device.  The RK points come from
    ``scipy.integrate.solve_ivp`` (init.py:128-144) with the vector field evaluated by the CUDA library.
    """

    def __init__(self, dt=0.01, method="RK45", use_df=False):
        self.dt = dt
        self.method = method
        self.stack_initvals = Stack(use_df=use_df)

    def __repr__(self):
        return f"{self.__class__.__name__}(dt={self.dt}, method={self.method})"

    def __call__(self, f, df, y0, t0, num_derivatives):
        y0 = as_device_tensor(y0).reshape(-1)
        ts, ys = self.rk_data(f=f, t0=t0, dt=self.dt, num_steps=num_derivatives + 1, y0=y0, method=self.method)
        m, sc = self.stack_initvals(f=f, df=df, y0=y0, t0=t0, num_derivatives=num_derivatives)
        return RungeKutta.rk_init_improve(m=m, sc=sc, t0=t0, ts=ts, ys=ys)

    @staticmethod
    def rk_data(f, t0, dt, num_steps, y0, method):
        """init.py:128-144: fixed steps forced through t_eval, tolerances of 1e12."""
        import scipy.integrate

        y0 = as_device_tensor(y0).reshape(-1)
        t_eval = np.arange(t0, t0 + num_steps * dt, dt)

        def fun(t, y):
            return f(t, torch.as_tensor(y, dtype=y0.dtype, device=y0.device)).cpu().numpy()

        sol = scipy.integrate.solve_ivp(fun=fun, t_span=(min(t_eval), max(t_eval)), y0=y0.cpu().numpy(), atol=1e12, rtol=1e12,
                                        t_eval=t_eval, method=method)
        return sol.t, torch.as_tensor(sol.y.T.copy(), dtype=y0.dtype, device=y0.device)

    @staticmethod
    def _factor_step(sc, phi, sq, p, pinv):
        """The (n, n) part of init.py:280-313 in NumPy: returns (sc_new, sc_pred-free quantities the mean update needs)."""
        import scipy.linalg

        sc = pinv[:, None] * sc
        x = phi @ sc
        stacked = np.vstack((x.T, sq.T))
        sc_pred = np.linalg.qr(stacked, mode="r").T                 # sqrt.propagate_cholesky_factor
        cross = (x @ sc.T).T
        sgain = scipy.linalg.cho_solve((sc_pred, True), cross.T).T
        h = (p[:, None] * sc_pred)[0, :]
        kgain = (sc_pred @ h.T) / (h @ h.T)
        sc_loc = sc_pred - kgain[:, None] * h[None, :]
        return p[:, None] * sc_loc, sgain, x, kgain

    @staticmethod
    def rk_init_improve(m, sc, t0, ts, ys):
        """init.py:146-277."""
        nu = m.shape[0] - 1
        dev, dtype = m.device, m.dtype
        A, _ = _lib.prior(nu)
        phi, sq = np.asarray(A), _lib.lapack_diffusion_sqrt(nu)
        on_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=dev)
        sc = np.asarray(sc.detach().cpu().numpy() if isinstance(sc, torch.Tensor) else sc, dtype=float)

        init = dict(m=m, sc=sc, m_pred=None, sgain=None, x=None, p=None, pinv=None)
        states, carry, t_loc = [], init, t0
        for t, y in zip(ts[1:], ys[1:]):
            p, pinv = _lib.nordsieck(nu, float(t - t_loc))
            sc_new, sgain, x, kgain = RungeKutta._factor_step(carry["sc"], phi, sq, p, pinv)
            m_pred = on_dev(phi) @ (on_dev(pinv)[:, None] * carry["m"])                       # (n, d) on the device
            z = on_dev(p)[0] * m_pred[0]
            m_new = on_dev(p)[:, None] * (m_pred - on_dev(kgain)[:, None] * (z - y)[None, :])
            carry = dict(m=m_new, sc=sc_new, m_pred=m_pred, sgain=sgain, x=x, p=p, pinv=pinv)
            states.append(carry)
            t_loc = t
        if not states:
            return m, on_dev(sc)

        prev, m_fut, sc_fut = states[-1], states[-1]["m"], states[-1]["sc"]
        n = nu + 1
        zeros = np.zeros((n, n))
        for fs in list(reversed(states[:-1])) + [init]:
            p, pinv = prev["p"], prev["pinv"]
            sgain = prev["sgain"]
            # kalman.smoother_step_sqrt (kalman.py:47-64) in the preconditioned coordinates of the LATER step
            m_s = on_dev(pinv)[:, None] * fs["m"]
            new_mean = m_s - on_dev(sgain) @ (prev["m_pred"] - on_dev(pinv)[:, None] * m_fut)
            M = np.block([[prev["x"].T, (pinv[:, None] * fs["sc"]).T], [sq.T, zeros.T],
                          [zeros.T, (pinv[:, None] * sc_fut).T @ sgain.T]])
            new_sc = np.linalg.qr(M, mode="r")[n:2 * n, n:].T
            m_fut, sc_fut = on_dev(p)[:, None] * new_mean, p[:, None] * new_sc
            prev = fs
        return m_fut, on_dev(sc_fut)
