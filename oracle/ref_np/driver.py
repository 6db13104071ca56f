"""ORACLE (test infrastructure): step rules and the accept/reject driver loop of the reference.

* step rules            tornadox/step.py:28-115
* perform_full_step     tornadox/odefilter.py:171-223
* solution_generator    tornadox/odefilter.py:80-159  (incl. the ``stop_at`` time stopper :355-373)
* Taylor-mode init      tornadox/init.py:22-103 via truncated power-series arithmetic
  (the reference uses ``jax.experimental.jet``; for the polynomial/rational vector fields in scope
  a Cauchy-product series algebra yields the same derivatives).
"""

import math

import numpy as np

from . import filters, prior

# ---------------------------------------------------------------------------------------------
# step rules                                                          tornadox/step.py
# ---------------------------------------------------------------------------------------------


class ConstantSteps:
    """tornadox/step.py:28-52."""

    def __init__(self, dt):
        self.dt = dt

    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        return self.dt

    def is_accepted(self, scaled_error_estimate):
        return True

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        return None

    def first_dt(self, f, t0, tmax, y0, df, df_diagonal):
        return self.dt


class AdaptiveSteps:
    """tornadox/step.py:55-108."""

    def __init__(self, abstol=1e-4, reltol=1e-2, max_changes=(0.2, 10.0), safety_scale=0.95,
                 min_step=1e-15, max_step=1e15):
        self.abstol = abstol
        self.reltol = reltol
        self.max_changes = max_changes
        self.safety_scale = safety_scale
        self.min_step = min_step
        self.max_step = max_step

    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        """step.py:76-87."""
        if local_convergence_rate is None:
            raise ValueError("Please provide a local convergence rate.")
        small, large = self.max_changes
        ratio = 1.0 / scaled_error_estimate
        change = self.safety_scale * ratio ** (1.0 / local_convergence_rate)
        change = np.maximum(small, np.minimum(change, large))
        return change * previous_dt

    def is_accepted(self, scaled_error_estimate):
        """step.py:90-91."""
        return scaled_error_estimate < 1

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        """step.py:93-104."""
        unscaled_error_estimate = np.asarray(unscaled_error_estimate)
        if (not np.isscalar(unscaled_error_estimate.size)) and (
            unscaled_error_estimate.shape != reference_state.shape
        ):
            raise ValueError("Unscaled error estimate needs same shape as reference state.")
        tolerance = self.abstol + self.reltol * reference_state
        ratio = unscaled_error_estimate / tolerance
        dim = len(ratio) if ratio.ndim > 0 else 1
        return np.linalg.norm(ratio) / np.sqrt(dim)

    def first_dt(self, f, t0, tmax, y0, df, df_diagonal):
        """step.py:106-108."""
        return propose_first_dt(f, t0, y0)


def propose_first_dt(f, t0, y0):
    """step.py:111-115."""
    return 0.01 * np.linalg.norm(y0) / np.linalg.norm(f(t0, y0))


# ---------------------------------------------------------------------------------------------
# Taylor-mode initialisation                                       tornadox/init.py:22-103
# ---------------------------------------------------------------------------------------------


class Series:
    """Truncated power series sum_k c[k] h^k with array-valued coefficients (c has shape (K, ...))."""

    __array_priority__ = 1000

    def __init__(self, c):
        self.c = np.asarray(c, dtype=float)

    @property
    def order(self):
        return self.c.shape[0]

    def _lift(self, other):
        if isinstance(other, Series):
            return other
        shape = np.shape(other) if np.ndim(other) > 0 else (1,) * (self.c.ndim - 1)
        c = np.zeros((self.order,) + shape)
        c[0] = other
        return Series(c)

    def __add__(self, other):
        o = self._lift(other)
        return Series(self.c + o.c)

    __radd__ = __add__

    def __neg__(self):
        return Series(-self.c)

    def __sub__(self, other):
        return self + (-self._lift(other))

    def __rsub__(self, other):
        return self._lift(other) - self

    def __mul__(self, other):
        if not isinstance(other, Series):
            return Series(self.c * other)
        K = self.order
        out = np.zeros(np.broadcast_shapes(self.c.shape, other.c.shape))
        for k in range(K):
            for j in range(k + 1):
                out[k] = out[k] + self.c[j] * other.c[k - j]
        return Series(out)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if not isinstance(other, Series):
            return Series(self.c / other)
        K = self.order
        out = np.zeros(np.broadcast_shapes(self.c.shape, other.c.shape))
        for k in range(K):
            acc = self.c[k] * np.ones(out.shape[1:])
            for j in range(1, k + 1):
                acc = acc - other.c[j] * out[k - j]
            out[k] = acc / other.c[0]
        return Series(out)

    def __pow__(self, r):
        if isinstance(r, (int, np.integer)) and r >= 0:
            out = self._lift(np.ones(self.c.shape[1:]))
            for _ in range(int(r)):
                out = out * self
            return out
        # real exponent: y = x^r  ->  x y' = r x' y   (x0 != 0)
        K = self.order
        out = np.zeros_like(self.c)
        out[0] = self.c[0] ** r
        for k in range(1, K):
            acc = 0.0
            for j in range(1, k + 1):
                acc = acc + (r * j - (k - j)) * self.c[j] * out[k - j]
            out[k] = acc / (k * self.c[0])
        return Series(out)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return Series(self.c[(slice(None),) + idx])

    def linear(self, op):
        """Apply a linear map coefficient-wise (roll, reshape, Laplacian, split ...)."""
        return Series(np.stack([op(ck) for ck in self.c]))


def series_stack(items, order):
    """Stack scalar-valued Series/floats into one vector-valued Series."""
    cols = []
    for it in items:
        if isinstance(it, Series):
            cols.append(it.c)
        else:
            c = np.zeros((order,) + np.shape(it))
            c[0] = it
            cols.append(c)
    return Series(np.stack(cols, axis=1))


def series_concat(items):
    return Series(np.concatenate([it.c for it in items], axis=1))


def taylor_mode(series_f, y0, t0, num_derivatives):
    """Derivatives [y, y', ..., y^(nu)] at t0 -- the quantity init.py:32-66 computes with ``jet``.

    ``series_f(t_series, y_series)`` must evaluate the vector field on ``Series`` arguments.
    With y(t0+h) = sum_k Y_k h^k and y' = f(t, y):  (k+1) Y_{k+1} = [h^k] f(t0 + h, y(t0 + h)).
    """
    y0 = np.asarray(y0, dtype=float)
    nu = num_derivatives
    coeffs = [y0]
    for k in range(nu):
        order = k + 1
        yc = np.zeros((order,) + y0.shape)
        for j, cj in enumerate(coeffs):
            yc[j] = cj
        tc = np.zeros((order,))
        tc[0] = t0
        if order > 1:
            tc[1] = 1.0
        fk = series_f(Series(tc), Series(yc)).c[k]
        coeffs.append(fk / (k + 1))
    derivs = [math.factorial(k) * c for k, c in enumerate(coeffs)]
    return np.stack(derivs)


def taylor_init(series_f, y0, t0, num_derivatives):
    """``TaylorMode.__call__`` -- init.py:25-29: (derivative stack, zero (n, n) factor)."""
    n = num_derivatives + 1
    return taylor_mode(series_f, y0, t0, num_derivatives), np.zeros((n, n))


# ---------------------------------------------------------------------------------------------
# Runge-Kutta initialisation                                         tornadox/init.py:109-315
# ---------------------------------------------------------------------------------------------
def stack_init(f, y0, t0, num_derivatives):
    """``Stack(use_df=False)`` -- init.py:382-391: [y0, f(y0), 0, ...] with a diag(0, 0, 1e3, ...) factor."""
    n = num_derivatives + 1
    y0 = np.asarray(y0, dtype=float)
    m = np.stack([y0, f(t0, y0)] + [np.zeros_like(y0)] * (n - 2))
    return m, np.diag(np.array([0.0, 0.0] + [1e3] * (n - 2)))


def rk_data(f, t0, dt, num_steps, y0, method="RK45"):
    """init.py:128-144: SciPy's solve_ivp with tolerances so loose that it takes its own fixed steps, read at t_eval."""
    import scipy.integrate

    t_eval = np.arange(t0, t0 + num_steps * dt, dt)
    sol = scipy.integrate.solve_ivp(fun=f, t_span=(min(t_eval), max(t_eval)), y0=y0, atol=1e12, rtol=1e12, t_eval=t_eval,
                                    method=method)
    return sol.t, sol.y.T


def _rk_forward_step(y, sc, m, sq, p, pinv, phi):
    """init.py:280-313."""
    import scipy.linalg

    m = pinv[:, None] * m
    sc = pinv[:, None] * sc
    m_pred = phi @ m
    x = phi @ sc
    sc_pred = prior.propagate_cholesky(x, sq)
    cross = (x @ sc.T).T
    sgain = scipy.linalg.cho_solve((sc_pred, True), cross.T).T
    sc_pred_np = p[:, None] * sc_pred
    h = sc_pred_np[0, :]
    s = h @ h.T
    kgain = (sc_pred @ h.T) / s
    z = (p[:, None] * m_pred)[0]
    m_loc = m_pred - kgain[:, None] * (z - y)[None, :]
    sc_loc = sc_pred - kgain[:, None] * h[None, :]
    return p[:, None] * m_loc, p[:, None] * sc_loc, m_pred, sc_pred, sgain, x


def _smoother_step_sqrt(m, sc, m_fut, sc_fut, sgain, sq, mp, x):
    """kalman.py:47-64."""
    new_mean = m - sgain @ (mp - m_fut)
    n = m.shape[0]
    zeros = np.zeros((n, n))
    M = np.block([[x.T, sc.T], [sq.T, zeros.T], [zeros.T, sc_fut.T @ sgain.T]])
    R = np.linalg.qr(M, mode="r")
    return new_mean, R[n:2 * n, n:].T


def rk_init_improve(m, sc, t0, ts, ys):
    """init.py:146-277: fit the stacked initial state to the RK points with a Kalman filter + RTS smoother (both in
    square-root form, Nordsieck-preconditioned) over the n x n Kronecker factor."""
    nu = m.shape[0] - 1
    phi, sq = prior.transition_1d(nu), prior.diffusion_sqrt_1d(nu)
    init = dict(m=m, sc=sc, m_pred=None, sgain=None, x=None, p=None, pinv=None)
    states, carry, t_loc = [], init, t0
    for t, y in zip(ts[1:], ys[1:]):
        p, pinv = prior.nordsieck_raw(nu, t - t_loc)
        mm, ss, m_pred, sc_pred, sgain, x = _rk_forward_step(y, carry["sc"], carry["m"], sq, p, pinv, phi)
        carry = dict(m=mm, sc=ss, m_pred=m_pred, sgain=sgain, x=x, p=p, pinv=pinv)
        states.append(carry)
        t_loc = t
    prev, m_fut, sc_fut = states[-1], states[-1]["m"], states[-1]["sc"]
    for fs in list(reversed(states[:-1])) + [init]:
        p, pinv = prev["p"], prev["pinv"]
        mm, ss = _smoother_step_sqrt(pinv[:, None] * fs["m"], pinv[:, None] * fs["sc"], pinv[:, None] * m_fut,
                                     pinv[:, None] * sc_fut, prev["sgain"], sq, prev["m_pred"], prev["x"])
        m_fut, sc_fut = p[:, None] * mm, p[:, None] * ss
        prev = fs
    return m_fut, sc_fut


def runge_kutta_init(f, y0, t0, num_derivatives, dt=0.01, method="RK45"):
    """``RungeKutta(dt, method, use_df=False).__call__`` -- init.py:119-126."""
    ts, ys = rk_data(f, t0, dt, num_derivatives + 1, np.asarray(y0, dtype=float), method)
    m, sc = stack_init(f, y0, t0, num_derivatives)
    return rk_init_improve(m, sc, t0, ts, ys)


# Series versions of the in-scope vector fields (same arithmetic as oracle.ref_np.problems).


def series_vanderpol(mu):
    def f(_, Y):
        return series_stack([Y[1], mu * (1.0 - Y[0] ** 2) * Y[1] - Y[0]], Y.order)

    return f


def series_vanderpol_julia(mu):
    def f(_, Y):
        return series_stack([Y[1], mu * ((1.0 - Y[0] ** 2) * Y[1] - Y[0])], Y.order)

    return f


def series_lorenz96(forcing):
    def f(_, y):
        A = y.linear(lambda c: np.roll(c, -1))
        B = y.linear(lambda c: np.roll(c, 2))
        C = y.linear(lambda c: np.roll(c, 1))
        return (A - B) * C - y + forcing

    return f


def series_brusselator(N):
    const = (1.0 / 50.0) * (N + 1) ** 2

    def lap(c, pad, k):
        padded = np.concatenate([[pad if k == 0 else 0.0], c, [pad if k == 0 else 0.0]])
        return np.convolve(padded, np.array([1.0, -2.0, 1.0]), mode="valid")

    def f(_, y):
        u, v = y[:N], y[N:]
        conv_u = Series(np.stack([lap(c, 1.0, k) for k, c in enumerate(u.c)]))
        conv_v = Series(np.stack([lap(c, 3.0, k) for k, c in enumerate(v.c)]))
        u_new = 1.0 + u**2 * v - 4 * u + const * conv_u
        v_new = 3 * u - u**2 * v + const * conv_v
        return series_concat([u_new, v_new])

    return f


def series_burgers_1d(dx, diffusion_param=0.02):
    """ivp.py:86-96 on Series arguments."""

    def f(_, x):
        n = x.c.shape[1]
        mid, left, right = x[1:n - 1], x[0:n - 2], x[2:n]
        interior = -(mid * (mid - left)) / dx + diffusion_param * (right - 2.0 * mid + left) / (dx**2)
        x0, x1, xm = x[0:1], x[1:2], x[n - 2:n - 1]
        boundary = -(x0 * (x0 - xm)) / dx + diffusion_param * (x1 - 2.0 * x0 + xm) / (dx**2)
        return series_concat([boundary, interior, boundary])

    return f


def series_wave_1d(dx, diffusion_param=0.01):
    """ivp.py:124-130 on Series arguments."""

    def f(_, y):
        n = y.c.shape[1] // 2
        x, v = y[0:n], y[n:2 * n]
        interior = diffusion_param * (x[2:n] - 2.0 * x[1:n - 1] + x[0:n - 2]) / (dx**2)
        zero = x[0:1] * 0.0
        return series_concat([zero, v[1:n - 1], zero, x[0:1], interior, x[n - 1:n]])

    return f


def series_fhn_2d(nx, ny, dx, a=2.8e-4, b=5e-3, k=-0.005, tau=0.1):
    from .problems import laplace_2d

    def lap(s):
        return s.linear(lambda c: laplace_2d(c.reshape((nx, ny)), dx=dx).reshape((-1,)))

    def f(_, x):
        half = nx * ny
        u, v = x[:half], x[half:]
        u_new = a * lap(u) + u - u**3 - v + k
        v_new = (b * lap(v) + u - v) / tau
        return series_concat([u_new, v_new])

    return f


def series_threebody():
    """tornadox/ivp.py:188-216 (used only to pin ``taylor_mode`` against tests/test_init.py:45-148)."""
    mu = 0.012277471
    mp = 1 - mu

    def f(_, Y):
        D1 = ((Y[0] + mu) ** 2 + Y[1] ** 2) ** (3 / 2)
        D2 = ((Y[0] - mp) ** 2 + Y[1] ** 2) ** (3 / 2)
        y1p = Y[0] + 2 * Y[3] - mp * (Y[0] + mu) / D1 - mu * (Y[0] - mp) / D2
        y2p = Y[1] - 2 * Y[2] - mp * Y[1] / D1 - mu * Y[1] / D2
        return series_stack([Y[2], Y[3], y1p, y2p], Y.order)

    return f


def series_pleiades():
    """tornadox/ivp.py:344-406 on Series arguments; the self-interaction terms (0/0 -> 0 in the reference) are skipped."""

    def f(_, u):
        comps = [u[14 + k] for k in range(14)]
        acc = {}
        for i in range(7):
            ddx, ddy = 0.0, 0.0
            for j in range(7):
                if i == j:
                    continue
                dxs, dys = u[j] - u[i], u[7 + j] - u[7 + i]
                r3 = (dxs ** 2 + dys ** 2) ** (3 / 2)
                ddx = ddx + (j + 1) * dxs / r3
                ddy = ddy + (j + 1) * dys / r3
            acc[i] = (ddx, ddy)
        comps += [acc[i][0] for i in range(7)] + [acc[i][1] for i in range(7)]
        return series_stack(comps, u.order)

    return f


# ---------------------------------------------------------------------------------------------
# driver loop                                                      tornadox/odefilter.py
# ---------------------------------------------------------------------------------------------


class TimeStopper:
    """tornadox/odefilter.py:355-373."""

    def __init__(self, locations):
        self._locations = iter(locations)
        self._next_location = next(self._locations)

    def adjust_dt_to_time_stops(self, t, dt):
        if t >= self._next_location:
            try:
                self._next_location = next(self._locations)
            except StopIteration:
                self._next_location = np.inf
        if t + dt > self._next_location:
            dt = self._next_location - t
        return dt


def perform_full_step(attempt, steprule, nu, state, initial_dt, tmax):
    """tornadox/odefilter.py:171-223.  ``attempt(state, dt) -> (proposed_state, info)``."""
    dt = initial_dt
    accepted = False
    proposed = None
    step_info = dict(num_f_evaluations=0, num_df_evaluations=0, num_df_diagonal_evaluations=0,
                     num_attempted_steps=0)
    trace = []
    while not accepted:
        proposed, attempt_info = attempt(state, dt)
        step_info["num_attempted_steps"] += 1
        for key in ("num_f_evaluations", "num_df_evaluations", "num_df_diagonal_evaluations"):
            if key in attempt_info:
                step_info[key] += attempt_info[key]
        internal_norm = steprule.scale_error_estimate(
            unscaled_error_estimate=dt * proposed.error_estimate
            if proposed.error_estimate is not None
            else None,
            reference_state=proposed.reference_state,
        )
        accepted = steprule.is_accepted(internal_norm)
        suggested_dt = steprule.suggest(dt, internal_norm, local_convergence_rate=nu + 1)
        trace.append((state.t, dt, internal_norm, bool(accepted)))
        if accepted:
            dt = min(suggested_dt, tmax - proposed.t)
        else:
            dt = min(suggested_dt, tmax - state.t)
        assert dt >= 0, f"Invalid step size: dt={dt}"
    return proposed, dt, step_info, trace


def solution_generator(attempt, initial_state, steprule, nu, problem, stop_at=None):
    """tornadox/odefilter.py:80-159 (without the progress bar).  Yields (state, info, trace)."""
    time_stopper = TimeStopper(stop_at) if stop_at is not None else None
    state = initial_state
    info = dict(num_f_evaluations=0, num_df_evaluations=0, num_df_diagonal_evaluations=0,
                num_steps=0, num_attempted_steps=0)
    yield state, info, []
    dt = steprule.first_dt(*problem)
    while state.t < problem.tmax:
        if time_stopper is not None:
            dt = time_stopper.adjust_dt_to_time_stops(state.t, dt)
        state, dt, step_info, trace = perform_full_step(attempt, steprule, nu, state, dt, problem.tmax)
        info["num_steps"] += 1
        for key in ("num_f_evaluations", "num_df_evaluations", "num_df_diagonal_evaluations",
                    "num_attempted_steps"):
            info[key] += step_info[key]
        yield state, info, trace


def solve_kronecker_ek0(problem, nu, steprule, mean0, cov_sqrtm0, stop_at=None):
    """Drain the generator with the KroneckerEK0 attempt; returns (states, info, trace)."""
    attempt = lambda s, dt: filters.kronecker_ek0_attempt(s, dt, problem.f, nu)
    init = filters.kronecker_ek0_initial_state(mean0, cov_sqrtm0, problem.t0)
    return _drain(solution_generator(attempt, init, steprule, nu, problem, stop_at))


def solve_diagonal_ek1(problem, nu, steprule, mean0, cov_sqrtm0, stop_at=None):
    """Drain the generator with the DiagonalEK1 attempt; returns (states, info, trace)."""
    attempt = lambda s, dt: filters.diagonal_ek1_attempt(s, dt, problem.f, problem.df_diagonal, nu)
    init = filters.batched_initial_state(mean0, cov_sqrtm0, problem.t0)
    return _drain(solution_generator(attempt, init, steprule, nu, problem, stop_at))


def _drain(gen):
    states, full_trace, info = [], [], None
    for state, info, trace in gen:
        states.append(state)
        full_trace.extend(trace)
    return states, dict(info), full_trace


__all__ = [name for name in dir() if not name.startswith("_")] + ["prior"]
