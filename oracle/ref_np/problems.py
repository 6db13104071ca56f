"""ORACLE (test infrastructure): NumPy restatement of the in-scope vector fields of ``tornadox/ivp.py``.

vanderpol (ivp.py:28-49), vanderpol_julia (ivp.py:52-73), burgers_1d (ivp.py:76-107), wave_1d (ivp.py:110-140),
threebody (ivp.py:188-216), brusselator (ivp.py:219-265), lorenz96 (ivp.py:268-294, 335-341), lorenz96_loop (ivp.py:298-332),
pleiades (ivp.py:344-406), fhn_2d (ivp.py:409-481) with its 5-point Neumann Laplacian (ivp.py:484-505).
Each factory returns ``Problem(f, t0, tmax, y0, df, df_diagonal)`` -- the field order of the
reference's ``InitialValueProblem`` namedtuple (ivp.py:10-25).  ``df`` (dense Jacobian, the
reference gets it from ``jax.jacfwd``) is built analytically and is only meant for small d.
"""

from collections import namedtuple

import numpy as np
import scipy.signal

Problem = namedtuple("Problem", "f t0 tmax y0 df df_diagonal", defaults=(None, None))


def vanderpol(t0=0.0, tmax=30, y0=None, stiffness_constant=1e1):
    """ivp.py:28-49."""
    mu = stiffness_constant
    y0 = np.array([2.0, 0.0]) if y0 is None else np.asarray(y0, dtype=float)

    def f(_, Y):
        return np.array([Y[1], mu * (1.0 - Y[0] ** 2) * Y[1] - Y[0]])

    def df(_, Y):
        return np.array([[0.0, 1.0], [-2.0 * mu * Y[0] * Y[1] - 1.0, mu * (1.0 - Y[0] ** 2)]])

    def df_diagonal(_, Y):
        return np.array([0.0, mu * (1.0 - Y[0] ** 2)])

    return Problem(f, t0, tmax, y0, df, df_diagonal)


def vanderpol_julia(t0=0.0, tmax=6.3, y0=None, stiffness_constant=1e1):
    """ivp.py:52-73: mu scales the whole second component."""
    mu = stiffness_constant
    y0 = np.array([2.0, 0.0]) if y0 is None else np.asarray(y0, dtype=float)

    def f(_, Y):
        return np.array([Y[1], mu * ((1.0 - Y[0] ** 2) * Y[1] - Y[0])])

    def df(_, Y):
        return np.array([[0.0, 1.0], [mu * (-2.0 * Y[0] * Y[1] - 1.0), mu * (1.0 - Y[0] ** 2)]])

    def df_diagonal(_, Y):
        return np.array([0.0, mu * (1.0 - Y[0] ** 2)])

    return Problem(f, t0, tmax, y0, df, df_diagonal)


def threebody(tmax=17.0652165601579625588917206249):
    """ivp.py:188-216 (restricted three-body problem, Y = [y1, y2, y1', y2']); the reference DEFINES its Jacobian diagonal
    as zero (:204-205)."""
    mu = 0.012277471
    mp = 1 - mu

    def f(_, Y):
        D1 = ((Y[0] + mu) ** 2 + Y[1] ** 2) ** (3 / 2)
        D2 = ((Y[0] - mp) ** 2 + Y[1] ** 2) ** (3 / 2)
        y1p = Y[0] + 2 * Y[3] - mp * (Y[0] + mu) / D1 - mu * (Y[0] - mp) / D2
        y2p = Y[1] - 2 * Y[2] - mp * Y[1] / D1 - mu * Y[1] / D2
        return np.array([Y[2], Y[3], y1p, y2p])

    def df(t, Y):
        return _dense_jacobian(f, t, np.asarray(Y, dtype=float))

    def df_diagonal(*args, **kwargs):
        return np.array([0.0, 0.0, 0.0, 0.0])

    y0 = np.array([0.994, 0, 0, -2.00158510637908252240537862224])
    return Problem(f, 0.0, tmax, y0, df, df_diagonal)


PLEIADES_Y0 = np.array([3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0, 3.0, -3.0, 2.0, 0, 0, -4.0, 4.0, 0, 0, 0, 0, 0, 1.75, -1.5, 0, 0, 0,
                        -1.25, 1, 0, 0])


def pleiades(t0=0.0, tmax=3.0):
    """ivp.py:344-406 (PLEI of Hairer I): u = [x(7); y(7); x'(7); y'(7)], body j has mass j.  The 0/0 of a body with itself
    is mapped to 0 (``nan_to_num``, :384-385).  The Jacobian diagonal (``diagonal(jacfwd(f))``, :393-397) vanishes: positions'
    derivatives are velocities, accelerations depend on positions only."""

    def f(t, u):
        x, y, dx, dy = u[0:7], u[7:14], u[14:21], u[21:28]
        xi, xj = x[:, None], x[None, :]
        yi, yj = y[:, None], y[None, :]
        rij = ((xi - xj) ** 2 + (yi - yj) ** 2) ** (3 / 2)
        mj = np.arange(1, 8)[None, :]
        with np.errstate(invalid="ignore", divide="ignore"):
            ddx = np.sum(np.nan_to_num(mj * (xj - xi) / rij), axis=1)
            ddy = np.sum(np.nan_to_num(mj * (yj - yi) / rij), axis=1)
        return np.concatenate((dx, dy, ddx, ddy))

    def df(t, u):
        # complex step with the self-interaction masked out (nan_to_num is not analytic)
        u = np.asarray(u, dtype=float)
        J = np.zeros((28, 28))
        off = ~np.eye(7, dtype=bool)
        for k in range(28):
            uc = u.astype(complex)
            uc[k] += 1e-30j
            x, y = uc[0:7], uc[7:14]
            ddx, ddy = np.zeros(7, complex), np.zeros(7, complex)
            for i in range(7):
                for j in range(7):
                    if off[i, j]:
                        r3 = ((x[i] - x[j]) ** 2 + (y[i] - y[j]) ** 2) ** 1.5
                        ddx[i] += (j + 1) * (x[j] - x[i]) / r3
                        ddy[i] += (j + 1) * (y[j] - y[i]) / r3
            J[:, k] = np.imag(np.concatenate((uc[14:21], uc[21:28], ddx, ddy))) / 1e-30
        return J

    def df_diagonal(t, u):
        return np.zeros(28)

    return Problem(f, t0, tmax, PLEIADES_Y0.copy(), df, df_diagonal)


def mesh_1d(bbox, dx):
    """ivp.py:81, 115: ``arange(bbox[0], bbox[1] + dx, step=dx)`` (NumPy's arange is what jnp.arange evaluates)."""
    if bbox is None:
        bbox = [0.0, 1.0]
    return np.arange(bbox[0], bbox[1] + dx, step=dx), bbox


def burgers_1d(t0=0.0, tmax=10.0, y0=None, bbox=None, dx=0.02, diffusion_param=0.02):
    """ivp.py:76-107.  Upwind convection + diffusion in the interior; BOTH end points get the same "cyclic" value built
    from x[0], x[1] and x[-2].  The reference's Jacobian diagonal is ``diag(jacfwd(f))`` (:106): written out here."""
    mesh, bbox = mesh_1d(bbox, dx)
    if y0 is None:
        y0 = np.exp(-500.0 * (mesh - (0.6 * (bbox[1] - bbox[0]))) ** 2)

    def f(_, x):
        nonlinear_convection = x[1:-1] * (x[1:-1] - x[:-2]) / dx
        diffusion = diffusion_param * (x[2:] - 2.0 * x[1:-1] + x[:-2]) / (dx**2)
        interior = (-nonlinear_convection + diffusion).reshape(-1)
        boundary = np.array(-x[0] * (x[0] - x[-2]) / dx + diffusion_param * (x[1] - 2.0 * x[0] + x[-2]) / (dx**2)).reshape(-1)
        return np.concatenate((boundary, interior, boundary))

    def df_diagonal(_, x):
        d = np.empty_like(x)
        d[1:-1] = -(2.0 * x[1:-1] - x[:-2]) / dx - 2.0 * diffusion_param / (dx**2)
        d[0] = -(2.0 * x[0] - x[-2]) / dx - 2.0 * diffusion_param / (dx**2)
        d[-1] = 0.0                      # f[-1] is the boundary value, which does not depend on x[-1]
        return d

    def df(t, x):
        return _dense_jacobian(f, t, x)

    return Problem(f, t0, tmax, np.asarray(y0, dtype=float), df, df_diagonal)


def wave_1d(t0=0.0, tmax=20.0, y0=None, bbox=None, dx=0.02, diffusion_param=0.01):
    """ivp.py:110-140.  y = [x(N); x'(N)]; the end points are frozen in a peculiar way: d/dt x = 0 and d/dt x' = x there.
    Every diagonal entry of the Jacobian is zero."""
    mesh, bbox = mesh_1d(bbox, dx)
    if y0 is None:
        x0 = np.exp(-70.0 * (mesh - (0.6 * (bbox[1] - bbox[0]))) ** 2)
        y0 = np.concatenate((x0, np.zeros_like(x0)))

    def f(_, x):
        _x, _dx = np.split(x, 2)
        interior = diffusion_param * (_x[2:] - 2.0 * _x[1:-1] + _x[:-2]) / (dx**2)
        _ddx = np.concatenate((_x[:1], interior, _x[-1:]))
        new_dx = np.concatenate((np.zeros(1), _dx[1:-1], np.zeros(1)))
        return np.concatenate((new_dx, _ddx))

    def df_diagonal(_, x):
        return np.zeros_like(x)

    def df(t, x):
        return _dense_jacobian(f, t, x)

    return Problem(f, t0, tmax, np.asarray(y0, dtype=float), df, df_diagonal)


def _dense_jacobian(f, t, x, h=1e-30):
    """Complex-step Jacobian (exact to rounding for the polynomial fields above; small d only)."""
    n = x.shape[0]
    J = np.zeros((n, n))
    for j in range(n):
        xc = x.astype(complex)
        xc[j] += 1j * h
        J[:, j] = np.imag(f(t, xc)) / h
    return J


def brusselator(N=20, t0=0.0, tmax=10.0, y0=None):
    """ivp.py:219-265.  y = [u(N); v(N)], c = (N+1)^2/50, boundary pads u=1, v=3."""
    alpha = 1.0 / 50.0
    const = alpha * (N + 1) ** 2
    weights = np.array([1.0, -2.0, 1.0])

    def f(_, y, n=N, w=weights, c=const):
        u, v = y[:n], y[n:]
        u_ = np.concatenate([[1.0], u, [1.0]])
        v_ = np.concatenate([[3.0], v, [3.0]])
        conv_u = np.convolve(u_, w, mode="valid")
        conv_v = np.convolve(v_, w, mode="valid")
        u_new = 1.0 + u**2 * v - 4 * u + c * conv_u
        v_new = 3 * u - u**2 * v + c * conv_v
        return np.concatenate([u_new, v_new])

    def df_diagonal(_, y, n=N, c=const):
        u, v = y[:n], y[n:]
        return np.concatenate([2 * u * v - 4 - 2 * c, -(u**2) - 2 * c])

    def df(_, y, n=N, c=const):
        u, v = y[:n], y[n:]
        lap = c * (np.diag(-2.0 * np.ones(n)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1))
        J = np.zeros((2 * n, 2 * n))
        J[:n, :n] = lap + np.diag(2 * u * v - 4)
        J[:n, n:] = np.diag(u**2)
        J[n:, :n] = np.diag(3 - 2 * u * v)
        J[n:, n:] = lap + np.diag(-(u**2))
        return J

    if y0 is None:
        u0 = np.arange(1, N + 1) / N + 1
        v0 = 3.0 * np.ones(N)
        y0 = np.concatenate([u0, v0])
    return Problem(f, t0, tmax, np.asarray(y0, dtype=float), df, df_diagonal)


def lorenz96(t0=0.0, tmax=30.0, y0=None, num_variables=10, forcing=8.0):
    """ivp.py:268-294 with the default y0 of ivp.py:335-341."""
    if y0 is None:
        y0 = np.ones(num_variables) * forcing
        y0[0] += 0.01

    def f(_, y, c=forcing):
        A = np.roll(y, shift=-1)
        B = np.roll(y, shift=2)
        C = np.roll(y, shift=1)
        D = y
        return (A - B) * C - D + c

    def df(_, y, c=forcing):
        n = y.shape[0]
        J = -np.eye(n)
        idx = np.arange(n)
        J[idx, (idx + 1) % n] += y[(idx - 1) % n]
        J[idx, (idx - 2) % n] += -y[(idx - 1) % n]
        J[idx, (idx - 1) % n] += y[(idx + 1) % n] - y[(idx - 2) % n]
        return J

    def df_diagonal(_, y, c=forcing):
        return -1.0 * np.ones(y.shape[0])

    return Problem(f, t0, tmax, np.asarray(y0, dtype=float), df, df_diagonal)


def lorenz96_loop(t0=0.0, tmax=30.0, y0=None, num_variables=10, forcing=8.0):
    """ivp.py:298-332: the same vector field written as a loop over the coordinates (identical arithmetic per entry)."""
    return lorenz96(t0=t0, tmax=tmax, y0=y0, num_variables=num_variables, forcing=forcing)


def laplace_2d(grid, dx):
    """5-point Laplacian with edge replication (Neumann) -- ivp.py:484-505."""
    padded = np.pad(grid, pad_width=1, mode="edge")
    kernel = 1 / (dx**2) * np.array([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]])
    out = scipy.signal.convolve2d(padded, kernel, mode="same")
    return out[1:-1, 1:-1]


def fhn_grid_size(bbox, dx):
    """ivp.py:424-430 -- ny, nx = int(extent/dx); must be square."""
    if bbox is None:
        bbox = [[0.0, 0.0], [1.0, 1.0]]
    ny, nx = int((bbox[1][0] - bbox[0][0]) / dx), int((bbox[1][1] - bbox[0][1]) / dx)
    if ny != nx:
        raise ValueError(
            f"The diagonal Jacobian only works for quadratic spatial grids. Got {bbox}."
        )
    return nx, ny


def fhn_2d(t0=0.0, tmax=20.0, y0=None, bbox=None, dx=0.02, a=2.8e-4, b=5e-3, k=-0.005, tau=0.1, seed=2):
    """ivp.py:409-481.

    The reference's default y0 is ``jax.random.uniform(PRNGKey(2))`` (ivp.py:432-434) which cannot
    be reproduced without JAX; here the default is ``numpy.random.default_rng(seed).uniform`` and
    parity tests always pass y0 explicitly.
    """
    nx, ny = fhn_grid_size(bbox, dx)
    if y0 is None:
        y0 = np.random.default_rng(seed).uniform(size=(2 * ny * nx,))

    def f(_, x):
        u, v = np.split(x, 2)
        du = laplace_2d(u.reshape((nx, ny)), dx=dx).reshape((-1,))
        dv = laplace_2d(v.reshape((nx, ny)), dx=dx).reshape((-1,))
        u_new = a * du + u - u**3 - v + k
        v_new = (b * dv + u - v) / tau
        return np.concatenate((u_new, v_new))

    # ivp.py:447-463 -- Laplacian diagonal: 2 (corner), 3 (edge), 4 (interior), times -1/dx^2
    diag = 3.0 * np.ones((nx, ny))
    diag[1:-1, 1:-1] = 4.0
    diag[0, 0] = diag[0, -1] = diag[-1, 0] = diag[-1, -1] = 2.0
    dlaplace = -1.0 * diag.reshape((-1,)) / (dx**2)

    def df_diagonal(_, x):
        u, v = np.split(x, 2)
        d_u = a * dlaplace + 1.0 - 3.0 * u**2
        d_v = (b * dlaplace - 1.0) / tau
        return np.concatenate((d_u, d_v))

    def df(_, x):
        """Dense Jacobian by exact differentiation of ``f`` (column-wise; small grids only)."""
        m = x.shape[0]
        u, _v = np.split(x, 2)
        J = np.zeros((m, m))
        for j in range(m):
            e = np.zeros(m)
            e[j] = 1.0
            zero = np.zeros_like(x)
            lin = f(None, e) - f(None, zero)  # exact: e**3 contributes -1 on the u-u diagonal only
            J[:, j] = lin
        half = m // 2
        idx = np.arange(half)
        J[idx, idx] += 1.0 - 3.0 * u**2  # undo the -e**3 = -1 on the diagonal, add -3u^2
        return J

    return Problem(f, t0, tmax, np.asarray(y0, dtype=float), df, df_diagonal)
