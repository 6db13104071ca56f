"""Initial value problems whose right-hand sides are evaluated INSIDE the fused CUDA step kernels.

Mirrors the factory functions of the reference (tornadox/ivp.py): ``vanderpol`` (:28-49), ``vanderpol_julia`` (:52-73),
``burgers_1d`` (:76-107), ``wave_1d`` (:110-140), ``threebody`` (:188-216), ``brusselator`` (:219-265), ``lorenz96`` (:268-294),
``lorenz96_loop`` (:298-332), ``pleiades`` (:344-406), ``fhn_2d`` (:409-481); ``wave_2d`` (:143-185) does not run upstream
(``jax.ops.index_update`` is gone) and is not provided.  Each returns an
``InitialValueProblem(f, t0, tmax, y0, df, df_diagonal)`` (:10-25); ``f`` and ``df_diagonal`` are callables
on device tensors (they launch ``tdx_eval_f``) that additionally carry the ``KernelSpec`` the solvers
dispatch on.  ``df`` (dense Jacobian) is ``None``: the O(d^2) dense solvers are out of scope.
A user-supplied ``f`` (any callable on device tensors, optionally with ``df_diagonal``) is integrated too: the solvers then
evaluate it with the caller's own torch code at the predicted state of every attempt and run the same fused kernels on the
result (``TDX_IVP_EXTERNAL``); only the device-resident step controller needs an in-kernel vector field.
"""

from collections import namedtuple

import numpy as np
import torch

from . import _lib
from .engine import KernelSpec, StepEngine


class InitialValueProblem(
    namedtuple("_InitialValueProblem", "f t0 tmax y0 df df_diagonal", defaults=(None, None))
):
    """Initial value problems (reference: ivp.py:10-25)."""

    @property
    def dimension(self):
        if np.isscalar(self.y0):
            return 1
        return self.y0.shape[0]

    @property
    def t_span(self):
        return self.t0, self.tmax

    @property
    def spec(self):
        return getattr(self.f, "spec", None)


class DeviceVectorField:
    """Callable ``(t, y) -> f(t, y)`` (or the Jacobian diagonal) evaluated by the CUDA library."""

    def __init__(self, spec, jacobian_diagonal=False):
        self.spec = spec
        self.jacobian_diagonal = jacobian_diagonal
        self._engines = {}

    def _engine(self, y):
        key = (y.dtype, y.device)
        if key not in self._engines:
            self._engines[key] = StepEngine(self.spec, nu=1, dtype=y.dtype, device=y.device)
        return self._engines[key]

    def __call__(self, t, y):
        y = as_device_tensor(y)
        eng = self._engine(y)
        if self.jacobian_diagonal:
            return eng.eval_f(t, y, want_jac=True)[1]
        return eng.eval_f(t, y)

    def __repr__(self):
        what = "df_diagonal" if self.jacobian_diagonal else "f"
        return f"<{what} of {self.spec.name} evaluated in-kernel>"


def as_device_tensor(y, dtype=None, device=None):
    """State vectors live on the GPU: NumPy input is copied there once; tensors stay where they are."""
    if isinstance(y, torch.Tensor):
        if device is not None:
            y = y.to(device)
        if dtype is None and y.dtype not in (torch.float32, torch.float64):
            dtype = torch.float64
        return y if dtype is None or y.dtype == dtype else y.to(dtype)
    arr = np.asarray(y, dtype=np.float64)
    return torch.as_tensor(arr, dtype=dtype or torch.float64, device=device or "cuda")


def _problem(spec, t0, tmax, y0):
    return InitialValueProblem(
        f=DeviceVectorField(spec),
        t0=t0,
        tmax=tmax,
        y0=y0,
        df=None,
        df_diagonal=DeviceVectorField(spec, jacobian_diagonal=True),
    )


def vanderpol(t0=0.0, tmax=30, y0=None, stiffness_constant=1e1):
    """Van der Pol oscillator, y' = [y2, mu (1 - y1^2) y2 - y1]  (ivp.py:28-49)."""
    y0 = np.array([2.0, 0.0]) if y0 is None else y0
    spec = KernelSpec(_lib.TDX_IVP_VANDERPOL, (float(stiffness_constant),), 1, 1, 2, "vanderpol")
    return _problem(spec, t0, tmax, y0)


def vanderpol_julia(t0=0.0, tmax=6.3, y0=None, stiffness_constant=1e1):
    """Van der Pol in the scaling of the Julia benchmarks, y' = [y2, mu ((1 - y1^2) y2 - y1)]  (ivp.py:52-73)."""
    y0 = np.array([2.0, 0.0]) if y0 is None else y0
    spec = KernelSpec(_lib.TDX_IVP_VANDERPOL, (float(stiffness_constant), 1.0), 1, 1, 2, "vanderpol_julia")
    return _problem(spec, t0, tmax, y0)


def threebody(tmax=17.0652165601579625588917206249):
    """Restricted three-body problem, Y = [y1, y2, y1', y2']; the Jacobian diagonal is defined as zero  (ivp.py:188-216)."""
    y0 = np.array([0.994, 0, 0, -2.00158510637908252240537862224])
    spec = KernelSpec(_lib.TDX_IVP_THREEBODY, (), 1, 4, 1, "threebody")
    return _problem(spec, 0.0, tmax, y0)


def pleiades(t0=0.0, tmax=3.0):
    """Seven-body Pleiades problem (PLEI of Hairer I), u = [x(7); y(7); x'(7); y'(7)]  (ivp.py:344-406)."""
    y0 = np.array([3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0, 3.0, -3.0, 2.0, 0, 0, -4.0, 4.0, 0, 0, 0, 0, 0, 1.75, -1.5, 0, 0, 0,
                   -1.25, 1, 0, 0])
    spec = KernelSpec(_lib.TDX_IVP_PLEIADES, (), 1, 28, 1, "pleiades")
    return _problem(spec, t0, tmax, y0)


def _mesh_1d(bbox, dx):
    """``arange(bbox[0], bbox[1] + dx, step=dx)`` as in ivp.py:81, 115."""
    if bbox is None:
        bbox = [0.0, 1.0]
    return np.arange(bbox[0], bbox[1] + dx, step=dx), bbox


def burgers_1d(t0=0.0, tmax=10.0, y0=None, bbox=None, dx=0.02, diffusion_param=0.02):
    """Viscous Burgers equation on a 1-d mesh, upwind convection, "cyclic" end points  (ivp.py:76-107)."""
    mesh, bbox = _mesh_1d(bbox, dx)
    if y0 is None:
        y0 = np.exp(-500.0 * (mesh - (0.6 * (bbox[1] - bbox[0]))) ** 2)
    n = int(np.shape(y0)[0])
    if n < 3:
        raise ValueError("burgers_1d needs at least 3 mesh points")
    spec = KernelSpec(_lib.TDX_IVP_BURGERS1D, (float(dx), float(diffusion_param)), 1, n, 1, "burgers_1d")
    return _problem(spec, t0, tmax, y0)


def wave_1d(t0=0.0, tmax=20.0, y0=None, bbox=None, dx=0.02, diffusion_param=0.01):
    """Wave equation as a first-order system y = [x; x'] on a 1-d mesh  (ivp.py:110-140)."""
    mesh, bbox = _mesh_1d(bbox, dx)
    if y0 is None:
        x0 = np.exp(-70.0 * (mesh - (0.6 * (bbox[1] - bbox[0]))) ** 2)
        y0 = np.concatenate((x0, np.zeros_like(x0)))
    n = int(np.shape(y0)[0]) // 2
    if n < 3 or 2 * n != int(np.shape(y0)[0]):
        raise ValueError("wave_1d needs y0 = [x; x'] on at least 3 mesh points")
    spec = KernelSpec(_lib.TDX_IVP_WAVE1D, (float(dx), float(diffusion_param)), 1, n, 2, "wave_1d")
    return _problem(spec, t0, tmax, y0)


def brusselator(N=20, t0=0.0, tmax=10.0, y0=None):
    """Brusselator chain, y = [u(N); v(N)], c = (N+1)^2/50, pads u=1, v=3  (ivp.py:219-265)."""
    if y0 is None:
        u0 = np.arange(1, N + 1) / N + 1
        v0 = 3.0 * np.ones(N)
        y0 = np.concatenate([u0, v0])
    spec = KernelSpec(_lib.TDX_IVP_BRUSSELATOR, (), 1, int(N), 2, "brusselator")
    return _problem(spec, t0, tmax, y0)


def lorenz96(t0=0.0, tmax=30.0, y0=None, num_variables=10, forcing=8.0):
    """Lorenz-96, cyclic (y[i+1] - y[i-2]) y[i-1] - y[i] + F  (ivp.py:268-294, default y0 :335-341)."""
    if y0 is None:
        y0 = np.ones(num_variables) * forcing
        y0[0] += 0.01
    else:
        num_variables = int(y0.shape[0])
    spec = KernelSpec(_lib.TDX_IVP_LORENZ96, (float(forcing),), 1, int(num_variables), 1, "lorenz96")
    return _problem(spec, t0, tmax, y0)


def lorenz96_loop(t0=0.0, tmax=30.0, y0=None, num_variables=10, forcing=8.0):
    """The reference's loop formulation of Lorenz-96 (ivp.py:298-332, written that way for jet): same vector field, same
    kernel."""
    return lorenz96(t0=t0, tmax=tmax, y0=y0, num_variables=num_variables, forcing=forcing)


def fhn_2d(t0=0.0, tmax=20.0, y0=None, bbox=None, dx=0.02, a=2.8e-4, b=5e-3, k=-0.005, tau=0.1, prng_key=None):
    """FitzHugh-Nagumo reaction-diffusion on a square grid with Neumann boundaries  (ivp.py:409-481).

    ``prng_key`` seeds the default y0 ~ U[0, 1).  The reference draws it with ``jax.random`` (threefry),
    which is not reproducible without JAX; here it is ``numpy.random.default_rng(prng_key or 2)``.
    """
    if bbox is None:
        bbox = [[0.0, 0.0], [1.0, 1.0]]
    ny, nx = int((bbox[1][0] - bbox[0][0]) / dx), int((bbox[1][1] - bbox[0][1]) / dx)
    if ny != nx:
        raise ValueError(
            f"The diagonal Jacobian only works for quadratic spatial grids. Got {bbox}."
        )
    if y0 is None:
        seed = 2 if prng_key is None else prng_key
        y0 = np.random.default_rng(seed).uniform(size=(2 * ny * nx,))
    spec = KernelSpec(_lib.TDX_IVP_FHN2D, (float(dx), float(a), float(b), float(k), float(tau)), nx, ny, 2, "fhn_2d")
    return _problem(spec, t0, tmax, y0)
