"""Shared test helpers: build matching (oracle problem, product ivp) pairs and compare states."""

import numpy as np

from oracle.ref_np import driver as odriver
from oracle.ref_np import problems as oproblems

FP64_RTOL = 1e-10   # north_star: means and Cholesky factors within 1e-10 (fp64) / 1e-5 (fp32)
FP32_RTOL = 1e-5


def make_pair(kind, size, **kw):
    """Return (oracle Problem, product InitialValueProblem, series_f) for a vector field at a given size."""
    from tornadox_b200 import ivp as pivp

    if kind == "vanderpol":
        mu = kw.get("mu", 1.0)
        o = oproblems.vanderpol(t0=0.0, tmax=kw.get("tmax", 1.0), stiffness_constant=mu)
        p = pivp.vanderpol(t0=0.0, tmax=kw.get("tmax", 1.0), stiffness_constant=mu)
        return o, p, odriver.series_vanderpol(mu)
    if kind == "vanderpol_julia":
        mu = kw.get("mu", 2.0)
        return (oproblems.vanderpol_julia(t0=0.0, tmax=0.5, stiffness_constant=mu),
                pivp.vanderpol_julia(t0=0.0, tmax=0.5, stiffness_constant=mu), odriver.series_vanderpol_julia(mu))
    if kind == "threebody":
        return oproblems.threebody(tmax=0.02), pivp.threebody(tmax=0.02), odriver.series_threebody()
    if kind == "pleiades":
        return oproblems.pleiades(tmax=0.2), pivp.pleiades(tmax=0.2), odriver.series_pleiades()
    if kind == "lorenz96":
        o = oproblems.lorenz96(num_variables=size)
        p = pivp.lorenz96(num_variables=size)
        return o, p, odriver.series_lorenz96(8.0)
    if kind == "brusselator":
        o = oproblems.brusselator(N=size)
        p = pivp.brusselator(N=size)
        return o, p, odriver.series_brusselator(size)
    if kind == "burgers_1d":                      # size = number of mesh points
        dx = 1.0 / (size - 1)
        o = oproblems.burgers_1d(dx=dx)
        assert o.y0.shape[0] == size, (o.y0.shape, size)
        return o, pivp.burgers_1d(dx=dx), odriver.series_burgers_1d(dx)
    if kind == "wave_1d":
        dx = 1.0 / (size - 1)
        o = oproblems.wave_1d(dx=dx)
        assert o.y0.shape[0] == 2 * size, (o.y0.shape, size)
        return o, pivp.wave_1d(dx=dx), odriver.series_wave_1d(dx)
    if kind == "fhn_2d":
        dx = 1.0 / size
        y0 = np.random.default_rng(2).uniform(size=(2 * size * size,))
        o = oproblems.fhn_2d(dx=dx, y0=y0)
        p = pivp.fhn_2d(dx=dx, y0=y0)
        nx, ny = oproblems.fhn_grid_size(None, dx)
        assert nx == size, (nx, size)
        return o, p, odriver.series_fhn_2d(nx, ny, dx)
    raise KeyError(kind)


def random_state(rng, n, d, cov="dense", mean_scale=1.0):
    """A mid-solve looking state: O(1) mean rows and small factor blocks."""
    mean = mean_scale * rng.standard_normal((n, d))
    if cov == "zero":
        sc = np.zeros((d, n, n))
    elif cov == "dense":
        sc = 1e-2 * rng.standard_normal((d, n, n))
    elif cov == "lower":
        sc = np.tril(1e-2 * rng.standard_normal((d, n, n)))
    else:
        raise KeyError(cov)
    return mean, sc


def rel_err(a, b):
    """max |a - b| relative to the max-norm of the reference array b."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = np.max(np.abs(b)) if b.size else 0.0
    diff = np.max(np.abs(a - b)) if b.size else 0.0
    return diff / scale if scale > 0 else diff


NOISE_FACTOR = 50.0


def assert_close(a, b, rtol, what, noise=0.0):
    """|a - b| <= rtol * max|b| + NOISE_FACTOR * noise.

    ``noise`` is the change the REFERENCE's own output shows when its inputs are perturbed by ~1 ulp
    (recorded next to each golden step by oracle/make_golden.py).  It is non-negligible only where the
    algorithm itself is ill-conditioned: right after an exact Taylor-mode initialisation the residual
    z = p1 mp[1] - f(p0 mp[0]) cancels 8+ digits, and error_estimate = |z| and the factors (proportional to
    |z|) carry that rounding noise in ANY fp64 implementation, including JAX/XLA vs NumPy.
    """
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = np.max(np.abs(b)) if b.size else 0.0
    diff = np.max(np.abs(a - b)) if b.size else 0.0
    bound = rtol * scale + NOISE_FACTOR * float(noise)
    if scale == 0 and noise == 0:
        bound = rtol
    assert diff <= bound, f"{what}: |diff| {diff:.3e} > {bound:.3e} (rtol {rtol:.1e} * {scale:.3e} + {NOISE_FACTOR} * noise {float(noise):.3e})"


# ---- golden vectors generated from the reference's own sources (oracle/make_golden.py) ----------------
import functools
import pathlib

GOLDEN_PATH = pathlib.Path(__file__).resolve().parent / "golden" / "reference_vectors.npz"


@functools.lru_cache(maxsize=1)
def golden():
    out = {}
    # disjoint keys: different problems / (oracle/make_golden_ek0_const.py) ConstantSteps solves of the reference's KroneckerEK0
    for path in (GOLDEN_PATH, GOLDEN_PATH.with_name("reference_vectors_small.npz"), GOLDEN_PATH.with_name("reference_vectors_ek0_const.npz")):
        with np.load(path) as data:
            out.update({k: data[k] for k in data.files})
    return out


def oracle_problem(name):
    """The oracle problem + series vector field matching the golden problem ``name`` (oracle/make_golden.py)."""
    g = golden()
    y0 = g[f"ivp/{name}/y0"]
    if name == "vanderpol":
        return oproblems.vanderpol(t0=0.0, tmax=1.0, stiffness_constant=1.0), odriver.series_vanderpol(1.0)
    if name == "lorenz96":
        return oproblems.lorenz96(t0=0.0, tmax=0.5, num_variables=10, forcing=8.0), odriver.series_lorenz96(8.0)
    if name == "brusselator":
        return oproblems.brusselator(N=7, t0=0.0, tmax=0.05), odriver.series_brusselator(7)
    if name == "fhn_2d":
        return oproblems.fhn_2d(t0=0.0, tmax=0.05, y0=y0, dx=0.2), odriver.series_fhn_2d(5, 5, 0.2)
    if name == "burgers_1d":
        return oproblems.burgers_1d(t0=0.0, tmax=0.05, dx=0.1), odriver.series_burgers_1d(0.1)
    if name == "wave_1d":
        return oproblems.wave_1d(t0=0.0, tmax=0.05, dx=0.1), odriver.series_wave_1d(0.1)
    if name == "vanderpol_julia":
        return oproblems.vanderpol_julia(t0=0.0, tmax=0.5, stiffness_constant=2.0), odriver.series_vanderpol_julia(2.0)
    if name == "threebody":
        return oproblems.threebody(tmax=0.02), odriver.series_threebody()
    if name == "pleiades":
        return oproblems.pleiades(t0=0.0, tmax=0.2), odriver.series_pleiades()
    if name == "lorenz96_loop":
        return oproblems.lorenz96_loop(t0=0.0, tmax=0.3, num_variables=6, forcing=8.0), odriver.series_lorenz96(8.0)
    raise KeyError(name)


def product_problem(name):
    """The tornadox_b200 ivp matching the golden problem ``name``."""
    from tornadox_b200 import ivp as pivp

    y0 = golden()[f"ivp/{name}/y0"]
    if name == "vanderpol":
        return pivp.vanderpol(t0=0.0, tmax=1.0, stiffness_constant=1.0)
    if name == "lorenz96":
        return pivp.lorenz96(t0=0.0, tmax=0.5, num_variables=10, forcing=8.0)
    if name == "brusselator":
        return pivp.brusselator(N=7, t0=0.0, tmax=0.05)
    if name == "fhn_2d":
        return pivp.fhn_2d(t0=0.0, tmax=0.05, y0=y0, dx=0.2)
    if name == "burgers_1d":
        return pivp.burgers_1d(t0=0.0, tmax=0.05, dx=0.1)
    if name == "wave_1d":
        return pivp.wave_1d(t0=0.0, tmax=0.05, dx=0.1)
    if name == "vanderpol_julia":
        return pivp.vanderpol_julia(t0=0.0, tmax=0.5, stiffness_constant=2.0)
    if name == "threebody":
        return pivp.threebody(tmax=0.02)
    if name == "pleiades":
        return pivp.pleiades(t0=0.0, tmax=0.2)
    if name == "lorenz96_loop":
        return pivp.lorenz96_loop(t0=0.0, tmax=0.3, num_variables=6, forcing=8.0)
    raise KeyError(name)
