"""GPU, BASELINE.json's FULL sizes: the CUDA path against the NumPy oracle where the oracle finishes in seconds, and
through size-independent properties elsewhere:

* DiagonalEK1 is per-coordinate apart from f's stencil, so a WINDOW of the state (a few grid rows / chain points plus one
  halo row / point on each side) run through the oracle must reproduce the kernel's output on the window's interior;
* KroneckerEK0 couples all coordinates through one scalar, and its oracle is a handful of d-wide NumPy operations, so it
  is compared in full;
* chunk invariance: the host-buffer pipeline (16 independent sub-shards with halos) must equal the one-launch result
  bit for bit.
"""

import math

import numpy as np
import pytest
import torch

from helpers import FP64_RTOL, assert_close
from oracle.ref_np import driver as odriver
from oracle.ref_np import filters as ofilters
from oracle.ref_np import problems as oproblems

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gpu_state(n, d, seed, with_blocks=True):
    g = torch.Generator(device=DEV).manual_seed(seed)
    mean = torch.randn((n, d), generator=g, device=DEV, dtype=torch.float64)
    mean[0] = torch.rand(d, generator=g, device=DEV, dtype=torch.float64)
    sc = 1e-2 * torch.randn((d, n, n), generator=g, device=DEV, dtype=torch.float64) if with_blocks else None
    return mean, sc


def _fhn_window_problem(side, dx, r0, r1, a=2.8e-4, b=5e-3, k=-0.005, tau=0.1):
    """f and df_diagonal of fhn_2d restricted to grid rows [r0, r1) (all columns): the oracle's own Laplacian
    (edge replication) on the sub-grid -- exact on rows that have both neighbours inside the window or sit on the true grid
    edge -- and ivp.py:447-463's Jacobian diagonal with GLOBAL row indices."""
    rows = r1 - r0

    def f(_, x):
        u, v = np.split(x, 2)
        du = oproblems.laplace_2d(u.reshape(rows, side), dx=dx).reshape(-1)
        dv = oproblems.laplace_2d(v.reshape(rows, side), dx=dx).reshape(-1)
        return np.concatenate((a * du + u - u**3 - v + k, (b * dv + u - v) / tau))

    gr = np.arange(r0, r1)[:, None] * np.ones((1, side), dtype=int)
    gc = np.arange(side)[None, :] * np.ones((rows, 1), dtype=int)
    nrep = (gr == 0).astype(int) + (gr == side - 1) + (gc == 0) + (gc == side - 1)
    dlaplace = (-(4.0 - nrep) / dx**2).reshape(-1)

    def df_diagonal(_, x):
        u, _v = np.split(x, 2)
        return np.concatenate((a * dlaplace + 1.0 - 3.0 * u**2, (b * dlaplace - 1.0) / tau * np.ones_like(u)))

    return f, df_diagonal


def test_config5_dek1_fhn2048_windows_and_chunk_invariance(cuda_device):
    """BASELINE config 5 (fhn_2d 2048 x 2048, d = 8 388 608, DiagonalEK1 nu = 4)."""
    from tornadox_b200 import ivp
    from tornadox_b200.engine import StepEngine

    side, nu = 2048, 4
    n, d, npts = nu + 1, 2 * side * side, side * side
    dx = 1.0 / side
    p = ivp.fhn_2d(dx=dx, y0=np.zeros(2))
    eng = StepEngine(p.spec, nu, device=DEV)
    mean, sc = _gpu_state(n, d, 11)
    dt, abstol, reltol = 2e-7, 1e-4, 1e-2
    m, s, err, ref, sq = eng.dek1_attempt(0.3, dt, mean, sc, abstol, reltol)
    torch.cuda.synchronize()

    total_ratio_sq = 0.0
    for r0, r1 in ((0, 7), (1021, 1031), (side - 6, side)):       # top edge, a tile-row boundary, bottom edge
        rows = r1 - r0
        idx = torch.cat([torch.arange(f * npts + r0 * side, f * npts + r1 * side, device=DEV) for f in range(2)])
        wm, wsc = mean[:, idx].cpu().numpy(), sc[idx].cpu().numpy()
        f, dfd = _fhn_window_problem(side, dx, r0, r1)
        ost, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, wm, wsc, None, None), dt, f, dfd, nu)
        lo = 0 if r0 == 0 else 1                                   # drop the halo rows of interior windows
        hi = rows if r1 == side else rows - 1
        keep = np.concatenate([np.arange(fl * rows * side + lo * side, fl * rows * side + hi * side) for fl in range(2)])
        sel = idx.cpu().numpy()[keep]
        assert_close(m[:, sel].cpu().numpy(), ost.mean[:, keep], FP64_RTOL, f"mean rows {r0}:{r1}")
        assert_close(s[sel].cpu().numpy(), ost.cov_sqrtm[keep], FP64_RTOL, f"cov_sqrtm rows {r0}:{r1}")
        assert_close(err[sel].cpu().numpy(), ost.error_estimate[keep], FP64_RTOL, f"error_estimate rows {r0}:{r1}")
        assert_close(ref[sel].cpu().numpy(), ost.reference_state[keep], FP64_RTOL, f"reference_state rows {r0}:{r1}")
        total_ratio_sq += float(np.sum((dt * ost.error_estimate[keep] / (abstol + reltol * ost.reference_state[keep])) ** 2))
    # the fused norm is the sum over ALL coordinates of what the windows verified element-wise
    ratio = dt * err / (abstol + reltol * ref)
    assert abs(float(sq.sum()) - float((ratio * ratio).sum())) <= 1e-11 * float(sq.sum())
    assert total_ratio_sq > 0.0

    # chunk invariance at full size: 16 sub-shards with exchanged halos == one launch, bit for bit
    h_mean = torch.empty(mean.shape, dtype=mean.dtype, pin_memory=True).copy_(mean)
    h_sc = torch.empty(sc.shape, dtype=sc.dtype, pin_memory=True).copy_(sc)
    hm, hs, herr, href, hsq = eng.dek1_attempt_host(0.3, dt, h_mean, h_sc, abstol, reltol, nchunks=16)
    assert torch.equal(hm, m.cpu()) and torch.equal(hs, s.cpu())
    assert torch.equal(herr, err.cpu()) and torch.equal(href, ref.cpu())
    assert abs(hsq - float(sq.sum())) <= 1e-12 * hsq


def test_config5_ek0_fhn2048_full_oracle(cuda_device):
    """BASELINE config 5, KroneckerEK0 nu = 4: the whole 8.4M-dimensional attempt against the oracle."""
    from tornadox_b200 import ivp
    from tornadox_b200.engine import StepEngine

    side, nu = 2048, 4
    n, d = nu + 1, 2 * side * side
    dx = 1.0 / side
    o = oproblems.fhn_2d(dx=dx, y0=np.zeros(d))
    p = ivp.fhn_2d(dx=dx, y0=np.zeros(2))
    eng = StepEngine(p.spec, nu, device=DEV)
    mean, _ = _gpu_state(n, d, 12, with_blocks=False)
    Cl = 1e-2 * np.random.default_rng(3).standard_normal((n, n))
    dt, abstol, reltol = 2e-7, 1e-4, 1e-2
    m, C, err, ref, itol = eng.ek0_attempt(0.3, dt, mean, torch.as_tensor(Cl, device=DEV), abstol, reltol)
    ost, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(0.3, mean.cpu().numpy(), Cl, None, None), dt, o.f, nu)
    ref_norm = odriver.AdaptiveSteps(abstol, reltol).scale_error_estimate(dt * ost.error_estimate, ost.reference_state)
    assert_close(m.cpu().numpy(), ost.mean, FP64_RTOL, "mean")
    assert_close(C.cpu().numpy(), ost.cov_sqrtm, FP64_RTOL, "cov_sqrtm_2")
    assert_close(float(err), ost.error_estimate, FP64_RTOL, "error_estimate")
    assert_close(ref.cpu().numpy(), ost.reference_state, FP64_RTOL, "reference_state")
    assert abs(math.sqrt(float(itol) / d) - ref_norm) <= FP64_RTOL * abs(ref_norm)


def test_config4_dek1_brusselator_1m_windows(cuda_device):
    """BASELINE config 4 (brusselator N = 1 000 000, DiagonalEK1 nu = 4): windows of chain points."""
    from tornadox_b200 import ivp
    from tornadox_b200.engine import StepEngine

    N, nu = 1_000_000, 4
    n, d = nu + 1, 2 * N
    p = ivp.brusselator(N=N)
    eng = StepEngine(p.spec, nu, device=DEV)
    mean, sc = _gpu_state(n, d, 13)
    mean[0, :N] += 1.0
    mean[0, N:] += 3.0
    dt, abstol, reltol = 1e-9, 1e-4, 1e-2
    m, s, err, ref, _ = eng.dek1_attempt(0.3, dt, mean, sc, abstol, reltol)
    const = (N + 1) ** 2 / 50.0                                   # ivp.py:223-224 with the GLOBAL N
    for i0, i1 in ((0, 300), (499_871, 500_129), (N - 300, N)):
        w = i1 - i0
        ow = oproblems.brusselator(N=w)                           # the oracle's f / df_diagonal take c as a keyword
        idx = torch.cat([torch.arange(f * N + i0, f * N + i1, device=DEV) for f in range(2)])
        f = lambda t, y: ow.f(t, y, c=const)
        dfd = lambda t, y: ow.df_diagonal(t, y, c=const)
        ost, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, mean[:, idx].cpu().numpy(), sc[idx].cpu().numpy(), None, None),
                                               dt, f, dfd, nu)
        lo = 0 if i0 == 0 else 1                                   # interior windows: the oracle pads u = 1, v = 3 there
        hi = w if i1 == N else w - 1
        keep = np.concatenate([np.arange(fl * w + lo, fl * w + hi) for fl in range(2)])
        sel = idx.cpu().numpy()[keep]
        assert_close(m[:, sel].cpu().numpy(), ost.mean[:, keep], FP64_RTOL, f"mean points {i0}:{i1}")
        assert_close(s[sel].cpu().numpy(), ost.cov_sqrtm[keep], FP64_RTOL, f"cov_sqrtm points {i0}:{i1}")
        assert_close(err[sel].cpu().numpy(), ost.error_estimate[keep], FP64_RTOL, f"error_estimate points {i0}:{i1}")


def test_config2_dek1_lorenz96_100k_constant_steps(cuda_device):
    """BASELINE config 2 (lorenz96 N = 100 000, DiagonalEK1 nu = 3, ConstantSteps): five chained steps in full."""
    from tornadox_b200 import ek1, ivp, step

    N, nu, dt = 100_000, 3, 0.01
    rng = np.random.default_rng(4)
    y0 = 8.0 + rng.standard_normal(N)
    o = oproblems.lorenz96(num_variables=N, y0=y0)
    p = ivp.lorenz96(y0=y0)
    sol = ek1.DiagonalEK1(num_derivatives=nu, steprule=step.ConstantSteps(dt))
    st = sol.initialize(*p)
    m0, c0 = odriver.taylor_init(odriver.series_lorenz96(8.0), y0, 0.0, nu)
    assert_close(st.y.mean.cpu().numpy(), m0, 1e-12, "Taylor-mode init")
    ost = ofilters.batched_initial_state(m0, c0, 0.0)
    for k in range(5):
        st, _ = sol.attempt_step(st, dt, *p)
        ost, _ = ofilters.diagonal_ek1_attempt(ost, dt, o.f, o.df_diagonal, nu)
        assert_close(st.y.mean.cpu().numpy(), ost.mean, 1e-9, f"mean after step {k + 1}")
        S = st.y.cov_sqrtm.cpu().numpy()
        assert_close(S @ S.transpose(0, 2, 1), ost.cov_sqrtm @ ost.cov_sqrtm.transpose(0, 2, 1), 1e-8, f"covariance after step {k + 1}")


def test_config3_ek0_fhn256_adaptive_same_step_sequence(cuda_device):
    """BASELINE config 3 (fhn_2d 256 x 256, d = 131 072, KroneckerEK0 nu = 4, AdaptiveSteps): identical accepted-step
    sequence, attempt counters and final state over the first stretch of the solve."""
    from tornadox_b200 import ek0, ivp, step

    side, nu = 256, 4
    dx = 1.0 / side
    y0 = np.random.default_rng(2).uniform(size=2 * side * side)
    tmax = 2e-4
    o = oproblems.fhn_2d(dx=dx, y0=y0, tmax=tmax)
    p = ivp.fhn_2d(dx=dx, y0=y0, tmax=tmax)
    m0, c0 = odriver.taylor_init(odriver.series_fhn_2d(side, side, dx), y0, 0.0, nu)
    states, info, trace = odriver.solve_kronecker_ek0(o, nu, odriver.AdaptiveSteps(), m0, c0)
    sol = ek0.KroneckerEK0(num_derivatives=nu, steprule=step.AdaptiveSteps())
    ts, final, final_info = [], None, None
    for st, inf in sol.solution_generator(p):
        ts.append(st.t)
        final, final_info = st, inf
    assert len(ts) == len(states) and len(ts) > 5, (len(ts), len(states))
    np.testing.assert_allclose(ts, [s.t for s in states], rtol=1e-9, atol=0)
    assert dict(final_info) == info
    # y0 is white noise on a 1/256 grid: the k-th derivative row has magnitude ~1e5^k and the solve amplifies rounding
    # differences over its ~10 steps, so every row is compared against its own scale
    fm = final.y.mean.cpu().numpy()
    for r in range(nu + 1):
        assert_close(fm[r], states[-1].mean[r], 1e-9 if r == 0 else 1e-6, f"final mean row {r}")
