"""GPU parity: the CUDA attempt_step (through the C ABI) against the NumPy oracle on identical inputs."""

import math

import numpy as np
import pytest
import torch

from helpers import FP64_RTOL, assert_close, make_pair, random_state
from oracle.ref_np import driver as odriver
from oracle.ref_np import filters as ofilters

pytestmark = pytest.mark.gpu

CASES = [
    ("vanderpol", 1),
    ("vanderpol_julia", 1),
    ("threebody", 4),
    ("pleiades", 28),
    ("lorenz96", 10),
    ("lorenz96", 1000),
    ("lorenz96", 4099),
    ("brusselator", 20),
    ("brusselator", 517),
    ("brusselator", 4096),
    ("burgers_1d", 3),
    ("burgers_1d", 11),
    ("burgers_1d", 257),
    ("burgers_1d", 1001),
    ("wave_1d", 3),
    ("wave_1d", 11),
    ("wave_1d", 129),
    ("wave_1d", 1001),
    ("fhn_2d", 2),
    ("fhn_2d", 7),
    ("fhn_2d", 50),
    ("fhn_2d", 64),
]


def _engine(pivp, nu, dtype=torch.float64):
    from tornadox_b200.engine import StepEngine

    return StepEngine(pivp.spec, nu, dtype=dtype, device="cuda:0")


def _t(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


@pytest.mark.parametrize("kind,size", CASES)
def test_eval_f_and_jacobian_diagonal(cuda_device, kind, size):
    o, p, _ = make_pair(kind, size)
    rng = np.random.default_rng(5)
    y = o.y0 + 0.1 * rng.standard_normal(o.y0.shape)
    eng = _engine(p, 1)
    fy, jd = eng.eval_f(0.0, _t(y), want_jac=True)
    assert_close(fy.cpu().numpy(), o.f(0.0, y), 1e-13, "f")
    assert_close(jd.cpu().numpy(), o.df_diagonal(0.0, y), 1e-13, "df_diagonal")


@pytest.mark.parametrize("kind,size", CASES)
@pytest.mark.parametrize("nu", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("cov", ["zero", "dense"])
def test_dek1_attempt_matches_oracle(cuda_device, kind, size, nu, cov):
    o, p, _ = make_pair(kind, size)
    n, d = nu + 1, o.y0.shape[0]
    rng = np.random.default_rng(100 * nu + size)
    mean, sc = random_state(rng, n, d, cov)
    mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
    dt = 0.0123 if kind != "fhn_2d" else 1e-3
    abstol, reltol = 1e-4, 1e-2
    state = ofilters.FilterState(0.3, mean, sc, None, None)
    ref_state, _ = ofilters.diagonal_ek1_attempt(state, dt, o.f, o.df_diagonal, nu)
    steprule = odriver.AdaptiveSteps(abstol, reltol)
    ref_norm = steprule.scale_error_estimate(dt * ref_state.error_estimate, ref_state.reference_state)

    eng = _engine(p, nu)
    m, s, err, ref, sq = eng.dek1_attempt(0.3, dt, _t(mean), _t(sc), abstol=abstol, reltol=reltol)
    assert_close(m.cpu().numpy(), ref_state.mean, FP64_RTOL, "mean")
    assert_close(s.cpu().numpy(), ref_state.cov_sqrtm, FP64_RTOL, "cov_sqrtm (LAPACK sign convention)")
    S = s.cpu().numpy()
    assert_close(S @ S.transpose(0, 2, 1), ref_state.cov_sqrtm @ ref_state.cov_sqrtm.transpose(0, 2, 1), FP64_RTOL, "cov")
    assert_close(err.cpu().numpy(), ref_state.error_estimate, FP64_RTOL, "error_estimate")
    assert_close(ref.cpu().numpy(), ref_state.reference_state, FP64_RTOL, "reference_state")
    norm = math.sqrt(float(sq.sum()) / d)
    assert abs(norm - ref_norm) <= FP64_RTOL * abs(ref_norm), (norm, ref_norm)


@pytest.mark.parametrize("kind,size", CASES)
@pytest.mark.parametrize("nu", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("cov", ["zero", "dense"])
def test_ek0_attempt_matches_oracle(cuda_device, kind, size, nu, cov):
    o, p, _ = make_pair(kind, size)
    n, d = nu + 1, o.y0.shape[0]
    rng = np.random.default_rng(200 * nu + size)
    mean, _ = random_state(rng, n, d, "zero")
    mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
    Cl = np.zeros((n, n)) if cov == "zero" else 1e-2 * rng.standard_normal((n, n))
    dt = 0.0123 if kind != "fhn_2d" else 1e-3
    abstol, reltol = 1e-4, 1e-2
    state = ofilters.FilterState(0.3, mean, Cl, None, None)
    ref_state, _ = ofilters.kronecker_ek0_attempt(state, dt, o.f, nu)
    steprule = odriver.AdaptiveSteps(abstol, reltol)
    ref_norm = steprule.scale_error_estimate(dt * ref_state.error_estimate, ref_state.reference_state)

    eng = _engine(p, nu)
    m, C, err, ref, itol = eng.ek0_attempt(0.3, dt, _t(mean), _t(Cl), abstol=abstol, reltol=reltol)
    assert_close(m.cpu().numpy(), ref_state.mean, FP64_RTOL, "mean")
    assert_close(C.cpu().numpy(), ref_state.cov_sqrtm, FP64_RTOL, "cov_sqrtm_2")
    assert_close(float(err), ref_state.error_estimate, FP64_RTOL, "error_estimate")
    assert_close(ref.cpu().numpy(), ref_state.reference_state, FP64_RTOL, "reference_state")
    norm = math.sqrt(float(itol) / d)
    assert abs(norm - ref_norm) <= FP64_RTOL * abs(ref_norm), (norm, ref_norm)


@pytest.mark.parametrize("solver", ["ek0", "ek1"])
def test_config1_vanderpol_adaptive_solve_same_step_sequence(cuda_device, solver):
    """BASELINE config 1: vdp mu=1, t in [0, 1], nu=4, AdaptiveSteps, Taylor init: identical accepted steps."""
    from tornadox_b200 import ek0, ek1, step

    o, p, series_f = make_pair("vanderpol", 1, mu=1.0, tmax=1.0)
    nu = 4
    m0, c0 = odriver.taylor_init(series_f, o.y0, o.t0, nu)
    if solver == "ek0":
        states, info, trace = odriver.solve_kronecker_ek0(o, nu, odriver.AdaptiveSteps(), m0, c0)
        sol = ek0.KroneckerEK0(num_derivatives=nu, steprule=step.AdaptiveSteps())
    else:
        states, info, trace = odriver.solve_diagonal_ek1(o, nu, odriver.AdaptiveSteps(), m0, c0)
        sol = ek1.DiagonalEK1(num_derivatives=nu, steprule=step.AdaptiveSteps())
    ts = []
    final, final_info = None, None
    for st, inf in sol.solution_generator(p):
        ts.append(st.t)
        final, final_info = st, inf
    ref_ts = [s.t for s in states]
    assert len(ts) == len(ref_ts)
    np.testing.assert_allclose(ts, ref_ts, rtol=1e-9, atol=0)
    assert dict(final_info) == info
    assert_close(final.y.mean.cpu().numpy(), states[-1].mean, 1e-8, "final mean")


@pytest.mark.parametrize("size", [257, 300, 513])
def test_awkward_fhn_grids(cuda_device, size):
    """Grids that exercise the strip kernel's ragged last column block (300 = 256 + 44), a one-column block (257, 513) and, for
    odd sizes, the non-bulk staging path; DiagonalEK1 on the same grid for the smallest two."""
    o, p, _ = make_pair("fhn_2d", size)
    nu = 4
    n, d = nu + 1, o.y0.shape[0]
    rng = np.random.default_rng(size)
    mean, sc = random_state(rng, n, d, "dense")
    mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
    Cl = 1e-2 * rng.standard_normal((n, n))
    dt, abstol, reltol = 1e-4, 1e-4, 1e-2
    eng = _engine(p, nu)
    ref0, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(0.3, mean, Cl, None, None), dt, o.f, nu)
    m, C, err, ref, itol = eng.ek0_attempt(0.3, dt, _t(mean), _t(Cl), abstol=abstol, reltol=reltol)
    assert_close(m.cpu().numpy(), ref0.mean, FP64_RTOL, "ek0 mean")
    assert_close(C.cpu().numpy(), ref0.cov_sqrtm, FP64_RTOL, "ek0 cov_sqrtm_2")
    assert_close(ref.cpu().numpy(), ref0.reference_state, FP64_RTOL, "ek0 reference_state")
    ref_norm = odriver.AdaptiveSteps(abstol, reltol).scale_error_estimate(dt * ref0.error_estimate, ref0.reference_state)
    assert abs(math.sqrt(float(itol) / d) - ref_norm) <= FP64_RTOL * abs(ref_norm)
    if size <= 300:
        ref1, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, mean, sc, None, None), dt, o.f, o.df_diagonal, nu)
        m1, s1, err1, ref_1, _ = eng.dek1_attempt(0.3, dt, _t(mean), _t(sc), abstol=abstol, reltol=reltol)
        assert_close(m1.cpu().numpy(), ref1.mean, FP64_RTOL, "dek1 mean")
        assert_close(s1.cpu().numpy(), ref1.cov_sqrtm, FP64_RTOL, "dek1 cov_sqrtm")
        assert_close(err1.cpu().numpy(), ref1.error_estimate, FP64_RTOL, "dek1 error_estimate")


@pytest.mark.parametrize("kind,size", [("fhn_2d", 64), ("lorenz96", 4099)])
@pytest.mark.parametrize("nu", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_dek1_factor_buffers_without_16_byte_alignment(cuda_device, kind, size, nu, dtype):
    """Caller-owned buffers need element alignment only: factor arrays that start one element past a 16-byte boundary take
    the element-wise staging routes (no TMA bulk copies, no 16-byte units) and must give the same bits as aligned ones."""
    o, p, _ = make_pair(kind, size)
    n, d = nu + 1, o.y0.shape[0]
    rng = np.random.default_rng(7 * nu + size)
    mean, sc = random_state(rng, n, d, "dense")
    mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
    dt = 0.0123 if kind != "fhn_2d" else 1e-3
    eng = _engine(p, nu, dtype)
    m_in, sc_in = _t(mean, dtype), _t(sc, dtype)
    m0, s0, e0, r0, q0 = eng.dek1_attempt(0.3, dt, m_in, sc_in, abstol=1e-4, reltol=1e-2)

    def shifted(x):
        buf = torch.empty(x.numel() + 1, dtype=x.dtype, device=x.device)
        view = buf[1:].view(x.shape)
        view.copy_(x)
        assert view.data_ptr() % 16 != 0
        return view

    sc_shift, out_shift = shifted(sc_in), shifted(torch.zeros_like(sc_in))
    m1, s1, e1, r1, q1 = eng.dek1_attempt(0.3, dt, m_in, sc_shift, abstol=1e-4, reltol=1e-2, sc_out=out_shift)
    assert s1.data_ptr() == out_shift.data_ptr()
    for a, b, name in ((m0, m1, "mean"), (s0, s1, "cov_sqrtm"), (e0, e1, "error_estimate"), (r0, r1, "reference_state")):
        assert torch.equal(a, b), name
    assert float(q0.sum()) == float(q1.sum())
