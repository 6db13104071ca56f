"""GPU: the host-buffer entry point (tdx_dek1_attempt_step_host: chunked H2D | kernel | D2H pipeline) against the
NumPy oracle and against the device-resident call, for every chunk count incl. ragged splits and cyclic halos."""

import numpy as np
import pytest
import torch

from helpers import FP64_RTOL, assert_close, make_pair, random_state
from oracle.ref_np import driver as odriver
from oracle.ref_np import filters as ofilters

pytestmark = pytest.mark.gpu

CASES = [
    ("vanderpol", 1, [0]),
    ("lorenz96", 40000, [0, 1, 2, 5]),
    ("brusselator", 30011, [0, 3, 7]),
    ("fhn_2d", 50, [0, 1, 2, 3, 6]),
    ("fhn_2d", 128, [0, 16]),
]


def _host(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).pin_memory()


@pytest.mark.parametrize("kind,size,chunk_counts", CASES)
@pytest.mark.parametrize("jacobian", [True, False])
def test_host_pipeline_matches_oracle_and_device_path(cuda_device, kind, size, chunk_counts, jacobian):
    from tornadox_b200.engine import StepEngine

    o, p, _ = make_pair(kind, size)
    nu = 4
    n, d = nu + 1, o.y0.shape[0]
    rng = np.random.default_rng(size + 7)
    mean, sc = random_state(rng, n, d, "dense")
    mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
    dt = 0.0123 if kind != "fhn_2d" else 1e-3
    abstol, reltol = 1e-4, 1e-2
    df_diag = o.df_diagonal if jacobian else (lambda t, y: np.zeros_like(y))
    ost, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, mean, sc, None, None), dt, o.f, df_diag, nu)
    ref_norm = odriver.AdaptiveSteps(abstol=abstol, reltol=reltol).scale_error_estimate(dt * ost.error_estimate, ost.reference_state)

    eng = StepEngine(p.spec, nu, dtype=torch.float64, device="cuda:0")
    dm, dsc, derr, dref, dsq = eng.dek1_attempt(0.3, dt, torch.as_tensor(mean, device="cuda:0"), torch.as_tensor(sc, device="cuda:0"),
                                                abstol, reltol, jacobian=jacobian)
    h_mean, h_sc = _host(mean), _host(sc)
    for nchunks in chunk_counts:
        m, s, err, ref, sq = eng.dek1_attempt_host(0.3, dt, h_mean, h_sc, abstol, reltol, jacobian=jacobian, nchunks=nchunks)
        tag = f"{kind}/{size}/chunks={nchunks}"
        assert_close(m.numpy(), ost.mean, FP64_RTOL, f"{tag} mean")
        assert_close(s.numpy(), ost.cov_sqrtm, FP64_RTOL, f"{tag} cov_sqrtm")
        assert_close(err.numpy(), ost.error_estimate, FP64_RTOL, f"{tag} error_estimate")
        assert_close(ref.numpy(), ost.reference_state, FP64_RTOL, f"{tag} reference_state")
        assert abs(np.sqrt(sq / d) - ref_norm) <= 1e-10 * abs(ref_norm), tag
        # chunking must not change a single bit of the state (same kernel, same per-coordinate arithmetic)
        assert torch.equal(m, dm.cpu()) and torch.equal(s, dsc.cpu()), tag
        assert torch.equal(err, derr.cpu()) and torch.equal(ref, dref.cpu()), tag
        assert abs(sq - float(dsq.sum())) <= 1e-12 * abs(sq), tag
    # input buffers are untouched (a rejected attempt re-attempts from the same state)
    assert np.array_equal(h_mean.numpy(), mean) and np.array_equal(h_sc.numpy(), sc)


def test_solver_attempt_step_accepts_a_host_state(cuda_device):
    """Public API: DiagonalEK1.attempt_step with CPU tensors in the state routes through the host pipeline."""
    from tornadox_b200 import ek1, odefilter, rv, step

    o, p, _ = make_pair("fhn_2d", 64)
    sol = ek1.DiagonalEK1(num_derivatives=4, steprule=step.AdaptiveSteps())
    st = sol.initialize(*p)
    dt = 1e-4
    dev_new, _ = sol.attempt_step(st, dt, *p)
    host_state = odefilter.ODEFilterState(st.t, rv.BatchedMultivariateNormal(st.y.mean.cpu().pin_memory(), st.y.cov_sqrtm.cpu().pin_memory()),
                                          None, None)
    host_new, info = sol.attempt_step(host_state, dt, *p)
    assert host_new.y.mean.device.type == "cpu" and info["num_f_evaluations"] == 1
    assert torch.equal(host_new.y.mean, dev_new.y.mean.cpu())
    assert torch.equal(host_new.y.cov_sqrtm, dev_new.y.cov_sqrtm.cpu())
    n1 = step.norm_from_sq_sum(host_new.scaled_error_norm_sq_sum, host_new.scaled_error_norm_dim)
    n2 = step.norm_from_sq_sum(dev_new.scaled_error_norm_sq_sum, dev_new.scaled_error_norm_dim)
    assert abs(n1 - n2) <= 1e-12 * abs(n2)
    # chained host attempts alternate between the solver's two pinned output sets
    again, _ = sol.attempt_step(host_new, dt, *p)
    assert again.y.mean.data_ptr() != host_new.y.mean.data_ptr()
