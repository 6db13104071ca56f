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


@pytest.mark.parametrize("solver_name", ["ek0", "ek1", "dek0"])
@pytest.mark.parametrize("kind,size,tmax", [("vanderpol", 1, 1.0), ("fhn_2d", 32, 2e-3), ("lorenz96", 300, 1.0), ("brusselator", 200, 3e-3)])
def test_native_full_step_loop_equals_the_python_loop(cuda_device, solver_name, kind, size, tmax):
    """tdx_*_full_step (accept / reject in C, norm polled from pinned memory) against odefilter.perform_full_step in Python:
    same accepted times, same dt proposals, same counters, same final state."""
    from tornadox_b200 import ek0, ek1, step

    class PyLoopSteps(step.AdaptiveSteps):      # not the reference's exact class -> the solvers take the generic host loop
        pass

    _, p, _ = make_pair(kind, size, tmax=tmax) if kind == "vanderpol" else make_pair(kind, size)
    p = p._replace(tmax=tmax)
    cls = {"ek0": ek0.KroneckerEK0, "ek1": ek1.DiagonalEK1, "dek0": ek0.DiagonalEK0}[solver_name]
    runs = []
    for rule in (step.AdaptiveSteps(), PyLoopSteps()):
        sol = cls(num_derivatives=3, steprule=rule)
        ts, final, info = [], None, None
        for st, info in sol.solution_generator(p, device_loop=False):
            ts.append(st.t)
            final = st
        runs.append((ts, final, dict(info)))
    (ts_c, fin_c, info_c), (ts_py, fin_py, info_py) = runs
    assert info_c == info_py and len(ts_c) > 3
    np.testing.assert_allclose(ts_c, ts_py, rtol=1e-13, atol=0)
    assert torch.equal(fin_c.y.mean, fin_py.y.mean)


@pytest.mark.parametrize("solver_name", ["ek0", "ek1", "dek0"])
@pytest.mark.parametrize("kind,size,tmax", [("vanderpol", 1, 1.0), ("fhn_2d", 32, 2e-3), ("fhn_2d", 50, 1e-3), ("lorenz96", 300, 1.0),
                                            ("brusselator", 200, 3e-3), ("burgers_1d", 101, 0.02)])
@pytest.mark.parametrize("rule", ["adaptive", "constant"])
def test_device_resident_controller_equals_the_host_loop(cuda_device, solver_name, kind, size, tmax, rule):
    """simulate_final_state under the device-resident step controller (dt, accept / reject and the buffer flip decided by the
    kernels' last CTA, the host only enqueues batches) against the host-driven loop: same number of steps and attempts,
    same accepted times, same final state."""
    from tornadox_b200 import ek0, ek1, step

    _, p, _ = make_pair(kind, size, tmax=tmax) if kind == "vanderpol" else make_pair(kind, size)
    p = p._replace(tmax=tmax)
    cls = {"ek0": ek0.KroneckerEK0, "ek1": ek1.DiagonalEK1, "dek0": ek0.DiagonalEK0}[solver_name]

    def make_rule(solver):
        if rule == "adaptive":
            return step.AdaptiveSteps()
        return step.ConstantSteps(0.07 * tmax)

    host = cls(num_derivatives=3, steprule=make_rule(None))
    host.device_loop = False
    fin_h, info_h = host.simulate_final_state(p)
    dev = cls(num_derivatives=3, steprule=make_rule(None))
    dev.device_loop_batch = 7                       # several batches, and a tail of launches after the solve is over
    fin_d, info_d = dev.simulate_final_state(p)
    uses_ctl = True
    assert dict(info_d) == dict(info_h), (info_d, info_h)
    assert abs(fin_d.t - fin_h.t) <= 1e-12 * abs(fin_h.t)
    # device pow() in the dt proposal is within an ulp or two of the host's: the trajectories agree to rounding, not bit for bit
    assert_close(fin_d.y.mean.cpu().numpy(), fin_h.y.mean.cpu().numpy(), 1e-9, "final mean")
    if uses_ctl:
        st, ht, hd = dev.engine.ctl_read(history=True)
        assert st.done == 1 and st.failed == 0 and st.accepted == info_h["num_steps"]
        n_acc = int(st.accepted)
        ts = ht[:n_acc] if n_acc <= len(ht) else None
        if ts is not None:
            assert abs(ts[-1] - fin_h.t) <= 1e-12 * abs(fin_h.t) and np.all(np.diff(ts) > 0)


def test_solve_offload_streams_states_to_the_host(cuda_device):
    """solve(offload=True): accepted states are copied to pinned host memory as they are produced; same values as on the device."""
    from tornadox_b200 import ek0, ek1, step

    _, p, _ = make_pair("fhn_2d", 32)
    p = p._replace(tmax=1e-3)
    for cls in (ek1.DiagonalEK1, ek0.KroneckerEK0):
        dev = cls(num_derivatives=3, steprule=step.AdaptiveSteps()).solve(p)
        host = cls(num_derivatives=3, steprule=step.AdaptiveSteps()).solve(p, offload=True)
        assert host.mean.device.type == "cpu" and host.cov_sqrtm.device.type == "cpu"
        assert torch.equal(host.t, dev.t) and torch.equal(host.mean, dev.mean.cpu()) and torch.equal(host.cov_sqrtm, dev.cov_sqrtm.cpu())
        assert host.info == dev.info
    # KroneckerEK0 on a problem too large for the dense kron(I_d, C): the (n, n) factors are stored instead
    sol = ek0.KroneckerEK0(num_derivatives=3, steprule=step.AdaptiveSteps()).solve(p)
    assert sol.cov_sqrtm.shape[1:] == (4, 4)
