"""GPU: the device-resident step controller as ONE persistent launch per batch of attempts (``tdx_*_run_attempts``,
reference: the compiled accept / reject loop odefilter.py:225-344), dense output by following the running kernel
(odefilter.py:51-71), time stops on the device (odefilter.py:355-373), and the launch-geometry independence of the
kernels' norm sums (step.py:93-104 is ONE norm over all coordinates)."""

import os

import numpy as np
import pytest
import torch

from helpers import assert_close, make_pair
from oracle.ref_np import driver as odriver

pytestmark = pytest.mark.gpu

CASES = [("vanderpol", 1, 1.0), ("fhn_2d", 32, 2e-3), ("fhn_2d", 50, 1e-3), ("lorenz96", 300, 1.0), ("brusselator", 200, 3e-3)]


def _problem(kind, size, tmax):
    _, p, _ = make_pair(kind, size, tmax=tmax) if kind == "vanderpol" else make_pair(kind, size)
    return p._replace(tmax=tmax)


def _solver(name, rule, nu=3):
    from tornadox_b200 import ek0, ek1

    cls = {"ek0": ek0.KroneckerEK0, "ek1": ek1.DiagonalEK1, "dek0": ek0.DiagonalEK0}[name]
    return cls(num_derivatives=nu, steprule=rule)


@pytest.mark.parametrize("solver_name", ["ek0", "ek1", "dek0"])
@pytest.mark.parametrize("kind,size,tmax", CASES)
@pytest.mark.parametrize("rule", ["adaptive", "constant"])
def test_one_launch_per_batch_equals_one_launch_per_attempt(cuda_device, solver_name, kind, size, tmax, rule):
    """k attempts inside one persistent kernel (in-kernel grid barrier, controller update, buffer ring) against the same k
    attempts as k launches of the same kernel: the same bits everywhere."""
    from tornadox_b200 import step

    p = _problem(kind, size, tmax)
    make_rule = (lambda: step.AdaptiveSteps()) if rule == "adaptive" else (lambda: step.ConstantSteps(0.07 * tmax))

    multi = _solver(solver_name, make_rule())
    multi.device_loop_batch = 5
    fin_m, info_m = multi.simulate_final_state(p)
    st_m, ht_m, hd_m = multi.engine.ctl_read(history=True)

    single = _solver(solver_name, make_rule())
    state = single.initialize(*p)
    eng = single.engine
    ring = single._ctl_ring(state, 2)
    eng.ctl_begin(eng.native_steprule(single.steprule), state.t, single.steprule.first_dt(*p), p.tmax)
    while True:
        pair_a = (ring["means"][0], ring["factors"][0])
        pair_b = (ring["means"][1], ring["factors"][1])
        if solver_name == "ek0":
            eng.ek0_enqueue_attempts(3, pair_a, pair_b, ring["errs"][0])
        else:
            eng.dek1_enqueue_attempts(3, pair_a, pair_b, None, None, jacobian=single._uses_jacobian)
        ctl = eng.ctl_read()
        assert not ctl.failed
        if ctl.done:
            break
    st_s, ht_s, hd_s = eng.ctl_read(history=True)
    assert (st_m.attempts, st_m.accepted, st_m.cur, st_m.done, st_m.failed) == (st_s.attempts, st_s.accepted, st_s.cur, st_s.done, 0)
    assert st_m.t == st_s.t and info_m["num_attempted_steps"] == st_s.attempts and info_m["num_steps"] == st_s.accepted
    n_acc = min(int(st_m.accepted), len(ht_m))
    assert np.array_equal(ht_m[:n_acc], ht_s[:n_acc]) and np.array_equal(hd_m[:n_acc], hd_s[:n_acc])
    assert torch.equal(fin_m.y.mean, ring["means"][st_s.cur])
    fac = fin_m.y.cov_sqrtm_2 if solver_name == "ek0" else fin_m.y.cov_sqrtm
    assert torch.equal(fac, ring["factors"][st_s.cur])


@pytest.mark.parametrize("solver_name", ["ek0", "ek1"])
@pytest.mark.parametrize("kind,size,tmax", [("vanderpol", 1, 1.0), ("fhn_2d", 32, 2e-3), ("lorenz96", 300, 1.0), ("brusselator", 200, 3e-3)])
def test_solve_follows_the_running_kernel(cuda_device, solver_name, kind, size, tmax):
    """solve(): the whole solve is ONE persistent launch on its own stream, the host follows it through mapped pinned memory and
    copies every accepted state out of the state ring.  Same accepted times, counters and states as the host-driven loop."""
    from tornadox_b200 import step

    p = _problem(kind, size, tmax)
    host = _solver(solver_name, step.AdaptiveSteps())
    host.device_loop = False
    ref = host.solve(p)
    for offload in (False, True):
        sol = _solver(solver_name, step.AdaptiveSteps())
        sol.ring_size = 3
        out = sol.solve(p, offload=offload)
        assert out.info == ref.info, (out.info, ref.info)
        assert out.t.shape == ref.t.shape
        np.testing.assert_allclose(out.t.numpy(), ref.t.numpy(), rtol=1e-9, atol=0)
        assert out.mean.shape == ref.mean.shape and out.cov_sqrtm.shape == ref.cov_sqrtm.shape
        assert (out.mean.device.type == "cpu") == offload
        # device pow() in the dt proposal is within an ulp or two of the host's: equal to rounding, not bit for bit
        assert_close(out.mean.cpu().numpy(), ref.mean.cpu().numpy(), 1e-9, "means of all accepted states")
        # factors are unique up to column signs (LAPACK's beta = -sign(alpha) |x| flips with the last bit of alpha): compare S S^T
        S, R = out.cov_sqrtm.cpu().numpy(), ref.cov_sqrtm.cpu().numpy()
        assert_close(S @ np.swapaxes(S, -1, -2), R @ np.swapaxes(R, -1, -2), 1e-7, "covariances of all accepted states")


@pytest.mark.parametrize("solver_name", ["ek0", "ek1"])
def test_solution_generator_yields_copies_from_the_ring(cuda_device, solver_name):
    """solution_generator under the controller: every yielded state stays valid (it is a copy), info counters are those of
    the host loop, and abandoning the generator stops the kernel."""
    from tornadox_b200 import step

    p = _problem("fhn_2d", 32, 2e-3)
    host = _solver(solver_name, step.AdaptiveSteps())
    ref = [(st.t, st.y.mean.clone(), dict(info)) for st, info in host.solution_generator(p, device_loop=False)]
    sol = _solver(solver_name, step.AdaptiveSteps())
    got = [(st.t, st.y.mean, dict(info)) for st, info in sol.solution_generator(p)]
    assert len(got) == len(ref)
    for (t1, m1, i1), (t2, m2, i2) in zip(got, ref):
        assert abs(t1 - t2) <= 1e-9 * max(abs(t2), 1e-300) and i1 == i2
        assert_close(m1.cpu().numpy(), m2.cpu().numpy(), 1e-9, "yielded mean")
    # abandon a running solve
    sol2 = _solver(solver_name, step.AdaptiveSteps())
    gen = sol2.solution_generator(p._replace(tmax=1.0))
    for k, (st, info) in enumerate(gen):
        if k == 3:
            break
    gen.close()
    torch.cuda.synchronize()
    ctl = sol2.engine.ctl_read()
    assert not ctl.failed and ctl.t < 1.0          # stopped: by the abort flag, or at the end of its bounded launch


@pytest.mark.parametrize("solver_name", ["ek0", "ek1"])
@pytest.mark.parametrize("kind,size,tmax", [("vanderpol", 1, 1.0), ("fhn_2d", 32, 2e-3), ("lorenz96", 300, 1.0)])
def test_time_stops_on_the_device(cuda_device, solver_name, kind, size, tmax):
    """stop_at under the controller (odefilter.py:355-373): the solve lands exactly on every stop; same step sequence as the
    host loop with the reference's _TimeStopper and as the oracle's driver."""
    from tornadox_b200 import step

    o, p, series_f = make_pair(kind, size, tmax=tmax) if kind == "vanderpol" else make_pair(kind, size)
    o, p = o._replace(tmax=tmax), p._replace(tmax=tmax)
    stops = [0.1 * tmax, 0.25 * tmax, 0.25 * tmax + 1e-3 * tmax, 0.8 * tmax]
    host = _solver(solver_name, step.AdaptiveSteps())
    ts_host = [st.t for st, _ in host.solution_generator(p, stop_at=stops, device_loop=False)]
    dev = _solver(solver_name, step.AdaptiveSteps())
    ts_dev, infos = [], []
    for st, info in dev.solution_generator(p, stop_at=stops):
        ts_dev.append(st.t)
        infos.append(dict(info))
    assert len(ts_dev) == len(ts_host)
    np.testing.assert_allclose(ts_dev, ts_host, rtol=1e-9, atol=0)
    for s in stops:
        assert any(t == s for t in ts_dev), (s, ts_dev)
    fin, info = _solver(solver_name, step.AdaptiveSteps()).simulate_final_state(p, stop_at=stops)
    assert info == infos[-1] and abs(fin.t - ts_host[-1]) <= 1e-9 * abs(ts_host[-1])
    # the oracle's driver (op-for-op restatement of the reference's loop) takes the same steps
    m0, c0 = odriver.taylor_init(series_f, o.y0, o.t0, 3)
    solve = odriver.solve_kronecker_ek0 if solver_name == "ek0" else odriver.solve_diagonal_ek1
    states, oinfo, _ = solve(o, 3, odriver.AdaptiveSteps(), m0, c0, stop_at=stops)
    assert oinfo["num_steps"] == info["num_steps"] and oinfo["num_attempted_steps"] == info["num_attempted_steps"]
    np.testing.assert_allclose(ts_dev, [s.t for s in states], rtol=1e-9, atol=0)


@pytest.mark.parametrize("kind,size", [("fhn_2d", 64), ("fhn_2d", 50), ("fhn_2d", 300), ("lorenz96", 4099), ("brusselator", 3000)])
def test_norm_sums_do_not_depend_on_the_launch_geometry(cuda_device, kind, size, monkeypatch):
    """One partial per tile / strip row, fixed chunk order: the reduced sums are bit-identical whatever the number of CTAs
    (what makes the accepted-step sequence independent of the number of GPUs needs this first)."""
    from helpers import random_state
    from tornadox_b200 import ek0, ek1, odefilter, rv, step

    o, p, _ = make_pair(kind, size)
    nu = 4
    n, d = nu + 1, o.y0.shape[0]
    rng = np.random.default_rng(size)
    mean, blocks = random_state(rng, n, d, "dense")
    mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
    single = 1e-2 * rng.standard_normal((n, n))
    dt = 1e-4 if kind == "fhn_2d" else 1e-3
    dev = torch.device("cuda:0")
    results = {}
    for cap in (None, "1", "7", "40"):
        if cap is None:
            monkeypatch.delenv("TDX_MAX_CTAS", raising=False)
        else:
            monkeypatch.setenv("TDX_MAX_CTAS", cap)
        for name in ("dek1", "ek0"):
            cls = ek1.DiagonalEK1 if name == "dek1" else ek0.KroneckerEK0
            sol = cls(num_derivatives=nu, steprule=step.AdaptiveSteps())
            sol.initialize(*p)
            if name == "dek1":
                y = rv.BatchedMultivariateNormal(torch.as_tensor(mean, device=dev), torch.as_tensor(blocks, device=dev))
            else:
                y = rv.LeftIsotropicMatrixNormal(torch.as_tensor(mean, device=dev), d, torch.as_tensor(single, device=dev))
            st, _ = sol.attempt_step(odefilter.ODEFilterState(0.3, y, None, None), dt, *p)
            val = float(st.scaled_error_norm_sq_sum.sum())
            results.setdefault(name, []).append((cap, val, st.y.mean.clone()))
    for name, runs in results.items():
        base = runs[0]
        for cap, val, m in runs[1:]:
            assert val == base[1], (name, cap, val, base[1])
            assert torch.equal(m, base[2]), (name, cap)
