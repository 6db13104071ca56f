"""Multi-GPU correctness check (launch with torchrun): sharded DiagonalEK1 / KroneckerEK0 attempts with NCCL halo
exchange + scalar all-reduce must reproduce the unsharded NumPy oracle.  Prints MULTIGPU_CHECK_OK on rank 0."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from helpers import make_pair  # noqa: E402
from oracle.ref_np import driver as odriver  # noqa: E402
from oracle.ref_np import filters as ofilters  # noqa: E402
from tornadox_b200 import ek0, ek1, step  # noqa: E402
from tornadox_b200.engine import local_slices  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    worst = 0.0
    from helpers import random_state
    from tornadox_b200 import odefilter, rv

    for kind, size in (("fhn_2d", 64), ("fhn_2d", 50), ("brusselator", 1000), ("lorenz96", 4099)):
        o, p, _ = make_pair(kind, size)
        nu = 4
        n, d = nu + 1, o.y0.shape[0]
        # a mid-solve looking state (right after an exact Taylor init the residual is rounding noise)
        rng = np.random.default_rng(size)
        mean, blocks = random_state(rng, n, d, "dense")
        mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
        single = 1e-2 * rng.standard_normal((n, n))
        dt = 1e-4 if kind == "fhn_2d" else 1e-3
        for name in ("dek1", "ek0"):
            cls = ek1.DiagonalEK1 if name == "dek1" else ek0.KroneckerEK0
            sol = cls(num_derivatives=nu, steprule=step.AdaptiveSteps(), process_group=group)
            sol.initialize(*p)
            eng = sol.engine
            idx = np.concatenate([np.arange(d)[s] for s in local_slices(eng.spec, eng.begin, eng.end)])
            mean_l = torch.as_tensor(np.ascontiguousarray(mean[:, idx]), device=dev)
            if name == "dek1":
                y = rv.BatchedMultivariateNormal(mean_l, torch.as_tensor(np.ascontiguousarray(blocks[idx]), device=dev))
                ost, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, mean, blocks, None, None), dt, o.f,
                                                       o.df_diagonal, nu)
            else:
                y = rv.LeftIsotropicMatrixNormal(mean_l, mean_l.shape[1], torch.as_tensor(single, device=dev))
                ost, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(0.3, mean, single, None, None), dt, o.f, nu)
            st, _ = sol.attempt_step(odefilter.ODEFilterState(0.3, y, None, None), dt, *p)
            norm = step.norm_from_sq_sum(st.scaled_error_norm_sq_sum, d)
            ref_norm = odriver.AdaptiveSteps().scale_error_estimate(dt * ost.error_estimate, ost.reference_state)
            e_mean = np.max(np.abs(st.y.mean.cpu().numpy() - ost.mean[:, idx])) / np.max(np.abs(ost.mean))
            if name == "dek1":
                e_cov = np.max(np.abs(st.y.cov_sqrtm.cpu().numpy() - ost.cov_sqrtm[idx])) / np.max(np.abs(ost.cov_sqrtm))
                e_err = np.max(np.abs(st.error_estimate.cpu().numpy() - ost.error_estimate[idx])) / np.max(np.abs(ost.error_estimate))
            else:
                e_cov = np.max(np.abs(st.y.cov_sqrtm_2.cpu().numpy() - ost.cov_sqrtm)) / np.max(np.abs(ost.cov_sqrtm))
                e_err = abs(float(st.error_estimate) - ost.error_estimate) / abs(ost.error_estimate)
            e_norm = abs(norm - ref_norm) / abs(ref_norm)
            worst = max(worst, e_mean, e_cov, e_err, e_norm)
            assert max(e_mean, e_cov, e_err, e_norm) < 1e-10, (kind, size, name, dt, e_mean, e_cov, e_err, e_norm)
            if name == "dek1":
                # the shard's state in HOST memory: halo exchange of the edge rows + chunked pipeline must reproduce the same bits
                pin = lambda x: torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x)
                hm, hs, herr, href, hsq = eng.dek1_attempt_host(0.3, dt, pin(y.mean), pin(y.cov_sqrtm), 1e-4, 1e-2, nchunks=3)
                assert torch.equal(hm, st.y.mean.cpu()) and torch.equal(hs, st.y.cov_sqrtm.cpu()), (kind, size, "host path")
                assert torch.equal(herr, st.error_estimate.cpu()), (kind, size, "host path err")
                assert abs(math.sqrt(hsq / d) - norm) <= 1e-12 * norm, (kind, size, "host path norm")
            # f evaluation with halos
            yv = torch.as_tensor(o.y0[idx], device=dev)
            fy = eng.eval_f(0.0, yv)
            e_f = np.max(np.abs(fy.cpu().numpy() - o.f(0.0, o.y0)[idx])) / np.max(np.abs(o.f(0.0, o.y0)))
            assert e_f < 1e-13, (kind, e_f)
    # whole adaptive solves, sharded: the native accept / reject loop and the device-resident step controller must both
    # reproduce the unsharded oracle's accepted-step count and final state on every rank
    for kind, size, tmax in (("fhn_2d", 32, 2e-3), ("brusselator", 400, 3e-3), ("lorenz96", 600, 0.5)):
        o, p, series_f = make_pair(kind, size)
        o, p = o._replace(tmax=tmax), p._replace(tmax=tmax)
        nu = 3
        m0, c0 = odriver.taylor_init(series_f, o.y0, o.t0, nu)
        for name in ("dek1", "ek0"):
            if name == "dek1":
                states, info, _ = odriver.solve_diagonal_ek1(o, nu, odriver.AdaptiveSteps(), m0, c0)
                cls = ek1.DiagonalEK1
            else:
                states, info, _ = odriver.solve_kronecker_ek0(o, nu, odriver.AdaptiveSteps(), m0, c0)
                cls = ek0.KroneckerEK0
            for device_loop in (False, True):
                sol = cls(num_derivatives=nu, steprule=step.AdaptiveSteps(), process_group=group)
                sol.device_loop = device_loop
                sol.device_loop_batch = 5
                final, finfo = sol.simulate_final_state(p)
                eng = sol.engine
                idx = np.concatenate([np.arange(o.y0.shape[0])[s] for s in local_slices(eng.spec, eng.begin, eng.end)])
                assert dict(finfo) == info, (kind, name, device_loop, finfo, info)
                assert abs(final.t - states[-1].t) <= 1e-12 * abs(states[-1].t), (kind, name, device_loop)
                e_fin = np.max(np.abs(final.y.mean.cpu().numpy() - states[-1].mean[:, idx])) / np.max(np.abs(states[-1].mean))
                assert e_fin < 1e-8, (kind, name, device_loop, e_fin)
                worst = max(worst, e_fin * 1e-2)
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MULTIGPU_CHECK_OK world={world} worst_rel_err={float(t):.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
