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
    for kind, size in (("fhn_2d", 64), ("fhn_2d", 50), ("brusselator", 1000), ("lorenz96", 4099)):
        o, p, series_f = make_pair(kind, size)
        if kind == "lorenz96":
            # the default y0 sits on the equilibrium (f = 0 exactly at all but 3 points): the residual and hence the
            # factors are pure rounding noise there; perturb it so that the comparison is meaningful
            from oracle.ref_np import problems as oproblems
            from tornadox_b200 import ivp as pivp

            y0 = o.y0 + 0.5 * np.random.default_rng(0).standard_normal(o.y0.shape)
            o, p = oproblems.lorenz96(y0=y0), pivp.lorenz96(y0=y0)
        nu = 4
        m0, c0 = odriver.taylor_init(series_f, o.y0, 0.0, nu)
        d = o.y0.shape[0]
        dts = [1e-4, 2e-4] if kind == "fhn_2d" else [1e-3, 2e-3]
        for name in ("dek1", "ek0"):
            if name == "dek1":
                sol = ek1.DiagonalEK1(num_derivatives=nu, steprule=step.AdaptiveSteps(), process_group=group)
                ost = ofilters.batched_initial_state(m0, c0, 0.0)
            else:
                sol = ek0.KroneckerEK0(num_derivatives=nu, steprule=step.AdaptiveSteps(), process_group=group)
                ost = ofilters.kronecker_ek0_initial_state(m0, c0, 0.0)
            st = sol.initialize(*p)
            eng = sol.engine
            idx = np.concatenate([np.arange(d)[s] for s in local_slices(eng.spec, eng.begin, eng.end)])
            for dt in dts:
                st, _ = sol.attempt_step(st, dt, *p)
                if name == "dek1":
                    ost, _ = ofilters.diagonal_ek1_attempt(ost, dt, o.f, o.df_diagonal, nu)
                else:
                    ost, _ = ofilters.kronecker_ek0_attempt(ost, dt, o.f, nu)
                norm = math.sqrt(float(st.scaled_error_norm_sq_sum) / d)
                ref_norm = odriver.AdaptiveSteps().scale_error_estimate(dt * ost.error_estimate, ost.reference_state)
                e_mean = np.max(np.abs(st.y.mean.cpu().numpy() - ost.mean[:, idx])) / np.max(np.abs(ost.mean))
                if name == "dek1":
                    e_cov = np.max(np.abs(st.y.cov_sqrtm.cpu().numpy() - ost.cov_sqrtm[idx])) / np.max(np.abs(ost.cov_sqrtm))
                else:
                    e_cov = np.max(np.abs(st.y.cov_sqrtm_2.cpu().numpy() - ost.cov_sqrtm)) / np.max(np.abs(ost.cov_sqrtm))
                e_norm = abs(norm - ref_norm) / abs(ref_norm)
                worst = max(worst, e_mean, e_cov, e_norm)
                assert e_mean < 1e-9 and e_cov < 1e-8 and e_norm < 1e-9, (kind, size, name, dt, e_mean, e_cov, e_norm)
            # f evaluation with halos
            y = torch.as_tensor(o.y0[idx], device=dev)
            fy = eng.eval_f(0.0, y)
            e_f = np.max(np.abs(fy.cpu().numpy() - o.f(0.0, o.y0)[idx])) / np.max(np.abs(o.f(0.0, o.y0)))
            assert e_f < 1e-13, (kind, e_f)
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MULTIGPU_CHECK_OK world={world} worst_rel_err={float(t):.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
