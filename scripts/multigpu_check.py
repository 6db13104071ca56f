"""Multi-GPU correctness check (launch with torchrun): sharded DiagonalEK1 / KroneckerEK0 attempts (halos and scalar
all-reduces through NVLink peer memory inside the step kernels) must reproduce the unsharded NumPy oracle, and -- where the
shard boundaries fall on the chunk grid of the reductions -- the SAME BITS of the error norm and the same accepted-step
sequence as an unsharded run on one GPU (step.py:93-104 is one norm).  Prints MULTIGPU_CHECK_OK on rank 0 and
writes a JSON summary to gpurun_out/multigpu_check_<world>gpu.json."""
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
from tornadox_b200.engine import local_slices, reduction_is_sharding_independent  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    worst = 0.0
    report = {"world": world, "attempts": [], "solves": []}
    from helpers import random_state
    from tornadox_b200 import odefilter, rv

    # fhn_2d 1024: the fused KroneckerEK0 kernel's row-block sums (8 grid rows per unit, large grids only) incl. the strips under a
    # shard's lower edge, which walk against the summation order
    for kind, size, solvers in (("fhn_2d", 64, ("dek1", "ek0")), ("fhn_2d", 50, ("dek1", "ek0")), ("brusselator", 1000, ("dek1", "ek0")),
                                ("lorenz96", 4099, ("dek1", "ek0")), ("fhn_2d", 1024, ("ek0",))):
        o, p, _ = make_pair(kind, size)
        nu = 4
        n, d = nu + 1, o.y0.shape[0]
        # a mid-solve looking state (right after an exact Taylor init the residual is rounding noise)
        rng = np.random.default_rng(size)
        if "dek1" in solvers:
            mean, blocks = random_state(rng, n, d, "dense")
        else:
            mean, blocks = rng.standard_normal((n, d)), None
        mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
        single = 1e-2 * rng.standard_normal((n, n))
        dt = 1e-4 if kind == "fhn_2d" else 1e-3
        if size >= 1024:
            dt = 1e-6
        for name in solvers:
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
            # the same attempt unsharded on this rank's GPU: the reduced sum must have the same bits
            one = cls(num_derivatives=nu, steprule=step.AdaptiveSteps())
            one.initialize(*p)
            mean_g = torch.as_tensor(mean, device=dev)
            if name == "dek1":
                y1 = rv.BatchedMultivariateNormal(mean_g, torch.as_tensor(blocks, device=dev))
            else:
                y1 = rv.LeftIsotropicMatrixNormal(mean_g, d, torch.as_tensor(single, device=dev))
            st1, _ = one.attempt_step(odefilter.ODEFilterState(0.3, y1, None, None), dt, *p)
            sum_sharded, sum_one = float(st.scaled_error_norm_sq_sum.sum()), float(st1.scaled_error_norm_sq_sum.sum())
            independent = reduction_is_sharding_independent(eng.spec, world)
            if independent:
                assert sum_sharded == sum_one, (kind, size, name, "norm sum differs from the single-GPU bits", sum_sharded, sum_one)
                assert torch.equal(st.y.mean, st1.y.mean[:, torch.as_tensor(idx, device=dev)]), (kind, size, name, "mean bits")
            else:
                assert abs(sum_sharded - sum_one) <= 1e-12 * abs(sum_one), (kind, size, name, sum_sharded, sum_one)
            report["attempts"].append(dict(problem=f"{kind} {size}", solver=name, sharding_independent=bool(independent),
                                           norm_sq_sum_sharded=sum_sharded.hex(), norm_sq_sum_single_gpu=sum_one.hex(),
                                           bit_equal=bool(sum_sharded == sum_one)))
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
    # (brusselator 4096: 16 tiles per field, so that 2, 4 and 8 equal shards are all cuts of the global summation grid)
    for kind, size, tmax in (("fhn_2d", 32, 2e-3), ("brusselator", 400, 3e-3), ("lorenz96", 600, 0.5), ("brusselator", 4096, 4e-5)):
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
            # unsharded on this rank's GPU, under the device-resident controller: the reference sequence of accepted steps
            one = cls(num_derivatives=nu, steprule=step.AdaptiveSteps())
            one.device_loop_batch = 64
            fin1, info1 = one.simulate_final_state(p)
            st1, ht1, hd1 = one.engine.ctl_read(history=True)
            for device_loop in (False, True):
                sol = cls(num_derivatives=nu, steprule=step.AdaptiveSteps(), process_group=group)
                sol.device_loop = device_loop
                sol.device_loop_batch = 5
                final, finfo = sol.simulate_final_state(p)
                eng = sol.engine
                if device_loop:
                    stN, htN, hdN = eng.ctl_read(history=True)
                    independent = reduction_is_sharding_independent(eng.spec, world)
                    n_acc = min(int(stN.accepted), len(htN))
                    same = bool(stN.accepted == st1.accepted and stN.attempts == st1.attempts and np.array_equal(htN[:n_acc], ht1[:n_acc])
                                and np.array_equal(hdN[:n_acc], hd1[:n_acc]))
                    if independent:
                        assert same, (kind, name, "accepted-step sequence differs from the single-GPU run", stN.accepted, st1.accepted)
                    report["solves"].append(dict(problem=f"{kind} {size}", solver=name, sharding_independent=bool(independent),
                                                 accepted=int(stN.accepted), attempts=int(stN.attempts),
                                                 identical_step_sequence_to_single_gpu=same))
                idx = np.concatenate([np.arange(o.y0.shape[0])[s] for s in local_slices(eng.spec, eng.begin, eng.end)])
                assert dict(finfo) == info, (kind, name, device_loop, finfo, info)
                assert abs(final.t - states[-1].t) <= 1e-12 * abs(states[-1].t), (kind, name, device_loop)
                e_fin = np.max(np.abs(final.y.mean.cpu().numpy() - states[-1].mean[:, idx])) / np.max(np.abs(states[-1].mean))
                assert e_fin < 1e-8, (kind, name, device_loop, e_fin)
                worst = max(worst, e_fin * 1e-2)
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        import json

        report["worst_rel_err_vs_oracle"] = float(t)
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, f"multigpu_check_{world}gpu.json"), "w") as fh:
            json.dump(report, fh, indent=1)
        # cases whose shard boundaries are boundaries of the global summation grid ("sharding_independent") MUST reproduce the
        # single-GPU bits; the others are deterministic and identical on all ranks, and are only counted
        ind_a = [a for a in report["attempts"] if a["sharding_independent"]]
        ind_s = [a for a in report["solves"] if a["sharding_independent"]]
        n_bits = sum(1 for a in ind_a if a["bit_equal"])
        n_seq = sum(1 for a in ind_s if a["identical_step_sequence_to_single_gpu"])
        report["summary"] = dict(bit_equal_norms=f"{n_bits}/{len(ind_a)} aligned cases", identical_step_sequences=f"{n_seq}/{len(ind_s)} aligned cases",
                                 unaligned_cases=len(report["attempts"]) - len(ind_a) + len(report["solves"]) - len(ind_s))
        with open(os.path.join(out_dir, f"multigpu_check_{world}gpu.json"), "w") as fh:
            json.dump(report, fh, indent=1)
        ok = n_bits == len(ind_a) and n_seq == len(ind_s)
        print(f"MULTIGPU_CHECK_{'OK' if ok else 'BITS_DIFFER'} world={world} worst_rel_err={float(t):.3e} bit_equal_norms={n_bits}/{len(ind_a)} "
              f"identical_step_sequences={n_seq}/{len(ind_s)} (aligned cases; {report['summary']['unaligned_cases']} unaligned cases not counted)")
        if not ok:
            raise SystemExit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
