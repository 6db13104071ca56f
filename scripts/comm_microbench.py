"""What one attempt costs on top of the shard's own work when the solve is sharded over G GPUs (torchrun, one rank per GPU).

Every rank holds a SMALL shard (64 grid rows x 256 columns of fhn_2d, 16 tiles / 8 strips), so the attempt is nothing but its fixed
costs; K attempts run inside ONE persistent launch under the device-resident controller (no host in the loop).  Reported per
attempt, max over ranks, device-timed:
  local   : the same shard as a stand-alone problem (no peers): kernel latency chain + controller
  sharded : halo rows delivered into the neighbours' buffers + flag wait, chunk-sum exchange with all peers (DiagonalEK1: one per
            attempt, KroneckerEK0: two), controller
  sharded - local = the cross-GPU part (what verdict item 5 asks to be measured).
Writes gpurun_out/comm_microbench_<G>gpu.json (rank 0)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tornadox_b200 import ek0, ek1, ivp, step  # noqa: E402

K = 128
ROWS, COLS = 64, 256


def per_attempt(sol, problem, world, k=K):
    state = sol.initialize(*problem)
    eng = sol.engine
    rule = eng.native_steprule(sol.steprule)
    dt = 0.5 * sol.steprule.first_dt(*problem)
    ring = sol._ctl_ring(state, 2, clone_first=True)
    fac0 = state.y.cov_sqrtm if hasattr(state.y, "cov_sqrtm") and not hasattr(state.y, "cov_sqrtm_2") else state.y.cov_sqrtm_2

    def run():
        ring["means"][0].copy_(state.y.mean)          # every run starts from the initial state (slot 0)
        ring["factors"][0].copy_(fac0)
        eng.ctl_begin(rule, state.t, dt, 1e300)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sol._ctl_launch(k, ring)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    run()
    ms = min(run() for _ in range(3))
    ctl = eng.ctl_read()
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t) * 1e3 / max(1, int(ctl.attempts)), int(ctl.attempts), int(ctl.failed)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    group = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        group = torch.distributed.group.WORLD
    out = {"gpus": world, "attempts_per_launch": K, "shard": f"fhn_2d {ROWS} x {COLS} grid rows per rank, nu=4, fp64", "us_per_attempt": {}}
    rng = np.random.default_rng(2)
    for name, cls in (("DiagonalEK1", ek1.DiagonalEK1), ("KroneckerEK0", ek0.KroneckerEK0)):
        for rule_name, mk in (("adaptive", lambda: step.AdaptiveSteps()), ("constant", lambda: step.ConstantSteps(2e-6))):
            # local: the shard alone (a ROWS x COLS problem is not square: fhn_2d takes bbox / dx, so use a square of the same
            # number of tiles instead -- 128 x 128 has the same 16 tiles and 2 x 4 strips)
            side = int(round((ROWS * COLS) ** 0.5))
            y0 = rng.uniform(size=2 * side * side)
            local = per_attempt(cls(num_derivatives=4, steprule=mk()), ivp.fhn_2d(dx=1.0 / side, y0=y0, tmax=1.0), 1)
            res = {"local": local[0], "local_attempts": local[1]}
            if world > 1:
                # sharded: a (side * world) x ... grid would change the tile count per rank; take a square grid whose shard has
                # about the same number of coordinates: G * side^2 points in total
                gside = int(round((world * side * side) ** 0.5 / world)) * world
                y0g = np.random.default_rng(3).uniform(size=2 * gside * gside)
                sh = per_attempt(cls(num_derivatives=4, steprule=mk(), process_group=group), ivp.fhn_2d(dx=1.0 / gside, y0=y0g, tmax=1.0), world)
                res.update({"sharded": sh[0], "sharded_attempts": sh[1], "sharded_failed": sh[2], "sharded_grid": gside,
                            "cross_gpu": sh[0] - local[0]})
            out["us_per_attempt"][f"{name}/{rule_name}"] = res
            if rank == 0:
                print(name, rule_name, res, flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/comm_microbench_{world}gpu.json", "w") as fh:
            json.dump(out, fh, indent=1)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
