"""debug: sharded KroneckerEK0 under the controller with ConstantSteps"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ek0, ek1, ivp, step
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
group = None
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    group = torch.distributed.group.WORLD
side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
y0 = np.random.default_rng(3).uniform(size=2 * side * side)
prob = ivp.fhn_2d(dx=1.0 / side, y0=y0, tmax=1.0)
for cls in (ek0.KroneckerEK0, ek1.DiagonalEK1):
    for k in (1, 2, 8):
        sol = cls(num_derivatives=4, steprule=step.ConstantSteps(1e-5), process_group=group)
        state = sol.initialize(*prob)
        eng = sol.engine
        ring = sol._ctl_ring(state, 2, clone_first=True)
        eng.ctl_begin(eng.native_steprule(sol.steprule), state.t, 1e-5, 1e300)
        sol._ctl_launch(k, ring)
        torch.cuda.synchronize()
        ctl = eng.ctl_read()
        m = ring["means"][ctl.cur]
        print(f"rank {rank} {cls.__name__} k={k}: attempts {ctl.attempts} accepted {ctl.accepted} failed {ctl.failed} t {ctl.t:.3e} last_norm {ctl.last_norm} "
              f"finite mean {bool(torch.isfinite(m).all())}", flush=True)
        if world > 1:
            torch.distributed.barrier()
if world > 1:
    torch.distributed.destroy_process_group()
