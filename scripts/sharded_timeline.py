"""Stations of ONE sharded attempt (one launch per attempt) on every rank, BASELINE config 5 cut over the ranks (torchrun).
Needs the instrumented library (TDX_LIB=.../libtornadox_b200_timeline.so, built with -DTDX_TIMELINE)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tornadox_b200 import _lib, ek0, ek1, ivp, step


def read(eng, off, cnt):
    buf = np.zeros(cnt)
    _lib.check(eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), off, cnt))
    return buf


world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
group = None
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    group = torch.distributed.group.WORLD
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
y0 = np.random.default_rng(2).uniform(size=2 * size * size)
prob = ivp.fhn_2d(dx=1.0 / size, y0=y0, tmax=1.0)
for name, cls in (("DiagonalEK1", ek1.DiagonalEK1), ("KroneckerEK0", ek0.KroneckerEK0)):
    sol = cls(num_derivatives=4, steprule=step.AdaptiveSteps(), process_group=group)
    state = sol.initialize(*prob)
    eng = sol.engine
    dt = 0.5 * sol.steprule.first_dt(*prob)
    st = state
    lines = []
    for it in range(6):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # steady state: ten launches back to back (the host runs ahead, the ranks pace one another through the exchanges); the
        # stamps read afterwards are those of the LAST launch
        for _ in range(9):
            st, _ = sol.attempt_step(st, dt, *prob)
        a.record()
        st, _ = sol.attempt_step(st, dt, *prob)
        b.record()
        torch.cuda.synchronize()
        s = read(eng, 2048, 8)            # attempt slot 0: [start, CTA 0 out of work, end, last CTA arrived, reduced, (ctl), ...]
        ms = a.elapsed_time(b) * 1e3
        if name == "DiagonalEK1":
            lines.append(f"launch {it}: {ms:6.1f} us on the stream | CTA 0: tiles {1e-3 * (s[1] - s[0]):6.1f} us; last CTA arrives {1e-3 * (s[3] - s[0]):6.1f} us "
                         f"after CTA 0's start; local sums + exchange with the peers {1e-3 * (s[4] - s[3]):5.1f} us")
        else:
            lines.append(f"launch {it}: {ms:6.1f} us on the stream | CTA 0: phase A {1e-3 * (s[1] - s[0]):5.1f}, barrier + exchange + gain {1e-3 * (s[2] - s[1]):5.1f}, "
                         f"phase B {1e-3 * (s[3] - s[2]):5.1f}, end barrier + exchange {1e-3 * (s[4] - s[3]):5.1f} us")
    for r in range(world):
        if world > 1:
            torch.distributed.barrier()
        if r == rank and rank in (0, world // 2, world - 1):
            print(f"--- {name}, rank {rank} of {world}")
            print("\n".join(lines[2:]), flush=True)
    del sol, state, st
    torch.cuda.empty_cache()
if world > 1:
    torch.distributed.destroy_process_group()
