"""Per-CTA timeline of the fused KroneckerEK0 attempt (fhn_2d): one launch per attempt against the multi-attempt kernel.
Needs the instrumented library (python -m tornadox_b200.build --variant=timeline -DTDX_TIMELINE; TDX_LIB=.../libtornadox_b200_timeline.so)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tornadox_b200 import _lib, ek0, ivp, step


def read(eng, off, cnt):
    buf = np.zeros(cnt)
    _lib.check(eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), off, cnt))
    return buf


def report(eng, label, grid=296):
    st = read(eng, 4096, 8 * grid).reshape(grid, 8)
    t0 = st[:, 0].min()
    def q(x):
        x = (x - t0) * 1e-3
        return f"min {x.min():6.1f} / median {np.median(x):6.1f} / max {x.max():6.1f}"
    print(label)
    for name, col in (("start", 0), ("end of phase A", 1), ("gain seen", 2), ("end of phase B", 3), ("fenced", 4)):
        print(f"    {name:16s} {q(st[:, col])} us")
    dB = (st[:, 3] - st[:, 2]) * 1e-3
    print(f"    phase B duration per CTA: min {dB.min():.1f} / median {np.median(dB):.1f} / max {dB.max():.1f} us;  CTA 0: {dB[0]:.1f};"
          f"  slowest CTAs {np.argsort(dB)[-5:].tolist()} on SMs {st[np.argsort(dB)[-5:], 5].astype(int).tolist()}")


size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
y0 = np.random.default_rng(2).uniform(size=2 * size * size)
prob = ivp.fhn_2d(dx=1.0 / size, y0=y0, tmax=1.0)
for rule_name, mk in (("adaptive", lambda: step.AdaptiveSteps()), ("constant", lambda: step.ConstantSteps(1e-6))):
    sol = ek0.KroneckerEK0(num_derivatives=4, steprule=mk())
    state = sol.initialize(*prob)
    eng = sol.engine
    dt = 0.5 * step.AdaptiveSteps().first_dt(*prob)
    st = state
    for _ in range(3):
        st, _ = sol.attempt_step(st, dt, *prob)
    torch.cuda.synchronize()
    report(eng, f"{rule_name}: one launch per attempt (third launch)")
    del st
    ring = sol._ctl_ring(state, 2, clone_first=True)
    for k in (1, 4):
        eng.ctl_begin(eng.native_steprule(sol.steprule), state.t, dt, 1e300)
        sol._ctl_launch(k, ring)
        torch.cuda.synchronize()
        ctl = eng.ctl_read()
        report(eng, f"{rule_name}: controller kernel, attempt {int(ctl.attempts)} of {k}")
    del sol, state, ring
    torch.cuda.empty_cache()
