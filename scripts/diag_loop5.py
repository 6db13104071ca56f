"""Diagnostic: KroneckerEK0 / DiagonalEK1 under the multi-attempt controller, launch by launch, with the debug notes of a
library built with -DTDX_DEBUG_WAITS (which wait gave up: 1 consumer full, 2 producer empty, 3 m1, 4 mid flag)."""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import _lib, ek0, ek1, ivp, step

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
which = sys.argv[2] if len(sys.argv) > 2 else "ek0"
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
y = np.random.default_rng(2).uniform(size=2 * side * side)
prob = ivp.fhn_2d(dx=1 / side, y0=y, tmax=2e-6 * (2048 / side) ** 2)
cls = ek0.KroneckerEK0 if which == "ek0" else ek1.DiagonalEK1
sol = cls(num_derivatives=4, steprule=step.AdaptiveSteps())
state = sol.initialize(*prob)
torch.cuda.synchronize()
print("initialised", flush=True)
eng = sol.engine
ring = sol._ctl_ring(state, 2)
eng.ctl_begin(eng.native_steprule(sol.steprule), state.t, sol.steprule.first_dt(*prob), prob.tmax)
for launch in range(6):
    t0 = time.perf_counter()
    sol._ctl_launch(k, ring)
    ctl = eng.ctl_read()
    notes = np.zeros(8)
    _lib.check(eng.lib.tdx_debug_read_workspace(eng.handle, notes.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 3000, 8))
    print(f"{cls.__name__} {side} launch {launch}: {1e3 * (time.perf_counter() - t0):8.2f} ms  t={ctl.t:.3e} dt={ctl.dt:.3e} norm={ctl.last_norm:.3e} "
          f"attempts={ctl.attempts} accepted={ctl.accepted} cur={ctl.cur} done={ctl.done} failed={ctl.failed} notes={notes[1:5]}", flush=True)
    if ctl.done:
        break
