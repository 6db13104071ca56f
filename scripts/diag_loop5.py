"""Diagnostic: KroneckerEK0 / DiagonalEK1 at config 5 under the multi-attempt controller, launch by launch."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ek0, ek1, ivp, step

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
y = np.random.default_rng(2).uniform(size=2 * side * side)
prob = ivp.fhn_2d(dx=1 / side, y0=y, tmax=2e-6 * (2048 / side) ** 2)
for cls in (ek0.KroneckerEK0, ek1.DiagonalEK1):
    sol = cls(num_derivatives=4, steprule=step.AdaptiveSteps())
    state = sol.initialize(*prob)
    eng = sol.engine
    ring = sol._ctl_ring(state, 2)
    eng.ctl_begin(eng.native_steprule(sol.steprule), state.t, sol.steprule.first_dt(*prob), prob.tmax)
    for launch in range(12):
        t0 = time.perf_counter()
        sol._ctl_launch(4, ring)
        ctl = eng.ctl_read()
        print(f"{cls.__name__} launch {launch}: {1e3 * (time.perf_counter() - t0):8.2f} ms  t={ctl.t:.3e} dt={ctl.dt:.3e} norm={ctl.last_norm:.3e} "
              f"attempts={ctl.attempts} accepted={ctl.accepted} cur={ctl.cur} done={ctl.done} failed={ctl.failed}", flush=True)
        if ctl.done:
            break
