"""Per-attempt time of the multi-attempt (controller) kernels against one launch per attempt at BASELINE config 5 (development aid)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tornadox_b200 import ek0, ek1, ivp, step

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
y0 = np.random.default_rng(2).uniform(size=2 * size * size)
prob = ivp.fhn_2d(dx=1.0 / size, y0=y0, tmax=1.0)
for name, cls in (("DiagonalEK1", ek1.DiagonalEK1), ("KroneckerEK0", ek0.KroneckerEK0)):
    for rule_name, mk in (("adaptive", lambda: step.AdaptiveSteps()), ("constant", lambda: step.ConstantSteps(1e-6))):
        sol = cls(num_derivatives=4, steprule=mk())
        state = sol.initialize(*prob)
        eng = sol.engine
        dt = 0.5 * step.AdaptiveSteps().first_dt(*prob)
        # one launch per attempt through the public API
        st = state
        for _ in range(3):
            st, _ = sol.attempt_step(st, dt, *prob)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            st, _ = sol.attempt_step(st, dt, *prob)
        b.record()
        torch.cuda.synchronize()
        per_launch = a.elapsed_time(b) * 1e3 / k
        del st
        ring = sol._ctl_ring(state, 2, clone_first=True)
        best = None
        for _ in range(3):
            eng.ctl_begin(eng.native_steprule(sol.steprule), state.t, dt, 1e300)
            torch.cuda.synchronize()
            a.record()
            sol._ctl_launch(k, ring)
            b.record()
            torch.cuda.synchronize()
            ctl = eng.ctl_read()
            us = a.elapsed_time(b) * 1e3 / max(1, int(ctl.attempts))
            best = us if best is None else min(best, us)
        print(f"fhn {size}^2 {name:12s} {rule_name:8s}: one launch per attempt {per_launch:7.1f} us | {int(ctl.attempts)} attempts in one launch "
              f"{best:7.1f} us/attempt ({int(ctl.accepted)} accepted, failed={int(ctl.failed)})", flush=True)
        del sol, state, ring
        torch.cuda.empty_cache()
