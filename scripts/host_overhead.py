"""Where does the host time of one adaptive attempt go? (development aid)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ek1, ek0, ivp, step

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
y0 = np.random.default_rng(2).uniform(size=2*side*side)
problem = ivp.fhn_2d(dx=1.0/side, y0=y0)
for cls in (ek1.DiagonalEK1, ek0.KroneckerEK0):
    sol = cls(num_derivatives=4, steprule=step.AdaptiveSteps())
    state = sol.initialize(*problem)
    dt = 0.5*sol.steprule.first_dt(*problem)
    for _ in range(5):
        state, dt, info = sol.perform_full_step(state, dt, *problem)
    torch.cuda.synchronize()
    # (a) attempts without any host sync
    t0 = time.perf_counter(); n = 50
    s = state
    for _ in range(n):
        s, _ = sol.attempt_step(s, dt, *problem)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{cls.__name__}: enqueue {1e3*(t1-t0)/n:.3f} ms/attempt, drained {1e3*(t2-t0)/n:.3f} ms/attempt")
    # (b) full adaptive loop
    t0 = time.perf_counter(); att = 0
    for _ in range(n):
        state, dt, info = sol.perform_full_step(state, dt, *problem)
        att += info["num_attempted_steps"]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{cls.__name__}: adaptive loop {1e3*(t1-t0)/att:.3f} ms/attempt ({att} attempts for {n} steps), dt={dt:.3e}")
