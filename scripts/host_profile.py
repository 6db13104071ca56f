"""cProfile of the Python host path of one attempt (development aid)."""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ek1, ek0, ivp, step

side = 64
y0 = np.random.default_rng(2).uniform(size=2*side*side)
problem = ivp.fhn_2d(dx=1.0/side, y0=y0)
for cls in (ek0.KroneckerEK0, ek1.DiagonalEK1):
    sol = cls(num_derivatives=4, steprule=step.AdaptiveSteps())
    state = sol.initialize(*problem)
    dt = 1e-6
    for _ in range(20):
        s, _ = sol.attempt_step(state, dt, *problem)
    torch.cuda.synchronize()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n):
        s, _ = sol.attempt_step(state, dt, *problem)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{cls.__name__}: {1e6*(t1-t0)/n:.1f} us per attempt_step call (enqueue only)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        s, _ = sol.attempt_step(state, dt, *problem)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
