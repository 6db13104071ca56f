"""Does CUDA-graph replay of controller-driven attempts shrink the launch gaps? (development aid)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ek1, ek0, ivp, step

def probe(name, cls, prob, nu, k=64):
    sol = cls(num_derivatives=nu, steprule=step.AdaptiveSteps())
    state = sol.initialize(*prob)
    eng = sol.engine
    rule = eng.native_steprule(sol.steprule)
    is_dek1 = isinstance(sol, ek1.DiagonalEK1)
    pair_a = (state.y.mean, state.y.cov_sqrtm if is_dek1 else state.y.cov_sqrtm_2)
    pair_b = (torch.empty_like(pair_a[0]), torch.empty_like(pair_a[1]))
    err = torch.empty((), dtype=torch.float64, device="cuda")
    dt = 0.5 * sol.steprule.first_dt(*prob)
    def enqueue():
        if is_dek1: eng.dek1_enqueue_attempts(k, pair_a, pair_b, None, None)
        else: eng.ek0_enqueue_attempts(k, pair_a, pair_b, err, None)
    eng.ctl_begin(rule, 0.0, dt, 1e300); enqueue(); eng.ctl_read()
    torch.cuda.synchronize()
    eng.ctl_begin(rule, 0.0, dt, 1e300)
    t0 = time.perf_counter(); enqueue(); st = eng.ctl_read(); t1 = time.perf_counter()
    print(f"{name}: plain launches {1e6*(t1-t0)/k:.1f} us/attempt", flush=True)
    if not is_dek1:
        return      # the EK0 kernel's barrier sequence number is a launch parameter: not replayable yet
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng.ctl_begin(rule, 0.0, dt, 1e300)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            enqueue()
        torch.cuda.synchronize()
        eng.ctl_begin(rule, 0.0, dt, 1e300)
        t0 = time.perf_counter(); g.replay(); st = eng.ctl_read(); t1 = time.perf_counter()
        print(f"{name}: graph replay   {1e6*(t1-t0)/k:.1f} us/attempt (attempts {st.attempts})", flush=True)

y0 = np.random.default_rng(2).uniform(size=2 * 256 * 256)
probe("fhn 256^2 dek1", ek1.DiagonalEK1, ivp.fhn_2d(dx=1/256, y0=y0), 4)
probe("lorenz 100k dek1 nu=3", ek1.DiagonalEK1, ivp.lorenz96(num_variables=100_000), 3)
probe("fhn 256^2 ek0", ek0.KroneckerEK0, ivp.fhn_2d(dx=1/256, y0=y0), 4)
probe("lorenz 100k ek0", ek0.KroneckerEK0, ivp.lorenz96(num_variables=100_000), 3)
