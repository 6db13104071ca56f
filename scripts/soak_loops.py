"""Large-scale soak: ~100+ adaptive attempts of BASELINE configs 4 and 5 under the host-driven native loop and under the
device-resident controller; both must take the same steps and end in the same state (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ek0, ek1, ivp, step

def run(cls, prob, nu, dl):
    sol = cls(num_derivatives=nu, steprule=step.AdaptiveSteps())
    sol.device_loop = dl
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st, info = sol.simulate_final_state(prob)
    torch.cuda.synchronize()
    return st, info, time.perf_counter() - t0

y5 = np.random.default_rng(2).uniform(size=2 * 2048 * 2048)
cases = [("config4 brusselator 1M dek1", ek1.DiagonalEK1, ivp.brusselator(N=1_000_000, tmax=2e-9), 4),
         ("config5 fhn 2048^2 dek1", ek1.DiagonalEK1, ivp.fhn_2d(dx=1/2048, y0=y5, tmax=4e-6), 4),
         ("config5 fhn 2048^2 ek0", ek0.KroneckerEK0, ivp.fhn_2d(dx=1/2048, y0=y5, tmax=4e-6), 4)]
for name, cls, prob, nu in cases:
    a, ia, ta = run(cls, prob, nu, False)
    b, ib, tb = run(cls, prob, nu, True)
    same = dict(ia) == dict(ib)
    scale = a.y.mean.abs().amax(dim=1, keepdim=True)
    rel = float(((a.y.mean - b.y.mean).abs() / scale).max())
    print(f"{name}: host loop {ia['num_attempted_steps']} attempts / {ia['num_steps']} steps in {ta*1e3:.0f} ms; device loop "
          f"{ib['num_attempted_steps']} / {ib['num_steps']} in {tb*1e3:.0f} ms; same counters: {same}; t {a.t:.6e} vs {b.t:.6e}; "
          f"max row-relative difference of the final mean {rel:.2e}", flush=True)
    del a, b
    torch.cuda.empty_cache()
