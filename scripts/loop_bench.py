"""Wall time per attempt of simulate_final_state: device-resident controller vs host-driven loop (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ek0, ek1, ivp, step

def run(cls, prob, rule, nu, device_loop, batch=64):
    sol = cls(num_derivatives=nu, steprule=rule)
    sol.device_loop = device_loop
    sol.device_loop_batch = batch
    sol.simulate_final_state(prob)     # warm-up: context, kernels, allocator
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, info = sol.simulate_final_state(prob)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    return sec, info

cases = []
y0 = np.random.default_rng(2).uniform(size=2 * 256 * 256)
cases.append(("config3 fhn 256^2 KroneckerEK0 nu=4 adaptive", ek0.KroneckerEK0, ivp.fhn_2d(dx=1/256, y0=y0, tmax=3e-3), step.AdaptiveSteps, 4))
cases.append(("        fhn 256^2 DiagonalEK1 nu=4 adaptive", ek1.DiagonalEK1, ivp.fhn_2d(dx=1/256, y0=y0, tmax=3e-3), step.AdaptiveSteps, 4))
cases.append(("config2 lorenz96 N=100k DiagonalEK1 nu=3 ConstantSteps(0.01) 1000 steps", ek1.DiagonalEK1,
              ivp.lorenz96(num_variables=100_000, tmax=10.0), lambda: step.ConstantSteps(0.01), 3))
y5 = np.random.default_rng(2).uniform(size=2 * 2048 * 2048)
cases.append(("config5 fhn 2048^2 DiagonalEK1 nu=4 adaptive", ek1.DiagonalEK1, ivp.fhn_2d(dx=1/2048, y0=y5, tmax=2e-6), step.AdaptiveSteps, 4))
cases.append(("config5 fhn 2048^2 KroneckerEK0 nu=4 adaptive", ek0.KroneckerEK0, ivp.fhn_2d(dx=1/2048, y0=y5, tmax=2e-6), step.AdaptiveSteps, 4))
for name, cls, prob, mk, nu in cases:
    for dl in (False, True):
        sec, info = run(cls, prob, mk(), nu, dl)
        att = info["num_attempted_steps"]
        print(f"{name:75s} {'device loop' if dl else 'host loop  '}: {sec*1e3:8.2f} ms, {att:5d} attempts / {info['num_steps']:5d} steps -> {sec/att*1e6:7.1f} us/attempt", flush=True)
