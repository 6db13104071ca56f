"""Kernel-level timing of every BASELINE.json configuration (development aid; bench.py is the contract)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine

def timeit(fn, warm=3, it=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it

g = torch.Generator(device="cuda").manual_seed(0)
CONFIGS = [
    ("config2 lorenz96 N=100k dek1 nu=3", ivp.lorenz96(num_variables=100_000), 3, "dek1"),
    ("config3 fhn 256^2 ek0 nu=4", ivp.fhn_2d(dx=1/256, y0=np.zeros(2)), 4, "ek0"),
    ("config4 brusselator N=1M dek1 nu=4", ivp.brusselator(N=1_000_000), 4, "dek1"),
    ("config5 fhn 2048^2 dek1 nu=4", ivp.fhn_2d(dx=1/2048, y0=np.zeros(2)), 4, "dek1"),
    ("config5 fhn 2048^2 ek0 nu=4", ivp.fhn_2d(dx=1/2048, y0=np.zeros(2)), 4, "ek0"),
    ("extra   lorenz96 N=8M dek1 nu=4", ivp.lorenz96(num_variables=8_000_000), 4, "dek1"),
    ("extra   lorenz96 N=8M ek0 nu=4", ivp.lorenz96(num_variables=8_000_000), 4, "ek0"),
    ("extra   brusselator N=4M ek0 nu=4", ivp.brusselator(N=4_000_000), 4, "ek0"),
    ("extra   fhn 2048^2 dek1 nu=3", ivp.fhn_2d(dx=1/2048, y0=np.zeros(2)), 3, "dek1"),
    ("extra   fhn 2048^2 dek1 nu=5", ivp.fhn_2d(dx=1/2048, y0=np.zeros(2)), 5, "dek1"),
]
for name, prob, nu, solver in CONFIGS:
    n = nu + 1
    spec = prob.spec
    d = spec.n_fields * spec.grid_rows * spec.grid_cols
    eng = StepEngine(spec, nu, dtype=torch.float64, device="cuda:0")
    mean = torch.rand((n, d), generator=g, device="cuda", dtype=torch.float64)
    mo = torch.empty_like(mean)
    if solver == "dek1":
        sc = 1e-3 * torch.randn((d, n, n), generator=g, device="cuda", dtype=torch.float64)
        so = torch.empty_like(sc)
        ms = timeit(lambda: eng.dek1_attempt(0.0, 1e-6, mean, sc, 1e-4, 1e-2, mean_out=mo, sc_out=so, want_err_ref=True))
        byts = 2 * (n + n * n) * 8 * d
        del sc, so
    else:
        Cl = 1e-3 * torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
        ms = timeit(lambda: eng.ek0_attempt(0.0, 1e-6, mean, Cl, 1e-4, 1e-2, mean_out=mo, want_ref=True))
        byts = 2 * n * 8 * d
    print(f"{name:40s} d={d:9d}: {ms*1e3:8.1f} us/attempt  {byts/ms/1e6:7.0f} GB/s algorithmic ({byts/ms/1e6/6553*100:5.1f}% of 6553)  {d/ms*1e3:.3e} dims*steps/s", flush=True)
    del mean, mo, eng
    torch.cuda.empty_cache()
