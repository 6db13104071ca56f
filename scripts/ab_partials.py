"""A/B: DiagonalEK1 with one partial per tile (sharding-independent norm) against per-CTA partials (TDX_CTA_PARTIALS=1)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine

def timeit(fn, warm=3, it=30):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it

g = torch.Generator(device="cuda").manual_seed(0)
for name, prob, nu, solver in [("fhn 2048^2 dek1 nu=4", ivp.fhn_2d(dx=1/2048, y0=np.zeros(2)), 4, "dek1"),
                               ("fhn 2048^2 dek1 nu=4 const", ivp.fhn_2d(dx=1/2048, y0=np.zeros(2)), 4, "dek1c"),
                               ("fhn 2048^2 ek0 nu=4", ivp.fhn_2d(dx=1/2048, y0=np.zeros(2)), 4, "ek0"),
                               ("brusselator 1M dek1 nu=4", ivp.brusselator(N=1_000_000), 4, "dek1"),
                               ("lorenz96 8M ek0 nu=4", ivp.lorenz96(num_variables=8_000_000), 4, "ek0")]:
    n = nu + 1
    spec = prob.spec
    d = spec.n_fields * spec.grid_rows * spec.grid_cols
    eng = StepEngine(spec, nu, dtype=torch.float64, device="cuda:0")
    mean = torch.rand((n, d), generator=g, device="cuda", dtype=torch.float64)
    mo = torch.empty_like(mean)
    if solver.startswith("dek1"):
        sc = 1e-3 * torch.randn((d, n, n), generator=g, device="cuda", dtype=torch.float64)
        so = torch.empty_like(sc)
        tol = (0.0, 0.0) if solver == "dek1c" else (1e-4, 1e-2)
        ms = timeit(lambda: eng.dek1_attempt(0.0, 1e-6, mean, sc, tol[0], tol[1], mean_out=mo, sc_out=so, want_err_ref=False))
        byts = 2 * (n + n * n) * 8 * d
        del sc, so
    else:
        Cl = 1e-3 * torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
        ms = timeit(lambda: eng.ek0_attempt(0.0, 1e-6, mean, Cl, 1e-4, 1e-2, mean_out=mo, want_ref=False))
        byts = 2 * n * 8 * d
    print(f"{name:32s}: {ms*1e3:8.1f} us/attempt  {byts/ms/1e6/6553*100:5.1f}% of 6553 GB/s", flush=True)
    del mean, mo, eng
    torch.cuda.empty_cache()
