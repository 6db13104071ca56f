"""DiagonalEK1 kernel time for nu = 1..5 at 8.4M dimensions (development aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine
g = torch.Generator(device="cuda").manual_seed(0)
size = 2048; d = 2 * size * size
for nu in ([int(a) for a in sys.argv[1:]] or (1, 2, 3, 4, 5)):
    n = nu + 1
    eng = StepEngine(ivp.fhn_2d(dx=1/size, y0=np.zeros(2)).spec, nu, device="cuda:0")
    mean = torch.rand((n, d), generator=g, device="cuda", dtype=torch.float64)
    sc = 1e-3 * torch.randn((d, n, n), generator=g, device="cuda", dtype=torch.float64)
    mo, so = torch.empty_like(mean), torch.empty_like(sc)
    f = lambda: eng.dek1_attempt(0.0, 1e-7, mean, sc, 1e-4, 1e-2, mean_out=mo, sc_out=so, want_err_ref=True)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    byts = 2 * (n + n * n) * 8 * d
    print(f"nu={nu}: {ms*1e3:8.1f} us  {byts/ms/1e6:6.0f} GB/s algorithmic = {byts/ms/1e6/6553*100:5.1f}% of 6553", flush=True)
    del mean, sc, mo, so, eng
