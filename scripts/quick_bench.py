"""Quick device-side timing of the step kernels (development aid; bench.py is the contract)."""
import sys, time, math
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine

def timeit(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts)//2], ts[0]

def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    nu = 4; n = nu + 1
    prob = ivp.fhn_2d(dx=1.0/size, y0=np.zeros(2))  # y0 unused here
    d = 2*size*size
    g = torch.Generator(device="cuda").manual_seed(0)
    for dtype in (torch.float64, torch.float32):
        eng = StepEngine(prob.spec, nu, dtype=dtype, device="cuda:0")
        mean = torch.rand((n, d), generator=g, device="cuda", dtype=dtype)
        sc = 1e-3*torch.randn((d, n, n), generator=g, device="cuda", dtype=dtype)
        mo = torch.empty_like(mean); so = torch.empty_like(sc)
        w = mean.element_size()
        for want in (True, False):
            med, best = timeit(lambda: eng.dek1_attempt(0.0, 1e-4, mean, sc, 1e-4, 1e-2, mean_out=mo, sc_out=so, want_err_ref=want))
            byts = 2*(n+n*n)*w*d
            print(f"dek1 {dtype} d={d} err_ref={want}: median {med:.3f} ms best {best:.3f} ms  -> {byts/med/1e6:.0f} GB/s algorithmic, {d/med*1e3:.3e} dims*steps/s")
        Cl = 1e-3*torch.randn((n, n), generator=g, device="cuda", dtype=dtype)
        for want in (True, False):
            med, best = timeit(lambda: eng.ek0_attempt(0.0, 1e-4, mean, Cl, 1e-4, 1e-2, mean_out=mo, want_ref=want))
            byts = 2*n*w*d
            print(f"ek0  {dtype} d={d} ref={want}: median {med:.3f} ms best {best:.3f} ms  -> {byts/med/1e6:.0f} GB/s algorithmic, {d/med*1e3:.3e} dims*steps/s")
        del mean, sc, mo, so
    # copy roofline on this box
    a = torch.empty(1<<30, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a))
    print(f"copy 2 GiB+2 GiB: best {best:.3f} ms -> {2*a.numel()*2/best/1e6:.0f} GB/s")

main()
