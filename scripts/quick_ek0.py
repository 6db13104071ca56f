"""Device-side timing of the KroneckerEK0 attempt only (development aid; bench.py is the contract)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine

def timeit(fn, warm=3, it=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts)//2], ts[0]

sizes = [int(x) for x in sys.argv[1:]] or [2048]
nu = 4; n = nu + 1
g = torch.Generator(device="cuda").manual_seed(0)
for size in sizes:
    prob = ivp.fhn_2d(dx=1.0/size, y0=np.zeros(2))
    d = 2*size*size
    for dtype in (torch.float64, torch.float32):
        eng = StepEngine(prob.spec, nu, dtype=dtype, device="cuda:0")
        mean = torch.rand((n, d), generator=g, device="cuda", dtype=dtype)
        mo = torch.empty_like(mean)
        Cl = 1e-3*torch.randn((n, n), generator=g, device="cuda", dtype=dtype)
        w = mean.element_size()
        bufs = [mean, mo]
        def chained(want):
            eng.ek0_attempt(0.0, 1e-7, bufs[0], Cl, 1e-4, 1e-2, mean_out=bufs[1], want_ref=want)
            bufs.reverse()
        for want in (True, False):
            med, best = timeit(lambda: chained(want))
            byts = 2*n*w*d
            print(f"ek0 {size} {dtype} ref={want}: median {med*1e3:.1f} us best {best*1e3:.1f} us -> {byts/med/1e6:.0f} GB/s algorithmic, {d/med*1e3:.3e} dims*steps/s", flush=True)
