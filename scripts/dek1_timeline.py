"""Spread of the CTAs' finish times in the DiagonalEK1 kernel (needs a library built with -DTDX_DEK1_TIMING)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nu = 4; n = nu + 1
g = torch.Generator(device="cuda").manual_seed(0)
prob = ivp.fhn_2d(dx=1.0/size, y0=np.zeros(2))
d = 2*size*size
eng = StepEngine(prob.spec, nu, dtype=torch.float64, device="cuda:0")
mean = torch.rand((n, d), generator=g, device="cuda", dtype=torch.float64)
sc = 1e-3*torch.randn((d, n, n), generator=g, device="cuda", dtype=torch.float64)
mo, so = torch.empty_like(mean), torch.empty_like(sc)
for rep in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    eng.dek1_attempt(0.0, 1e-7, mean, sc, 1e-4, 1e-2, mean_out=mo, sc_out=so, want_err_ref=False)
    b.record()
    torch.cuda.synchronize()
    buf = np.zeros(1024)
    rc = eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1024, 1024)
    assert rc == 0
    t = buf.reshape(512, 2)
    G = int(np.count_nonzero(t[:, 0]))
    end = t[:G, 0]
    end = (end - end.min()) / 1e3
    print(f"run {rep}: grid {G}, kernel {a.elapsed_time(b)*1e3:.1f} us; CTA finish times relative to the first finisher: mean {end.mean():.1f} us, max {end.max():.1f} us, p90 {np.percentile(end, 90):.1f} us")
