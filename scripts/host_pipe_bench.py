"""Timing of the host-buffer pipeline for several chunk counts (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine
size, nu = 2048, 4
n, d = nu + 1, 2 * size * size
eng = StepEngine(ivp.fhn_2d(dx=1.0/size, y0=np.zeros(2)).spec, nu, device="cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
mean = torch.rand((n, d), generator=g, device="cuda", dtype=torch.float64)
sc = 1e-3 * torch.randn((d, n, n), generator=g, device="cuda", dtype=torch.float64)
pin = lambda x: torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x)
h_mean, h_sc = pin(mean), pin(sc)
mo, so = torch.empty_like(h_mean, pin_memory=True), torch.empty_like(h_sc, pin_memory=True)
del mean, sc
for want in (True, False):
    for nchunks in (8, 16, 32, 64, 128):
        for _ in range(2):
            eng.dek1_attempt_host(0.0, 1e-7, h_mean, h_sc, 1e-4, 1e-2, mean_out=mo, sc_out=so, want_err_ref=want, nchunks=nchunks)
        t0 = time.perf_counter()
        for _ in range(4):
            eng.dek1_attempt_host(0.0, 1e-7, h_mean, h_sc, 1e-4, 1e-2, mean_out=mo, sc_out=so, want_err_ref=want, nchunks=nchunks)
        ms = (time.perf_counter() - t0) / 4 * 1e3
        print(f"err/ref={want} nchunks={nchunks:4d}: {ms:.2f} ms/step  {d/ms*1e3:.3e} dims*steps/s  ({2*(n+n*n)*8*d/ms/1e6:.1f} GB/s state bytes both ways)", flush=True)
