"""In-kernel timeline of the fused KroneckerEK0 launch (needs a library built with -DTDX_EK0_TIMING)."""
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
mo = torch.empty_like(mean)
Cl = 1e-3*torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
for _ in range(4):
    eng.ek0_attempt(0.0, 1e-4, mean, Cl, 1e-4, 1e-2, mean_out=mo, want_ref=False)
torch.cuda.synchronize()
buf = np.zeros(4096 + 1024)
rc = eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1024, 4096 + 1024)
assert rc == 0
G = int(np.count_nonzero(buf[:4096].reshape(1024, 4)[:, 0]))
print("grid", G)
t = buf[:4*G].reshape(G, 4).copy()
smid = buf[4096:4096 + G].astype(int)
t0 = t[:, 0].min()
t = (t - t0) / 1e3
print("start   : min %.1f max %.1f us" % (t[:, 0].min(), t[:, 0].max()))
print("end A   : min %.1f mean %.1f max %.1f us" % (t[:, 1].min(), t[:, 1].mean(), t[:, 1].max()))
print("gain    : min %.1f max %.1f us" % (t[:, 2].min(), t[:, 2].max()))
print("end B   : min %.1f mean %.1f max %.1f us" % (t[:, 3].min(), t[:, 3].mean(), t[:, 3].max()))

dB = t[:, 3] - t[:, 2]
dA = t[:, 1] - t[:, 0]
print("phase B duration by column block:", np.round([dB[np.arange(G) % 8 == c].mean() for c in range(8)], 1))
print("phase B duration by chunk (first 10 / last 10):", np.round([dB[np.arange(G) // 8 == c].mean() for c in range(37)], 1))
order = np.argsort(smid)
pairs = {}
for b in range(G):
    pairs.setdefault(smid[b], []).append(b)
per_sm = np.array([[dB[b] for b in v] + [np.nan] * (2 - len(v)) for k, v in sorted(pairs.items())])
print("per-SM phase-B durations (pairs), first 20 SMs:\n", np.round(per_sm[:20], 1))
print("per-SM mean: min %.1f max %.1f ; within-SM diff mean %.1f" % (np.nanmean(per_sm, 1).min(), np.nanmean(per_sm, 1).max(), np.nanmean(np.abs(per_sm[:, 0] - per_sm[:, 1]))))
print("corr(dA, dB) = %.2f" % np.corrcoef(dA, dB)[0, 1])
sm_mean = np.nanmean(per_sm, 1)
print("slowest SMs:", [(k, round(m, 1)) for k, m in sorted(zip(sorted(pairs.keys()), sm_mean), key=lambda x: -x[1])[:12]])
print("fastest SMs:", [(k, round(m, 1)) for k, m in sorted(zip(sorted(pairs.keys()), sm_mean), key=lambda x: x[1])[:12]])
