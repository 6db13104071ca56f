"""Is the per-SM speed of the fused EK0 kernel reproducible across launches? (needs -DTDX_EK0_TIMING)"""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine

size = 2048
nu = 4; n = nu + 1
g = torch.Generator(device="cuda").manual_seed(0)
prob = ivp.fhn_2d(dx=1.0/size, y0=np.zeros(2))
d = 2*size*size
eng = StepEngine(prob.spec, nu, dtype=torch.float64, device="cuda:0")
mean = torch.rand((n, d), generator=g, device="cuda", dtype=torch.float64)
mo = torch.empty_like(mean)
Cl = 1e-3*torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
def run():
    eng.ek0_attempt(0.0, 1e-4, mean, Cl, 1e-4, 1e-2, mean_out=mo, want_ref=False)
    torch.cuda.synchronize()
    buf = np.zeros(4096 + 1024)
    rc = eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1024, 4096 + 1024)
    G = int(np.count_nonzero(buf[:4096].reshape(1024, 4)[:, 0]))
    t = buf[:4*G].reshape(G, 4).copy()
    smid = buf[4096:4096 + G].astype(int)
    return t, smid
for _ in range(3): run()
res = []
for _ in range(4):
    t, smid = run()
    dA = t[:, 1] - t[:, 0]; dB = t[:, 3] - t[:, 2]
    perA = np.zeros(148); perB = np.zeros(148); cnt = np.zeros(148)
    for b in range(len(smid)):
        perA[smid[b]] += dA[b]; perB[smid[b]] += dB[b]; cnt[smid[b]] += 1
    res.append((perA / cnt / 1e3, perB / cnt / 1e3))
    print("run: A mean %.1f max %.1f | B mean %.1f max %.1f | smid of block 0..7: %s" % (dA.mean()/1e3, dA.max()/1e3, dB.mean()/1e3, dB.max()/1e3, smid[:8]))
for i in range(1, 4):
    print("corr per-SM B run0 vs run%d: %.2f   A: %.2f" % (i, np.corrcoef(res[0][1], res[i][1])[0, 1], np.corrcoef(res[0][0], res[i][0])[0, 1]))
m = np.mean([r[1] for r in res], 0)
print("per-SM B duration averaged over runs: min %.1f max %.1f std %.1f" % (m.min(), m.max(), m.std()))
print("by smid//? groups of 2 (TPC):", np.round(m.reshape(74, 2).mean(1)[:20], 1))
print("sorted smids slow->fast:", np.argsort(-m)[:30])
