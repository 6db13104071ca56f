"""Which CTAs of the fused EK0 kernel are slow: position of the strip, walking direction, SM?  (needs -DTDX_EK0_TIMING)"""
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
    eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1024, 4096 + 1024)
    G = int(np.count_nonzero(buf[:4096].reshape(1024, 4)[:, 0]))
    return buf[:4*G].reshape(G, 4).copy(), buf[4096:4096 + G].astype(int)
for _ in range(3): run()
runs = [run() for _ in range(6)]
G = len(runs[0][1])
A = np.array([(t[:, 1] - t[:, 0]) / 1e3 for t, _ in runs]); B = np.array([(t[:, 3] - t[:, 2]) / 1e3 for t, _ in runs])
print("G", G, " A mean %.1f max %.1f | B mean %.1f max %.1f" % (A.mean(), A.max(1).mean(), B.mean(), B.max(1).mean()))
print("per-block reproducibility (corr run0 vs others)  A:", np.round([np.corrcoef(A[0], A[i])[0, 1] for i in range(1, 6)], 2),
      " B:", np.round([np.corrcoef(B[0], B[i])[0, 1] for i in range(1, 6)], 2))
mA, mB = A.mean(0), B.mean(0)
print("run-averaged per-block: A min %.1f max %.1f std %.2f | B min %.1f max %.1f std %.2f   (single-run std A %.2f B %.2f)"
      % (mA.min(), mA.max(), mA.std(), mB.min(), mB.max(), mB.std(), A.std(1).mean(), B.std(1).mean()))
col_blocks = size // 256
chunk, cb = np.arange(G) // col_blocks, np.arange(G) % col_blocks
print("B by column block:", np.round([mB[cb == k].mean() for k in range(col_blocks)], 1))
print("A by column block:", np.round([mA[cb == k].mean() for k in range(col_blocks)], 1))
print("B by chunk (row band):", np.round([mB[chunk == k].mean() for k in range(G // col_blocks)], 1))
print("A by chunk (row band):", np.round([mA[chunk == k].mean() for k in range(G // col_blocks)], 1))
smid = runs[0][1]
same_sm = all((r[1] == smid).all() for r in runs)
print("block -> SM mapping identical across runs:", same_sm)
first = np.zeros(G, bool); seen = set()
for b in range(G):
    if smid[b] not in seen: first[b] = True; seen.add(smid[b])
print("B first CTA of its SM %.1f, second %.1f | A %.1f %.1f" % (mB[first].mean(), mB[~first].mean(), mA[first].mean(), mA[~first].mean()))
perB = np.array([mB[smid == s].mean() for s in range(148)])
print("per-SM B: std %.2f; by GPC-ish groups of 18 SMs:" % perB.std(), np.round([perB[i:i+18].mean() for i in range(0, 148, 18)], 1))
# the partner effect: does a CTA finish late when its SM partner does?
partner = {}
for b in range(G): partner.setdefault(smid[b], []).append(b)
pairs = np.array([v for v in partner.values() if len(v) == 2])
print("corr of B between the two CTAs of an SM: %.2f" % np.corrcoef(mB[pairs[:, 0]], mB[pairs[:, 1]])[0, 1])
slow = np.argsort(-mB)[:16]
print("slowest blocks (id, chunk, cb, smid, B):", [(int(b), int(chunk[b]), int(cb[b]), int(smid[b]), round(float(mB[b]), 1)) for b in slow])
