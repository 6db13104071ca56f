"""Per-output discrepancy of one KroneckerEK0 attempt against the oracle (development aid): kind size nu."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import make_pair, random_state
from oracle.ref_np import filters as ofilters
from tornadox_b200.engine import StepEngine

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

kind, size, nu = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
o, p, _ = make_pair(kind, size)
n = nu + 1; d = o.y0.shape[0]
rng = np.random.default_rng(size)
mean, sc = random_state(rng, n, d, "dense")
mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
Cl = 1e-2 * rng.standard_normal((n, n))
dt = 1e-4 if kind == "fhn_2d" else 1e-3
eng = StepEngine(p.spec, nu, device="cuda:0")
ost, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(0.3, mean, Cl, None, None), dt, o.f, nu)
m, C, err, ref, itol = eng.ek0_attempt(0.3, dt, torch.as_tensor(mean, device="cuda"), torch.as_tensor(Cl, device="cuda"), 1e-4, 1e-2)
m = m.cpu().numpy()
print(kind, size, nu, "mean rows:", [rel(m[r], ost.mean[r]) for r in range(n)], "| mean (global max-norm)", rel(m, ost.mean))
print("  C", rel(C.cpu().numpy(), ost.cov_sqrtm), "err", abs(float(err) - ost.error_estimate) / ost.error_estimate,
      "ref", rel(ref.cpu().numpy(), ost.reference_state))
print("  C gpu\n", C.cpu().numpy(), "\n  C oracle\n", ost.cov_sqrtm)
i = int(np.argmax(np.abs(m - ost.mean).max(0)))
print("  worst column", i, "of", d, "gpu", m[:, i], "oracle", ost.mean[:, i])
