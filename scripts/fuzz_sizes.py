"""Parity fuzz over awkward fhn grid sizes (ragged / multiple column blocks of the strip kernel, odd sizes -> non-bulk path),
chain lengths around tile boundaries, both solvers, against the NumPy oracle (development aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import make_pair, random_state
from oracle.ref_np import filters as ofilters
from tornadox_b200.engine import StepEngine

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

NU = int(sys.argv[1]) if len(sys.argv) > 1 else 4
worst = 0.0
cases = [("fhn_2d", s) for s in (255, 256, 257, 258, 300, 511, 512, 513, 600)] + \
        [("lorenz96", s) for s in (255, 256, 257, 511, 513, 70001, 70002)] + \
        [("brusselator", s) for s in (127, 128, 129, 30001, 30002)] + [("burgers_1d", s) for s in (257, 513, 40001)] + \
        [("wave_1d", s) for s in (129, 257, 20002)]
for kind, size in cases:
    o, p, _ = make_pair(kind, size)
    nu = NU; n = nu + 1; d = o.y0.shape[0]
    rng = np.random.default_rng(size)
    mean, sc = random_state(rng, n, d, "dense")
    mean[0] = o.y0 + 0.05 * rng.standard_normal(d)
    Cl = 1e-2 * rng.standard_normal((n, n))
    dt = 1e-4 if kind == "fhn_2d" else 1e-3
    eng = StepEngine(p.spec, nu, device="cuda:0")
    ost, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(0.3, mean, Cl, None, None), dt, o.f, nu)
    m, C, err, ref, itol = eng.ek0_attempt(0.3, dt, torch.as_tensor(mean, device="cuda"), torch.as_tensor(Cl, device="cuda"), 1e-4, 1e-2)
    e0 = max(rel(m.cpu().numpy(), ost.mean), rel(C.cpu().numpy(), ost.cov_sqrtm), rel(ref.cpu().numpy(), ost.reference_state))
    tol_sum = float(np.sum((dt * ost.error_estimate / (1e-4 + 1e-2 * ost.reference_state)) ** 2))
    # the norm sum is dominated by coordinates with |reference_state| ~ 0, where tol ~ abstol: an absolute difference delta in
    # reference_state changes a term by the relative amount 2 reltol delta / abstol
    delta = float(np.max(np.abs(ref.cpu().numpy() - ost.reference_state)))
    e_norm = abs(float(itol) - tol_sum) / tol_sum
    assert e_norm < 1e-10 + 4 * 1e-2 * delta / 1e-4, (kind, size, "norm sum", e_norm, delta)
    e1 = 0.0
    if d <= 200_000:
        ost1, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, mean, sc, None, None), dt, o.f, o.df_diagonal, nu)
        m1, s1, err1, ref1, sq1 = eng.dek1_attempt(0.3, dt, torch.as_tensor(mean, device="cuda"), torch.as_tensor(sc, device="cuda"), 1e-4, 1e-2)
        e1 = max(rel(m1.cpu().numpy(), ost1.mean), rel(s1.cpu().numpy(), ost1.cov_sqrtm), rel(err1.cpu().numpy(), ost1.error_estimate))
    worst = max(worst, e0, e1)
    print(f"{kind:12s} {size:6d}: ek0 {e0:.2e}  dek1 {e1:.2e}", flush=True)
    assert e0 < 1e-10 and e1 < 1e-10, (kind, size, e0, e1)
print("FUZZ_OK worst", worst)
