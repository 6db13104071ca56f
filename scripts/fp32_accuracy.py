"""fp32 kernels at the BASELINE sizes against the fp64 oracle (north_star: 1e-5 in fp32): one attempt from a mid-solve looking
state, DiagonalEK1 on windows, KroneckerEK0 in full.  Prints the relative errors per configuration and output array."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle.ref_np import filters as ofilters, problems as oproblems
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine
from test_gpu_fullsize import _fhn_window_problem

DEV = "cuda:0"

def nordsieck(n, dt):
    import math
    return np.array([dt ** (n - 1 - k + 0.5) / math.factorial(n - 1 - k) for k in range(n)])


def state(n, d, seed, dt, blocks=True):
    """A mid-solve looking state: derivative rows of order one, factor rows scaled like the Nordsieck vector of the step (what
    the update leaves behind, ek1.py:246-247) -- an unscaled random factor overflows fp32 in the preconditioned QR at small dt."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    mean = torch.randn((n, d), generator=g, device=DEV, dtype=torch.float64)
    mean[0] = torch.rand(d, generator=g, device=DEV, dtype=torch.float64)
    sc = None
    if blocks:
        p = torch.as_tensor(nordsieck(n, dt), device=DEV)
        sc = 1e-2 * torch.randn((d, n, n), generator=g, device=DEV, dtype=torch.float64) * p[None, :, None]
    return mean, sc

def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    assert a.shape == b.shape and a.size > 0, (a.shape, b.shape)
    assert np.all(np.isfinite(b)), "non-finite oracle values"
    if not np.all(np.isfinite(a)):
        return float("inf")        # the fp32 kernel left fp32's exponent range (reported as such below)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

def fhn_dek1(side, dt):
    nu, n, d, npts = 4, 5, 2 * side * side, side * side
    p = ivp.fhn_2d(dx=1.0 / side, y0=np.zeros(2))
    eng = StepEngine(p.spec, nu, dtype=torch.float32, device=DEV)
    mean, sc = state(n, d, 11, dt)
    m, s, err, ref, sq = eng.dek1_attempt(0.3, dt, mean.float(), sc.float(), 1e-4, 1e-2)
    out = {}
    for r0, r1 in ((0, 7), (side // 2 - 3, side // 2 + 5)):
        rows = r1 - r0
        idx = torch.cat([torch.arange(f * npts + r0 * side, f * npts + r1 * side, device=DEV) for f in range(2)])
        f, dfd = _fhn_window_problem(side, 1.0 / side, r0, r1)
        wm, wsc = mean[:, idx].float().double().cpu().numpy(), sc[idx].float().double().cpu().numpy()   # the fp32-rounded inputs
        ost, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, wm, wsc, None, None), dt, f, dfd, nu)
        lo, hi = (0 if r0 == 0 else 1), (rows if r1 == side else rows - 1)
        keep = np.concatenate([np.arange(fl * rows * side + lo * side, fl * rows * side + hi * side) for fl in range(2)])
        sel = idx.cpu().numpy()[keep]
        for name, got, want in (("mean", m[:, sel], ost.mean[:, keep]), ("cov_sqrtm", s[sel], ost.cov_sqrtm[keep]),
                                ("error_estimate", err[sel], ost.error_estimate[keep])):
            out[name] = max(out.get(name, 0.0), rel(got.cpu().numpy(), want))
        S = s[sel].double().cpu().numpy()
        out["covariance S S^T"] = max(out.get("covariance S S^T", 0.0), rel(S @ S.transpose(0, 2, 1), ost.cov_sqrtm[keep] @ ost.cov_sqrtm[keep].transpose(0, 2, 1)))
    return out

def fhn_ek0(side, dt):
    nu, n, d = 4, 5, 2 * side * side
    o = oproblems.fhn_2d(dx=1.0 / side, y0=np.zeros(d))
    p = ivp.fhn_2d(dx=1.0 / side, y0=np.zeros(2))
    eng = StepEngine(p.spec, nu, dtype=torch.float32, device=DEV)
    mean, _ = state(n, d, 12, dt, blocks=False)
    Cl = (1e-2 * np.random.default_rng(3).standard_normal((n, n)) * nordsieck(n, dt)[:, None]).astype(np.float32)
    m, C, err, ref, itol = eng.ek0_attempt(0.3, dt, mean.float(), torch.as_tensor(Cl, device=DEV), 1e-4, 1e-2)
    ost, _ = ofilters.kronecker_ek0_attempt(ofilters.FilterState(0.3, mean.float().double().cpu().numpy(), Cl.astype(np.float64), None, None), dt, o.f, nu)
    return {"mean": rel(m.cpu().numpy(), ost.mean), "cov_sqrtm_2": rel(C.cpu().numpy(), ost.cov_sqrtm),
            "error_estimate": abs(float(err) - ost.error_estimate) / abs(ost.error_estimate)}

def chain_dek1(kind, N, nu, dt):
    n = nu + 1
    if kind == "lorenz96":
        p, o, d = ivp.lorenz96(num_variables=N), oproblems.lorenz96(num_variables=N), N
    else:
        p, o, d = ivp.brusselator(N=N), None, 2 * N
    eng = StepEngine(p.spec, nu, dtype=torch.float32, device=DEV)
    mean, sc = state(n, d, 13, dt)
    if kind == "lorenz96":
        mean[0] = 8.0 + mean[0]
    else:
        mean[0, :N] += 1.0
        mean[0, N:] += 3.0
    m, s, err, ref, _ = eng.dek1_attempt(0.3, dt, mean.float(), sc.float(), 1e-4, 1e-2)
    out = {}
    if kind == "lorenz96":
        ost, _ = ofilters.diagonal_ek1_attempt(ofilters.FilterState(0.3, mean.float().double().cpu().numpy(), sc.float().double().cpu().numpy(), None, None),
                                               dt, o.f, o.df_diagonal, nu)
        out = {"mean": rel(m.cpu().numpy(), ost.mean), "cov_sqrtm": rel(s.cpu().numpy(), ost.cov_sqrtm),
               "error_estimate": rel(err.cpu().numpy(), ost.error_estimate)}
    else:
        const = (N + 1) ** 2 / 50.0
        for i0, i1 in ((0, 300), (N // 2 - 129, N // 2 + 129)):
            w = i1 - i0
            ow = oproblems.brusselator(N=w)
            idx = torch.cat([torch.arange(f * N + i0, f * N + i1, device=DEV) for f in range(2)])
            ost, _ = ofilters.diagonal_ek1_attempt(
                ofilters.FilterState(0.3, mean[:, idx].float().double().cpu().numpy(), sc[idx].float().double().cpu().numpy(), None, None), dt,
                lambda t, y: ow.f(t, y, c=const), lambda t, y: ow.df_diagonal(t, y, c=const), nu)
            lo, hi = (0 if i0 == 0 else 1), (w if i1 == N else w - 1)
            keep = np.concatenate([np.arange(fl * w + lo, fl * w + hi) for fl in range(2)])
            sel = idx.cpu().numpy()[keep]
            for name, got, want in (("mean", m[:, sel], ost.mean[:, keep]), ("cov_sqrtm", s[sel], ost.cov_sqrtm[keep]),
                                    ("error_estimate", err[sel], ost.error_estimate[keep])):
                out[name] = max(out.get(name, 0.0), rel(got.cpu().numpy(), want))
    return out

for name, fn in (("config2 lorenz96 N=100k DiagonalEK1 nu=3 dt=0.01", lambda: chain_dek1("lorenz96", 100_000, 3, 0.01)),
                 ("config3 fhn 256^2 KroneckerEK0 nu=4 dt=2e-5", lambda: fhn_ek0(256, 2e-5)),
                 ("config3 fhn 256^2 DiagonalEK1 nu=4 dt=2e-5", lambda: fhn_dek1(256, 2e-5)),
                 ("config4 brusselator N=1M DiagonalEK1 nu=4 dt=1e-9", lambda: chain_dek1("brusselator", 1_000_000, 4, 1e-9)),
                 ("config5 fhn 2048^2 KroneckerEK0 nu=4 dt=2e-7", lambda: fhn_ek0(2048, 2e-7)),
                 ("config5 fhn 2048^2 DiagonalEK1 nu=4 dt=2e-7", lambda: fhn_dek1(2048, 2e-7))):
    res = fn()
    dt_case = float(name.split("dt=")[1])
    nu_case = int(name.split("nu=")[1].split()[0])
    pinv0 = 1.0 / nordsieck(nu_case + 1, dt_case)[0]
    note = f"   [p_inv[0] = {pinv0:.1e}" + (" > fp32 max 3.4e38: the Nordsieck preconditioner itself is not representable]" if pinv0 > 3.4e38 else "]")
    print(f"{name:52s} " + "  ".join(f"{k}: {v:.2e}" for k, v in res.items()) + note, flush=True)
    torch.cuda.empty_cache()
