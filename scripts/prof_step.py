"""Short run of the step kernels at a given size for ncu (development aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tornadox_b200 import ivp
from tornadox_b200.engine import StepEngine

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
which = sys.argv[2] if len(sys.argv) > 2 else "both"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dtype = {"f64": torch.float64, "f32": torch.float32}[sys.argv[4] if len(sys.argv) > 4 else "f64"]
nu = int(sys.argv[5]) if len(sys.argv) > 5 else 4; n = nu + 1
prob = ivp.fhn_2d(dx=1.0/size, y0=np.zeros(2))
d = 2*size*size
g = torch.Generator(device="cuda").manual_seed(0)
eng = StepEngine(prob.spec, nu, dtype=dtype, device="cuda:0")
mean = torch.rand((n, d), generator=g, device="cuda", dtype=dtype)
mo = torch.empty_like(mean)
if which in ("both", "dek1"):
    sc = 1e-3*torch.randn((d, n, n), generator=g, device="cuda", dtype=dtype)
    so = torch.empty_like(sc)
    for _ in range(iters):
        eng.dek1_attempt(0.0, 1e-4, mean, sc, 1e-4, 1e-2, mean_out=mo, sc_out=so, want_err_ref=True)
if which in ("both", "ek0", "ek0noref"):
    # ek0noref: without the materialised reference_state vector (the solver's default: |mean[0]| is deferred), which is what
    # bench.py's timed launches do -- the capture profiles/traffic.json records
    Cl = 1e-3*torch.randn((n, n), generator=g, device="cuda", dtype=dtype)
    for _ in range(iters):
        eng.ek0_attempt(0.0, 1e-4, mean, Cl, 1e-4, 1e-2, mean_out=mo, want_ref=(which != "ek0noref"))
if which in ("ek1ctl", "ek0ctl"):
    # the controller-driven (multi-attempt) instantiations: one launch of 2 attempts per iteration
    from tornadox_b200 import step
    rule = eng.native_steprule(step.AdaptiveSteps())
    means = [mean, mo]
    if which == "ek1ctl":
        sc = 1e-3*torch.randn((d, n, n), generator=g, device="cuda", dtype=dtype)
        facs = [sc, torch.empty_like(sc)]
    else:
        Cl = 1e-3*torch.randn((n, n), generator=g, device="cuda", dtype=dtype)
        facs = [Cl, torch.empty_like(Cl)]
        errs = [torch.zeros((), device="cuda", dtype=dtype) for _ in range(2)]
    for _ in range(iters):
        eng.ctl_begin(rule, 0.0, 1e-4, 1.0, nbuf=2)
        if which == "ek1ctl":
            eng.dek1_run_attempts(2, means, facs)
        else:
            eng.ek0_run_attempts(2, means, facs, errs)
torch.cuda.synchronize()
print("ok")
