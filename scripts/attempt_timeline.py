"""Where does an attempt of the multi-attempt kernels spend its time?  Needs the instrumented library
(python -m tornadox_b200.build --variant=timeline -DTDX_TIMELINE; run with TDX_LIB=tornadox_b200/lib/libtornadox_b200_timeline.so).
CTA 0 stamps the globaltimer at the stations of every attempt; this prints the mean gaps over the attempts of one launch."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tornadox_b200 import _lib, ek0, ek1, ivp, step


def stamps(eng, n_att):
    buf = np.zeros(64 * 8)
    _lib.check(eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 2048, 64 * 8))
    return buf.reshape(64, 8)[: min(n_att, 64)]


def run(name, cls, prob, rule, nu, k=48):
    sol = cls(num_derivatives=nu, steprule=rule)
    state = sol.initialize(*prob)
    eng = sol.engine
    ring = sol._ctl_ring(state, 2, clone_first=True)
    dt = sol.steprule.first_dt(*prob) * (0.5 if isinstance(rule, step.AdaptiveSteps) else 1.0)
    for _ in range(2):
        eng.ctl_begin(eng.native_steprule(rule), state.t, dt, 1e300)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sol._ctl_launch(k, ring)
        b.record()
        torch.cuda.synchronize()
    ctl = eng.ctl_read()
    s = stamps(eng, int(ctl.attempts))
    ms = a.elapsed_time(b)
    print(f"{name}: {ctl.attempts} attempts ({ctl.accepted} accepted) in {ms * 1e3:.1f} us -> {ms * 1e3 / max(1, ctl.attempts):.2f} us/attempt")
    if s.shape[0] < 3:
        return
    if isinstance(sol, ek0.KroneckerEK0):
        labels = ["phase A", "mid barrier+gain", "phase B", "end barrier(+ctl)", "next attempt's top"]
        t = s[:, :5]
        gaps = np.diff(t, axis=1)
        top = t[1:, 0] - t[:-1, 4]
        for lab, g in zip(labels[:4], gaps.T):
            print(f"    {lab:22s} {np.mean(g[1:]) * 1e-3:7.2f} us (min {np.min(g[1:]) * 1e-3:.2f}, max {np.max(g[1:]) * 1e-3:.2f})")
        print(f"    {labels[4]:22s} {np.mean(top) * 1e-3:7.2f} us")
        qr = s[:, 6] - s[:, 5]
        print(f"    stacked QR (CTA 0 producer lane) {np.mean(qr[1:]) * 1e-3:.2f} us; starts {np.mean((s[:, 5] - s[:, 2])[1:]) * 1e-3:.2f} us after the gain")
    else:
        t = s[:, :3]
        gaps = np.diff(t, axis=1)
        top = t[1:, 0] - t[:-1, 2]
        print(f"    tiles                  {np.mean(gaps[1:, 0]) * 1e-3:7.2f} us")
        print(f"    end barrier(+ctl)      {np.mean(gaps[1:, 1]) * 1e-3:7.2f} us")
        print(f"    next attempt's top     {np.mean(top) * 1e-3:7.2f} us")
        last = s[1:-1]
        print(f"    (last CTA: arrives {np.mean(last[:, 3] - last[:, 1]) * 1e-3:.2f} us after CTA 0 ran out of tiles; reduction "
              f"{np.mean(last[:, 4] - last[:, 3]) * 1e-3:.2f} us; controller + release {np.mean(last[:, 5] - last[:, 4]) * 1e-3:.2f} us; "
              f"CTA 0 starts the next attempt {np.mean(s[2:, 0] - s[1:-1, 5]) * 1e-3:.2f} us after the release)")


y3 = np.random.default_rng(2).uniform(size=2 * 256 * 256)
run("config3 fhn 256^2 KroneckerEK0 nu=4 adaptive", ek0.KroneckerEK0, ivp.fhn_2d(dx=1 / 256, y0=y3), step.AdaptiveSteps(), 4)
run("        fhn 256^2 DiagonalEK1 nu=4 adaptive", ek1.DiagonalEK1, ivp.fhn_2d(dx=1 / 256, y0=y3), step.AdaptiveSteps(), 4)
run("config2 lorenz96 100k DiagonalEK1 nu=3 constant", ek1.DiagonalEK1, ivp.lorenz96(num_variables=100_000), step.ConstantSteps(0.01), 3)
run("        lorenz96 100k KroneckerEK0 nu=3 adaptive", ek0.KroneckerEK0, ivp.lorenz96(num_variables=100_000), step.AdaptiveSteps(), 3)
if len(sys.argv) > 1 and sys.argv[1] == "small":
    sys.exit(0)
y5 = np.random.default_rng(2).uniform(size=2 * 2048 * 2048)
run("config5 fhn 2048^2 KroneckerEK0 nu=4 adaptive", ek0.KroneckerEK0, ivp.fhn_2d(dx=1 / 2048, y0=y5), step.AdaptiveSteps(), 4, k=16)
run("config5 fhn 2048^2 DiagonalEK1 nu=4 adaptive", ek1.DiagonalEK1, ivp.fhn_2d(dx=1 / 2048, y0=y5), step.AdaptiveSteps(), 4, k=8)
