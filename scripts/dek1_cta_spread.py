"""When do the CTAs of a DiagonalEK1 attempt run out of tiles?  One launch per attempt against the multi-attempt kernel.
Needs the instrumented library (python -m tornadox_b200.build --variant=timeline -DTDX_TIMELINE; TDX_LIB=.../libtornadox_b200_timeline.so)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tornadox_b200 import _lib, ek1, ivp, step


def read(eng, off, cnt):
    buf = np.zeros(cnt)
    _lib.check(eng.lib.tdx_debug_read_workspace(eng.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), off, cnt))
    return buf


def spread(eng, t0, label, grid=296):
    raw = read(eng, 1024, 2 * grid).reshape(grid, 2)
    end, sm = (raw[:, 0] - t0) * 1e-3, raw[:, 1].astype(int)
    q = np.percentile(end, [0, 10, 50, 90, 100])
    print(f"{label}: CTAs run out of tiles after min {q[0]:.1f} / p10 {q[1]:.1f} / median {q[2]:.1f} / p90 {q[3]:.1f} / max {q[4]:.1f} us")
    # the two CTAs of an SM
    pairs = {}
    for e, s_ in zip(end, sm):
        pairs.setdefault(s_, []).append(e)
    both = np.array([sorted(v) for v in pairs.values() if len(v) == 2])
    if len(both):
        print(f"    per SM: earlier CTA mean {both[:, 0].mean():.1f}, later CTA mean {both[:, 1].mean():.1f} us; SMs with 2 CTAs: {len(both)}; "
              f"slowest SMs: {sorted(pairs, key=lambda k: -max(pairs[k]))[:6]}")
    idx = np.argsort(end)
    print(f"    first CTAs to finish {idx[:6].tolist()}, last {idx[-6:].tolist()}")
    extra = read(eng, 5000, 2 * grid).reshape(grid, 2)
    began = (extra[:, 0] - t0) * 1e-3
    print(f"    CTAs start on their tiles after min {began.min():.2f} / median {np.median(began):.2f} / max {began.max():.2f} us")
    q = (read(eng, 5600, 4 * grid).reshape(grid, 4)[:, :3] - t0) * 1e-3
    print(f"    median CTA reaches tile 28 / 56 / 84 after {np.median(q[:, 0]):.1f} / {np.median(q[:, 1]):.1f} / {np.median(q[:, 2]):.1f} us")
    for b in idx[-4:][::-1]:
        mates = [int(o) for o in np.where(sm == sm[b])[0] if o != b]
        print(f"      CTA {int(b):3d} reaches tile 28 / 56 / 84 after {q[b, 0]:.1f} / {q[b, 1]:.1f} / {q[b, 2]:.1f} us")
        print(f"      CTA {int(b):3d} on SM {int(sm[b]):3d}: began {began[b]:7.2f}, ran out of tiles {end[b]:7.2f} us, {int(extra[b, 1])} tiles; "
              f"SM mate {mates} ended {[round(float(end[m]), 1) for m in mates]}")
    return raw[:, 0]


def epilogue(eng, tiles_end, grid=296):
    """controller kernel only: per CTA, after the tile loop: bulk stores landed -> proxy fence -> barriers + last tile partial"""
    ep = read(eng, 3072, 4 * grid).reshape(grid, 4)
    for lab, col, base in (("bulk stores landed (cp.async.bulk.wait_group 0)", 0, tiles_end), ("fence.proxy.async.global", 1, ep[:, 0]),
                           ("2 barriers + the last tile's partial", 2, ep[:, 1]),
                           ("arrival (barrier, __threadfence, ticket atomic, barrier)", 3, ep[:, 2])):
        g = (ep[:, col] - base) * 1e-3
        print(f"    {lab:58s}: mean {g.mean():6.2f} / max {g.max():6.2f} us")
    print(f"    last CTA ready to report in {((ep[:, 2].max() - tiles_end.min()) * 1e-3):.2f} us after the FIRST CTA ran out of tiles")


size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
y0 = np.random.default_rng(2).uniform(size=2 * size * size)
prob = ivp.fhn_2d(dx=1.0 / size, y0=y0, tmax=1.0)
sol = ek1.DiagonalEK1(num_derivatives=4, steprule=step.AdaptiveSteps())
state = sol.initialize(*prob)
eng = sol.engine
dt = 0.5 * sol.steprule.first_dt(*prob)
# one launch per attempt (the plain instantiation)
for _ in range(3):
    new, _ = sol.attempt_step(state, dt, *prob)
torch.cuda.synchronize()
t0 = read(eng, 2048, 8)[0]
spread(eng, t0, "one launch per attempt       ")
# the multi-attempt kernel
ring = sol._ctl_ring(state, 2, clone_first=True)
for k in (1, 4, 4, 4, 4, 4):
    eng.ctl_begin(eng.native_steprule(sol.steprule), state.t, dt, 1e300)
    sol._ctl_launch(k, ring)
    torch.cuda.synchronize()
    ctl = eng.ctl_read()
    st = read(eng, 2048, 64 * 8).reshape(64, 8)
    t0 = st[int(ctl.attempts) - 1, 0]
    ends = spread(eng, t0, f"controller kernel, attempt {int(ctl.attempts)} of {k}")
    epilogue(eng, ends)
