"""Per-launch timing of the sharded step under torchrun (development aid): pack vs step kernel."""
import os, sys, ctypes
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tornadox_b200 import _lib, ivp
from tornadox_b200.engine import StepEngine, _ptr

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
side = 2048; nu = 4; n = nu + 1
prob = ivp.fhn_2d(dx=1.0/side, y0=np.zeros(2))
eng = StepEngine(prob.spec, nu, device=dev, group=dist.group.WORLD, rank=rank, world=world)
d = eng.d_local
g = torch.Generator(device=dev).manual_seed(rank)
mean = torch.rand((n, d), generator=g, device=dev, dtype=torch.float64)
sc = 1e-3*torch.randn((d, n, n), generator=g, device=dev, dtype=torch.float64)
mo, so = torch.empty_like(mean), torch.empty_like(sc)
err = torch.empty(d, device=dev, dtype=torch.float64); ref = torch.empty_like(err)
sq = torch.zeros(1, device=dev, dtype=torch.float64)
st = eng._stream()
flags = _lib.TDX_STEP_COMM_HALO | _lib.TDX_STEP_COMM_REDUCE
args = _lib.StepArgs(0.0, 1e-6, 1e-4, 1e-2, flags, 0)
def attempt(ev=None):
    if ev: ev[0].record()
    _lib.check(eng.lib.tdx_comm_pack_halo(eng.handle, 1e-6, _ptr(mean), st))
    if ev: ev[1].record()
    _lib.check(eng.lib.tdx_dek1_attempt_step(eng.handle, ctypes.byref(args), _ptr(mean), _ptr(sc), _ptr(mo), _ptr(so), _ptr(err), _ptr(ref), None, None, _ptr(sq), st))
    if ev: ev[2].record()
for _ in range(5): attempt()
torch.cuda.synchronize(); dist.barrier()
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(30)]
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for ev in evs: attempt(ev)
t1.record()
torch.cuda.synchronize()
pack = np.median([e[0].elapsed_time(e[1]) for e in evs]); main = np.median([e[1].elapsed_time(e[2]) for e in evs])
print(f"rank {rank}: pack {pack*1e3:.1f} us, step kernel {main*1e3:.1f} us, total/attempt {t0.elapsed_time(t1)/30*1e3:.1f} us", flush=True)
# same shard size without any communication (interior flag => no halo reads, no reduce), for reference
args2 = _lib.StepArgs(0.0, 1e-6, 1e-4, 1e-2, _lib.TDX_STEP_TILES_INTERIOR, 0)
def plain():
    _lib.check(eng.lib.tdx_dek1_attempt_step(eng.handle, ctypes.byref(args2), _ptr(mean), _ptr(sc), _ptr(mo), _ptr(so), _ptr(err), _ptr(ref), None, None, _ptr(sq), st))
for _ in range(3): plain()
torch.cuda.synchronize()
t0.record()
for _ in range(30): plain()
t1.record(); torch.cuda.synchronize()
print(f"rank {rank}: interior-only kernel {t0.elapsed_time(t1)/30*1e3:.1f} us", flush=True)
# fused: one launch delivers the halos, does all tiles and all-reduces the norm
args3 = _lib.StepArgs(0.0, 1e-6, 1e-4, 1e-2, _lib.TDX_STEP_COMM_PACK | _lib.TDX_STEP_COMM_REDUCE, 0)
def fused():
    _lib.check(eng.lib.tdx_dek1_attempt_step(eng.handle, ctypes.byref(args3), _ptr(mean), _ptr(sc), _ptr(mo), _ptr(so), _ptr(err), _ptr(ref), None, None, _ptr(sq), st))
for _ in range(3): fused()
torch.cuda.synchronize(); dist.barrier()
t0.record()
for _ in range(30): fused()
t1.record(); torch.cuda.synchronize()
print(f"rank {rank}: fused pack+step+allreduce launch {t0.elapsed_time(t1)/30*1e3:.1f} us", flush=True)
eng.check_peers()
dist.barrier(); dist.destroy_process_group()
