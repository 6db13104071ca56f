"""Step engine: owns one ``tdx_ctx`` (one GPU, one shard) and launches the fused kernels on torch buffers.

Multi-GPU: one process per GPU; the state dimension is sharded by grid rows (fhn_2d) or chain ranges
(lorenz96, brusselator) of BOTH fields.  Per attempt the only communication is the stencil halo and the
all-reduce of one or two fp64 sums; by default both travel through NVLink peer memory written and read by the
step kernels themselves (``_connect_peers``: CUDA-IPC mapped exchange blocks, ``tdx_comm_*``), so every rank
takes the identical accept/reject decision without a single NCCL call between kernels.  ``HaloExchanger``
(torch.distributed send/recv, any backend) remains for callers without peer access and for the CPU tests
(reference: no counterpart, single device).
"""

import ctypes
from collections import namedtuple

import numpy as np
import torch

from . import _lib

KernelSpec = namedtuple("KernelSpec", "kind params grid_rows grid_cols n_fields name")
"""Which in-kernel vector field an ``InitialValueProblem`` maps to (global geometry)."""


_RED_CHUNKS = 16      # kRedChunks in csrc/tdx_device.cuh


def _ek0_row_block(global_rows, col_blocks):
    """Row-block height of the fused fhn_2d KroneckerEK0 kernel's sums (ek0_row_block in csrc/tdx_ek0_fused.cuh)."""
    return 8 if -(-global_rows // 8) * col_blocks >= 256 else 1


def _chunk_aligned_bounds(spec, world):
    """Shard boundaries on the chunk grid of the kernels' reductions (see ``reduction_is_sharding_independent``), or None
    if the problem's size does not allow it.  The shards then differ in size by at most one chunk's rounding."""
    if _RED_CHUNKS % world:
        return None
    per = _RED_CHUNKS // world
    if spec.kind == _lib.TDX_IVP_FHN2D:
        rows = spec.grid_rows
        bounds = [((r * per) * rows) // _RED_CHUNKS for r in range(world + 1)]
        tiles_x = -(-spec.grid_cols // 32)
        tiles_y = -(-rows // 4)
        col_blocks = -(-spec.grid_cols // 256)
        rb = _ek0_row_block(rows, col_blocks)
        blocks_y = -(-rows // rb)
        for r, b in enumerate(bounds[1:-1], start=1):
            if b % 4 or (b // 4) * tiles_x != ((r * per) * tiles_y * tiles_x) // _RED_CHUNKS:
                return None
            if b % rb or (b // rb) * col_blocks != ((r * per) * blocks_y * col_blocks) // _RED_CHUNKS:
                return None
    else:
        tile = 256 if spec.n_fields == 1 else 128
        units = -(-spec.grid_cols // tile)
        bounds = [min((((r * per) * units) // _RED_CHUNKS) * tile, spec.grid_cols) for r in range(world + 1)]
        bounds[-1] = spec.grid_cols
    if any(b1 - b0 < 2 for b0, b1 in zip(bounds[:-1], bounds[1:])):
        return None
    return bounds


def shard_range(spec, rank, world):
    """Contiguous split of the sharded axis (grid rows for fhn_2d, points for 1-d chains).  Where the size allows it the
    boundaries are put on the chunk grid of the kernels' reductions, which makes the error norm (and with it the accepted-step
    sequence) bit-identical to a single-GPU run; otherwise the split is balanced to within one row / point."""
    extent = spec.grid_rows if spec.kind == _lib.TDX_IVP_FHN2D else spec.grid_cols
    if world == 1:
        return 0, extent
    if spec.kind == _lib.TDX_IVP_VANDERPOL:
        raise ValueError("vanderpol (d = 2) cannot be sharded across GPUs")
    if spec.kind in (_lib.TDX_IVP_BURGERS1D, _lib.TDX_IVP_WAVE1D):
        raise NotImplementedError(f"{spec.name} runs on one GPU only (its end-point rules couple the two ends of the mesh)")
    if spec.kind in (_lib.TDX_IVP_THREEBODY, _lib.TDX_IVP_PLEIADES):
        raise ValueError(f"{spec.name} (d = {spec.grid_cols}) cannot be sharded across GPUs")
    if extent < 2 * world:
        raise ValueError(f"cannot shard an axis of {extent} across {world} ranks")
    aligned = _chunk_aligned_bounds(spec, world)
    if aligned is not None:
        return aligned[rank], aligned[rank + 1]
    base, rem = divmod(extent, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduction_is_sharding_independent(spec, world, chunks=16):
    """True if the kernels' norm sums are bit-identical to a single-GPU run for this problem and number of ranks: every
    shard boundary has to fall on a boundary of the ``chunks`` equal parts of the reduction's global unit range (``RedPlan``
    in csrc/tdx_device.cuh).  Units: 256-coordinate tiles (fhn_2d: 4 grid rows x 32 columns of both fields; chains: 128
    points of both fields, or 256 points) for DiagonalEK1 and the 1-d KroneckerEK0 kernels, (grid row, 256-column block)
    pairs for the fused fhn_2d KroneckerEK0 kernel.  E.g. 2 / 4 / 8 / 16 equal shards of a 2048 x 2048 grid qualify."""
    if world == 1:
        return True

    def on_chunk_boundary(unit_index, units):
        return any((c * units) // chunks == unit_index for c in range(chunks + 1))

    bounds = [shard_range(spec, r, world)[0] for r in range(1, world)]
    if spec.kind == _lib.TDX_IVP_FHN2D:
        tile_rows, tile_cols, strip_cols = 4, 32, 256
        tiles_x = -(-spec.grid_cols // tile_cols)
        tiles_y = -(-spec.grid_rows // tile_rows)
        col_blocks = -(-spec.grid_cols // strip_cols)
        row_block = _ek0_row_block(spec.grid_rows, col_blocks)
        blocks_y = -(-spec.grid_rows // row_block)
        for b in bounds:
            if b % tile_rows or not on_chunk_boundary((b // tile_rows) * tiles_x, tiles_y * tiles_x):
                return False
            if b % row_block or not on_chunk_boundary((b // row_block) * col_blocks, blocks_y * col_blocks):
                return False
        return True
    tile = 256 if spec.n_fields == 1 else 128
    units = -(-spec.grid_cols // tile)
    return all(b % tile == 0 and on_chunk_boundary(b // tile, units) for b in bounds)


def local_slices(spec, begin, end):
    """Index ranges of the global flat state y = [field0; field1] owned by a shard (one per field)."""
    if spec.kind == _lib.TDX_IVP_FHN2D:
        npts = spec.grid_rows * spec.grid_cols
        lo, hi = begin * spec.grid_cols, end * spec.grid_cols
    else:
        npts = spec.grid_rows * spec.grid_cols
        lo, hi = begin, end
    return [slice(f * npts + lo, f * npts + hi) for f in range(spec.n_fields)]


class HaloExchanger:
    """Neighbour exchange of the shard-edge values (works on any torch backend: nccl on GPUs, gloo in tests).

    ``exchange(edge_first, edge_last)`` sends ``edge_first`` to the lower neighbour (it becomes its
    upper halo) and ``edge_last`` to the upper neighbour (its lower halo) and returns
    ``(halo_lo, halo_hi)``.  Non-cyclic problems have no neighbour beyond the global ends.
    """

    def __init__(self, group, rank, world, cyclic, lo_count, hi_count, dtype, device):
        import torch.distributed as dist

        self.dist = dist
        self.group, self.rank, self.world, self.cyclic = group, rank, world, cyclic
        self.lower = (rank - 1) % world if (cyclic or rank > 0) else None
        self.upper = (rank + 1) % world if (cyclic or rank < world - 1) else None
        self.halo_lo = torch.empty(lo_count, dtype=dtype, device=device) if self.lower is not None else None
        self.halo_hi = torch.empty(hi_count, dtype=dtype, device=device) if self.upper is not None else None

    def exchange(self, edge_first, edge_last):
        dist = self.dist
        ops = []
        if self.lower is not None:
            ops.append(dist.P2POp(dist.isend, edge_first, self._global(self.lower), self.group))
            ops.append(dist.P2POp(dist.irecv, self.halo_lo, self._global(self.lower), self.group))
        if self.upper is not None:
            ops.append(dist.P2POp(dist.isend, edge_last, self._global(self.upper), self.group))
            ops.append(dist.P2POp(dist.irecv, self.halo_hi, self._global(self.upper), self.group))
        if self.world == 2 and self.cyclic:
            # both neighbours are the same peer: order the four transfers consistently on both ranks
            ops = self._two_rank_cyclic(edge_first, edge_last)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return self.halo_lo, self.halo_hi

    def _global(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def _two_rank_cyclic(self, edge_first, edge_last):
        dist, peer = self.dist, self._global(1 - self.rank)
        if self.rank == 0:
            return [
                dist.P2POp(dist.isend, edge_first, peer, self.group),
                dist.P2POp(dist.isend, edge_last, peer, self.group),
                dist.P2POp(dist.irecv, self.halo_hi, peer, self.group),
                dist.P2POp(dist.irecv, self.halo_lo, peer, self.group),
            ]
        return [
            dist.P2POp(dist.irecv, self.halo_hi, peer, self.group),
            dist.P2POp(dist.irecv, self.halo_lo, peer, self.group),
            dist.P2POp(dist.isend, edge_first, peer, self.group),
            dist.P2POp(dist.isend, edge_last, peer, self.group),
        ]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class StepEngine:
    """One GPU's share of the filter step for a given (vector field, nu, dtype)."""

    def __init__(self, spec, nu, dtype=torch.float64, device=None, group=None, rank=0, world=1, peer_memory=True):
        if not torch.cuda.is_available():
            raise _lib.TornadoxB200Error(
                "tornadox_b200 needs a CUDA device (sm_100a); there is no CPU fallback for the filter step"
            )
        self.lib = _lib.load()
        self.spec, self.nu, self.n = spec, int(nu), int(nu) + 1
        self.dtype = dtype
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.group, self.rank, self.world = group, rank, world
        self.begin, self.end = shard_range(spec, rank, world)
        self.d_global = spec.n_fields * spec.grid_rows * spec.grid_cols

        handle = ctypes.c_void_p()
        tdx_dtype = {torch.float64: _lib.TDX_F64, torch.float32: _lib.TDX_F32}[dtype]
        _lib.check(self.lib.tdx_create(ctypes.byref(handle), self.device.index or 0, tdx_dtype), _lib.TornadoxB200Error)
        self.handle = handle
        desc = _lib.ProblemDesc()
        desc.kind, desc.nu = spec.kind, self.nu
        desc.grid_rows, desc.grid_cols = spec.grid_rows, spec.grid_cols
        desc.shard_begin, desc.shard_end = self.begin, self.end
        desc.n_params = len(spec.params)
        for i, v in enumerate(spec.params):
            desc.params[i] = float(v)
        _lib.check(self.lib.tdx_set_problem(self.handle, ctypes.byref(desc)))
        # use the LAPACK factor the reference itself computes (iwp.py:37) rather than the library's
        # extended-precision one: they differ by cond(Q) * eps, 1e-11 relative at nu = 5
        self.Ql = _lib.lapack_diffusion_sqrt(self.nu)
        _lib.check(self.lib.tdx_set_diffusion_sqrt(
            self.handle, self.Ql.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        self.d_local = int(self.lib.tdx_local_dim(self.handle))
        lo, hi = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.tdx_halo_sizes(self.handle, ctypes.byref(lo), ctypes.byref(hi)))
        self.halo_lo_n, self.halo_hi_n = lo.value, hi.value

        # device scalars, a ring of 16 rows so that the sums of recent attempts stay readable:
        # [0:2] norm sums (interior, rim launches), [2] EK0 z^2 sum
        self._scalar_ring = torch.zeros((16, 4), dtype=torch.float64, device=self.device)
        self._ring_pos = 0
        # interior/rim split with the NCCL exchange on a second stream: NCCL's own kernels cannot be scheduled while
        # the persistent step kernel holds every SM, so this does not pay off; kept for experiments only
        self.overlap = False
        # default multi-GPU path: halos and scalar sums travel through NVLink peer memory written by our own kernels
        self.peer = False
        self._comm_stream = torch.cuda.Stream(device=self.device) if world > 1 else None
        self.exchanger = None
        if world > 1:
            cyclic = spec.kind == _lib.TDX_IVP_LORENZ96
            # what we send: first edge has the size of the lower neighbour's hi halo, etc.
            self._edge_first = torch.empty(self._edge_sizes()[0], dtype=dtype, device=self.device)
            self._edge_last = torch.empty(self._edge_sizes()[1], dtype=dtype, device=self.device)
            self.exchanger = HaloExchanger(group, rank, world, cyclic, self.halo_lo_n, self.halo_hi_n, dtype, self.device)
            if peer_memory:
                self._connect_peers(cyclic)

    def _connect_peers(self, cyclic):
        """Map every rank's exchange block through CUDA IPC (tdx_comm_init / tdx_comm_connect)."""
        import torch.distributed as dist

        rank, world = self.rank, self.world
        lower = (rank - 1) % world if (cyclic or rank > 0) else -1
        upper = (rank + 1) % world if (cyclic or rank < world - 1) else -1
        handle = (ctypes.c_ubyte * _lib.TDX_COMM_HANDLE_BYTES)()
        _lib.check(self.lib.tdx_comm_init(self.handle, rank, world, lower, upper, ctypes.cast(handle, ctypes.c_void_p)),
                   _lib.TornadoxB200Error)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        for peer, raw in enumerate(handles):
            if peer == rank:
                continue
            buf = (ctypes.c_ubyte * _lib.TDX_COMM_HANDLE_BYTES).from_buffer_copy(raw)
            _lib.check(self.lib.tdx_comm_connect(self.handle, peer, ctypes.cast(buf, ctypes.c_void_p)),
                       _lib.TornadoxB200Error)
        dist.barrier(group=self.group)
        self.peer = True

    def barrier(self):
        """All ranks of a sharded solver meet here (after the initialisation, whose duration differs from rank to rank): the
        in-kernel waits for a neighbour's halo / share of a reduction give up after TDX_WAIT_TIMEOUT_MS."""
        if self.world > 1:
            torch.cuda.current_stream(self.device).synchronize()
            self.exchanger.dist.barrier(group=self.group)

    def check_peers(self):
        """Raise if a flag wait timed out (a neighbouring rank died or diverged)."""
        if not self.peer:
            return
        err = ctypes.c_int(0)
        _lib.check(self.lib.tdx_comm_error(self.handle, ctypes.byref(err), self._stream()))
        if err.value:
            raise _lib.TornadoxB200Error("peer-memory exchange timed out: a neighbouring rank did not deliver its halo")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.tdx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- helpers ------------------------------------------------------------------------------
    def _edge_sizes(self):
        k = self.spec.kind
        if k == _lib.TDX_IVP_FHN2D:
            return 2 * self.spec.grid_cols, 2 * self.spec.grid_cols
        if k == _lib.TDX_IVP_BRUSSELATOR:
            return 2, 2
        if k == _lib.TDX_IVP_LORENZ96:
            return 1, 2
        return 0, 0

    def _stream(self):
        # torch.cuda.current_stream() builds a Stream object (7 us); the raw handle is all the C ABI needs
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(self._dev_index))

    def _args(self, t, dt, abstol, reltol, flags=0):
        return _lib.StepArgs(float(t), float(dt), float(abstol), float(reltol), int(flags), 0)

    def _check(self, x, shape, name):
        if not isinstance(x, torch.Tensor) or x.device != self.device or x.dtype != self.dtype:
            raise ValueError(f"{name} must be a {self.dtype} tensor on {self.device}")
        if tuple(x.shape) != tuple(shape):
            raise ValueError(f"{name} has shape {tuple(x.shape)}, expected {tuple(shape)}")
        if not x.is_contiguous():
            raise ValueError(f"{name} must be C-contiguous")

    def _next_scalars(self):
        self._ring_pos = (self._ring_pos + 1) % self._scalar_ring.shape[0]
        return self._scalar_ring[self._ring_pos]

    def launch_count(self):
        return int(self.lib.tdx_launch_count(self.handle))

    def all_reduce_(self, scalar):
        if self.world > 1:
            self.exchanger.dist.all_reduce(scalar, group=self.group)
        return scalar

    def _halos_for_mean(self, dt, mean):
        """Predicted-state halos for a step from ``mean`` (multi-GPU only), exchanged on the current stream."""
        if self.world == 1:
            return None, None
        _lib.check(self.lib.tdx_pack_halo(self.handle, float(dt), _ptr(mean), _ptr(self._edge_first),
                                          _ptr(self._edge_last), self._stream()))
        return self.exchanger.exchange(self._edge_first, self._edge_last)

    def _halos_overlapped(self, dt, mean):
        """Pack on the compute stream, exchange on the communication stream; returns (halo_lo, halo_hi, event)."""
        main = torch.cuda.current_stream(self.device)
        _lib.check(self.lib.tdx_pack_halo(self.handle, float(dt), _ptr(mean), _ptr(self._edge_first),
                                          _ptr(self._edge_last), self._stream()))
        self._comm_stream.wait_stream(main)
        with torch.cuda.stream(self._comm_stream):
            halo_lo, halo_hi = self.exchanger.exchange(self._edge_first, self._edge_last)
            ready = torch.cuda.Event()
            ready.record(self._comm_stream)
        return halo_lo, halo_hi, ready

    def _all_reduce_on_comm_stream(self, scalars):
        """Sum over ranks without stalling the compute stream; returns the event to wait on before reading."""
        main = torch.cuda.current_stream(self.device)
        self._comm_stream.wait_stream(main)
        with torch.cuda.stream(self._comm_stream):
            self.exchanger.dist.all_reduce(scalars, group=self.group)
            done = torch.cuda.Event()
            done.record(self._comm_stream)
        return done

    # ---- vector field ---------------------------------------------------------------------------
    def eval_f(self, t, y, want_jac=False):
        """f(t, y) (and diag df) on the local shard -- replaces ivp.f / ivp.df_diagonal."""
        self._check(y, (self.d_local,), "y")
        halo_lo = halo_hi = None
        if self.world > 1:
            k = self.spec.kind
            npts = self.d_local // self.spec.n_fields
            yf = y.view(self.spec.n_fields, npts)
            nf_, nl_ = {_lib.TDX_IVP_FHN2D: (self.spec.grid_cols,) * 2, _lib.TDX_IVP_BRUSSELATOR: (1, 1),
                        _lib.TDX_IVP_LORENZ96: (1, 2)}[k]
            self._edge_first.copy_(yf[:, :nf_].reshape(-1))
            self._edge_last.copy_(yf[:, npts - nl_:].reshape(-1))
            halo_lo, halo_hi = self.exchanger.exchange(self._edge_first, self._edge_last)
        fy = torch.empty_like(y)
        jd = torch.empty_like(y) if want_jac else None
        _lib.check(self.lib.tdx_eval_f(self.handle, float(t), _ptr(y), _ptr(fy), _ptr(jd), _ptr(halo_lo),
                                       _ptr(halo_hi), self._stream()))
        return (fy, jd) if want_jac else fy

    # ---- DiagonalEK1 ------------------------------------------------------------------------------
    def dek1_attempt(self, t, dt, mean, sc, abstol=0.0, reltol=0.0, mean_out=None, sc_out=None, want_err_ref=True,
                     jacobian=True):
        """One DiagonalEK1 attempt (``jacobian=False``: DiagonalEK0).  Returns (mean_out, sc_out, err, ref, sq_norm_sum[device scalar])."""
        n, d = self.n, self.d_local
        self._check(mean, (n, d), "mean")
        self._check(sc, (d, n, n), "cov_sqrtm")
        mean_out = torch.empty_like(mean) if mean_out is None else mean_out
        sc_out = torch.empty_like(sc) if sc_out is None else sc_out
        err = torch.empty(d, dtype=self.dtype, device=self.device) if want_err_ref else None
        ref = torch.empty(d, dtype=self.dtype, device=self.device) if want_err_ref else None
        base_flags = 0 if jacobian else _lib.TDX_STEP_ZERO_JACOBIAN
        want_norm = abstol != 0.0 or reltol != 0.0
        call = self.lib.tdx_dek1_attempt_step
        bufs = (_ptr(mean), _ptr(sc), _ptr(mean_out), _ptr(sc_out), _ptr(err), _ptr(ref))
        scal = self._next_scalars()
        if self.world > 1 and self.peer:
            # ONE launch per attempt, no NCCL: a few CTAs first store this shard's edge state straight into the
            # neighbours' halo buffers (NVLink), all CTAs do the interior tiles first and await the arrival flags only
            # when they reach the rim tiles, and the last CTA all-reduces the norm sum through peer memory
            st = self._stream()
            a_all = self._args(t, dt, abstol, reltol, base_flags | _lib.TDX_STEP_COMM_PACK | _lib.TDX_STEP_COMM_REDUCE)
            sq = scal[0:1]
            _lib.check(call(self.handle, ctypes.byref(a_all), *bufs, None, None, _ptr(sq), st))
            return mean_out, sc_out, err, ref, sq
        if self.world > 1 and self.overlap:
            # interior tiles run while the halos travel; the rim tiles follow once they have arrived
            halo_lo, halo_hi, ready = self._halos_overlapped(dt, mean)
            a_int = self._args(t, dt, abstol, reltol, base_flags | _lib.TDX_STEP_TILES_INTERIOR)
            _lib.check(call(self.handle, ctypes.byref(a_int), *bufs, None, None, _ptr(scal[0:1]), self._stream()))
            torch.cuda.current_stream(self.device).wait_event(ready)
            a_rim = self._args(t, dt, abstol, reltol, base_flags | _lib.TDX_STEP_TILES_RIM)
            _lib.check(call(self.handle, ctypes.byref(a_rim), *bufs, _ptr(halo_lo), _ptr(halo_hi),
                            _ptr(scal[1:2]), self._stream()))
            sq = scal[0:2]
            if want_norm:
                sq.ready_event = self._all_reduce_on_comm_stream(sq)
            return mean_out, sc_out, err, ref, sq
        halo_lo, halo_hi = self._halos_for_mean(dt, mean)
        args = self._args(t, dt, abstol, reltol, base_flags)
        sq = scal[0:1]
        _lib.check(call(self.handle, ctypes.byref(args), *bufs, _ptr(halo_lo), _ptr(halo_hi), _ptr(sq), self._stream()))
        if want_norm:
            self.all_reduce_(sq)
        return mean_out, sc_out, err, ref, sq

    def dek1_attempt_host(self, t, dt, mean, sc, abstol=0.0, reltol=0.0, mean_out=None, sc_out=None, want_err_ref=True,
                          jacobian=True, nchunks=0):
        """One DiagonalEK1 attempt on a state in HOST memory (CPU tensors, pinned for full PCIe speed): the library streams
        the state through the GPU in chunks, copies in both directions overlapping the kernels (``tdx_dek1_attempt_step_host``).
        Returns (mean_out, sc_out, err, ref, sq_norm_sum[float]) as CPU tensors; synchronous."""
        n, d = self.n, self.d_local
        for x, shape, name in ((mean, (n, d), "mean"), (sc, (d, n, n), "cov_sqrtm")):
            if not isinstance(x, torch.Tensor) or x.device.type != "cpu" or x.dtype != self.dtype:
                raise ValueError(f"{name} must be a {self.dtype} CPU tensor")
            if tuple(x.shape) != tuple(shape) or not x.is_contiguous():
                raise ValueError(f"{name} must be C-contiguous with shape {tuple(shape)}")

        def host_empty(shape):
            return torch.empty(shape, dtype=self.dtype, pin_memory=True)

        mean_out = host_empty((n, d)) if mean_out is None else mean_out
        sc_out = host_empty((d, n, n)) if sc_out is None else sc_out
        err = host_empty((d,)) if want_err_ref else None
        ref = host_empty((d,)) if want_err_ref else None
        sq = ctypes.c_double(0.0)
        args = self._args(t, dt, abstol, reltol, 0 if jacobian else _lib.TDX_STEP_ZERO_JACOBIAN)
        if self.world > 1:
            # one shard of a multi-GPU run: the stencil halos of the shard's edges have to come from the neighbouring ranks
            # first -- upload just the edge rows / points, pack, exchange -- then the shard streams through its own pipeline
            halo_lo, halo_hi = self._exchange_halos_of_host_mean(dt, mean)
            torch.cuda.current_stream(self.device).synchronize()      # the library's pipeline runs on its own streams
            _lib.check(self.lib.tdx_dek1_attempt_step_host_sharded(
                self.handle, ctypes.byref(args), _ptr(mean), _ptr(sc), _ptr(mean_out), _ptr(sc_out), _ptr(err), _ptr(ref),
                _ptr(halo_lo), _ptr(halo_hi), ctypes.byref(sq), int(nchunks)))
            total = torch.tensor([sq.value], dtype=torch.float64, device=self.device)
            self.all_reduce_(total)
            return mean_out, sc_out, err, ref, float(total.item())
        _lib.check(self.lib.tdx_dek1_attempt_step_host(self.handle, ctypes.byref(args), _ptr(mean), _ptr(sc), _ptr(mean_out),
                                                       _ptr(sc_out), _ptr(err), _ptr(ref), ctypes.byref(sq), int(nchunks)))
        return mean_out, sc_out, err, ref, sq.value

    def _exchange_halos_of_host_mean(self, dt, host_mean):
        """Predicted edge state of a HOST-resident shard mean, exchanged with the neighbouring ranks: only the shard's first
        and last grid row (fhn_2d) / two points (chains) of every derivative row travel to the device for this."""
        n, d = self.n, self.d_local
        if getattr(self, "_edge_mean_dev", None) is None:
            self._edge_mean_dev = torch.zeros((n, d), dtype=self.dtype, device=self.device)
        npts = d // self.spec.n_fields
        width = self.spec.grid_cols if self.spec.kind == _lib.TDX_IVP_FHN2D else min(2, npts)
        for f in range(self.spec.n_fields):
            for a in (f * npts, f * npts + npts - width):
                self._edge_mean_dev[:, a:a + width].copy_(host_mean[:, a:a + width], non_blocking=True)
        return self._halos_for_mean(dt, self._edge_mean_dev)

    # ---- native accept / reject loop (odefilter.py:171-223 inside the library) --------------------------------------
    def native_steprule(self, steprule):
        """The C mirror of a step rule, or None if it is not exactly one of the reference's two rules."""
        from . import step as _step

        if type(steprule) is _step.AdaptiveSteps:
            lower, upper = steprule.max_changes
            return _lib.StepRule(1, 0, float(steprule.abstol), float(steprule.reltol), float(lower), float(upper),
                                 float(steprule.safety_scale), 0.0)
        if type(steprule) is _step.ConstantSteps:
            return _lib.StepRule(0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, float(steprule.dt))
        return None

    @property
    def external(self):
        """The vector field is the caller's own code (``TDX_IVP_EXTERNAL``): every attempt needs f at the predicted state."""
        return self.spec.kind == _lib.TDX_IVP_EXTERNAL

    def can_run_full_steps(self):
        # the library's own accept / reject loops cannot call back into the caller's f
        return (self.world == 1 or self.peer) and not self.external

    def predict_state(self, dt, mean):
        """E0 P A P^-1 m: the point where the attempt with step ``dt`` from ``mean`` evaluates f (ek1.py:304-306, ek0.py:146-147)."""
        self._check(mean, (self.n, self.d_local), "mean")
        out = torch.empty(self.d_local, dtype=self.dtype, device=self.device)
        _lib.check(self.lib.tdx_predict_state(self.handle, float(dt), _ptr(mean), _ptr(out), self._stream()))
        return out

    def set_external(self, fx, jac=None):
        """Hand the caller's f(t + dt, m_at) (and diag df) to the next attempt of an EXTERNAL vector field."""
        self._check(fx, (self.d_local,), "f(t, y)")
        if jac is not None:
            self._check(jac, (self.d_local,), "df_diagonal(t, y)")
        self._external_arrays = (fx, jac)            # keep them alive until the step has run
        _lib.check(self.lib.tdx_set_external(self.handle, _ptr(fx), _ptr(jac)))

    def evaluate_external(self, f, df_diagonal, t, dt, mean, jacobian):
        """predict -> the caller's f / df_diagonal on device tensors -> hand the arrays in."""
        m_at = self.predict_state(dt, mean)
        fx = torch.as_tensor(f(t + dt, m_at), dtype=self.dtype, device=self.device).reshape(-1).contiguous()
        jac = None
        if jacobian:
            if df_diagonal is None:
                raise ValueError("DiagonalEK1 needs df_diagonal(t, y) for a user-defined vector field")
            jac = torch.as_tensor(df_diagonal(t + dt, m_at), dtype=self.dtype, device=self.device).reshape(-1).contiguous()
        self.set_external(fx, jac)

    def _full_step_flags(self, base=0):
        return base | ((_lib.TDX_STEP_COMM_PACK | _lib.TDX_STEP_COMM_REDUCE) if self.world > 1 else 0)

    def dek1_full_step(self, rule, t, dt, tmax, mean, sc, want_err_ref=True, jacobian=True):
        """Attempt DiagonalEK1 steps from (mean, sc) until one is accepted; returns (mean_out, sc_out, err, ref, result)."""
        n, d = self.n, self.d_local
        self._check(mean, (n, d), "mean")
        self._check(sc, (d, n, n), "cov_sqrtm")
        mean_out, sc_out = torch.empty_like(mean), torch.empty_like(sc)
        err = torch.empty(d, dtype=self.dtype, device=self.device) if want_err_ref else None
        ref = torch.empty(d, dtype=self.dtype, device=self.device) if want_err_ref else None
        res = _lib.FullStepResult()
        flags = self._full_step_flags(0 if jacobian else _lib.TDX_STEP_ZERO_JACOBIAN)
        _lib.check(self.lib.tdx_dek1_full_step(self.handle, ctypes.byref(rule), float(t), float(dt), float(tmax), flags,
                                               _ptr(mean), _ptr(sc), _ptr(mean_out), _ptr(sc_out), _ptr(err), _ptr(ref),
                                               ctypes.byref(res), self._stream()))
        return mean_out, sc_out, err, ref, res

    def ek0_full_step(self, rule, t, dt, tmax, mean, Cl, want_ref=True):
        """Attempt KroneckerEK0 steps from (mean, Cl) until one is accepted; returns (mean_out, Cl_out, err, ref, result)."""
        n, d = self.n, self.d_local
        self._check(mean, (n, d), "mean")
        self._check(Cl, (n, n), "cov_sqrtm_2")
        mean_out, Cl_out = torch.empty_like(mean), torch.empty_like(Cl)
        err = torch.empty((), dtype=self.dtype, device=self.device)
        ref = torch.empty(d, dtype=self.dtype, device=self.device) if want_ref else None
        res = _lib.FullStepResult()
        _lib.check(self.lib.tdx_ek0_full_step(self.handle, ctypes.byref(rule), float(t), float(dt), float(tmax),
                                              self._full_step_flags(), _ptr(mean), _ptr(Cl), _ptr(mean_out), _ptr(Cl_out),
                                              _ptr(err), _ptr(ref), ctypes.byref(res), self._stream()))
        return mean_out, Cl_out, err, ref, res

    # ---- device-resident step controller (no host involvement between attempts) -------------------------------------
    def ctl_begin(self, rule, t0, dt0, tmax, nbuf=2, stop_at=None, follow=False):
        """Reset the device-resident controller.  ``nbuf``: size of the state ring; ``stop_at``: time stops
        (odefilter.py:355-373); ``follow``: the host follows the running solve (``ctl_poll`` / ``ctl_history`` / ``ctl_abort``)."""
        opt = _lib.CtlOptions()
        opt.nbuf, opt.follow = int(nbuf), 1 if follow else 0
        stops = None
        if stop_at is not None:
            stops = np.ascontiguousarray(np.asarray(list(stop_at), dtype=np.float64))
            opt.stop_at = stops.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            opt.n_stops = int(stops.shape[0])
        _lib.check(self.lib.tdx_ctl_begin_ex(self.handle, ctypes.byref(rule), float(t0), float(dt0), float(tmax),
                                             ctypes.byref(opt), self._stream()))

    def ctl_read(self, history=False):
        st = _lib.CtlState()
        if history:
            ht = np.zeros(_lib.TDX_CTL_HISTORY)
            hd = np.zeros(_lib.TDX_CTL_HISTORY)
            pd = ctypes.POINTER(ctypes.c_double)
            _lib.check(self.lib.tdx_ctl_read(self.handle, ctypes.byref(st), ht.ctypes.data_as(pd), hd.ctypes.data_as(pd), self._stream()))
            return st, ht, hd
        _lib.check(self.lib.tdx_ctl_read(self.handle, ctypes.byref(st), None, None, self._stream()))
        return st

    def ctl_poll(self):
        """(accepted, attempts) of the running solve from the mapped pinned feed; no synchronisation."""
        pr = _lib.CtlProgress()
        _lib.check(self.lib.tdx_ctl_poll(self.handle, ctypes.byref(pr)))
        return int(pr.accepted), int(pr.attempts)

    def ctl_history(self, state_index):
        """(t, dt, attempts so far) of accepted state ``state_index`` >= 1 from the feed."""
        t, dt, att = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        _lib.check(self.lib.tdx_ctl_history(self.handle, int(state_index), ctypes.byref(t), ctypes.byref(dt), ctypes.byref(att)))
        return t.value, dt.value, int(att.value)

    def ctl_abort(self):
        _lib.check(self.lib.tdx_ctl_abort(self.handle))

    def raise_if_failed(self, ctl):
        """Map the controller's failure codes to the exceptions of the host-driven loops."""
        if ctl.failed == 1:
            raise FloatingPointError(f"error estimate is NaN at t={ctl.t}, dt={ctl.dt}")
        if ctl.failed:
            why = {2: "an attempt was poisoned (a peer's halo or share of a reduction did not arrive in time)",
                   3: "the grid barrier between two attempts was abandoned",
                   4: "host feed",
                   5: "a wait for a neighbouring rank timed out"}.get(int(ctl.failed), f"code {ctl.failed}")
            raise _lib.TornadoxB200Error(f"a wait inside the step kernels timed out: {why}; the state is not valid "
                                         f"(t={ctl.t}, attempts={ctl.attempts})")

    @staticmethod
    def _ring(tensors):
        if tensors is None:
            return None
        return (ctypes.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else None) for t in tensors])

    def dek1_enqueue_attempts(self, k, pair_a, pair_b, err=None, ref=None, jacobian=True):
        flags = self._full_step_flags(0 if jacobian else _lib.TDX_STEP_ZERO_JACOBIAN)
        _lib.check(self.lib.tdx_dek1_enqueue_attempts(self.handle, int(k), flags, _ptr(pair_a[0]), _ptr(pair_a[1]), _ptr(pair_b[0]),
                                                      _ptr(pair_b[1]), _ptr(err), _ptr(ref), self._stream()))

    def ek0_enqueue_attempts(self, k, pair_a, pair_b, err, ref=None):
        _lib.check(self.lib.tdx_ek0_enqueue_attempts(self.handle, int(k), self._full_step_flags(), _ptr(pair_a[0]), _ptr(pair_a[1]),
                                                     _ptr(pair_b[0]), _ptr(pair_b[1]), _ptr(err), _ptr(ref), self._stream()))

    def dek1_run_attempts(self, k, means, factors, errs=None, refs=None, jacobian=True, stream=None, max_accepted=0):
        """Up to ``k`` DiagonalEK1 attempts in ONE persistent launch over the state ring ``means`` / ``factors`` (lists of
        tensors; ``errs`` / ``refs``: per-slot error_estimate / reference_state vectors or None).  ``max_accepted`` > 0 ends
        the launch after that many newly accepted steps."""
        flags = self._full_step_flags(0 if jacobian else _lib.TDX_STEP_ZERO_JACOBIAN)
        _lib.check(self.lib.tdx_dek1_run_attempts(self.handle, int(k), int(max_accepted), flags, self._ring(means), self._ring(factors),
                                                  len(means), self._ring(errs), self._ring(refs),
                                                  stream if stream is not None else self._stream()))

    def ek0_run_attempts(self, k, means, factors, errs, refs=None, stream=None, max_accepted=0):
        """Up to ``k`` KroneckerEK0 attempts in ONE persistent launch (``errs``: one device scalar per ring slot)."""
        _lib.check(self.lib.tdx_ek0_run_attempts(self.handle, int(k), int(max_accepted), self._full_step_flags(), self._ring(means),
                                                 self._ring(factors), len(means), self._ring(errs), self._ring(refs),
                                                 stream if stream is not None else self._stream()))

    # ---- KroneckerEK0 -----------------------------------------------------------------------------
    def ek0_attempt(self, t, dt, mean, Cl, abstol=0.0, reltol=0.0, mean_out=None, want_ref=True):
        """One KroneckerEK0 attempt.  Returns (mean_out, Cl_out, err[device scalar], ref, sq_norm_sum)."""
        n, d = self.n, self.d_local
        self._check(mean, (n, d), "mean")
        self._check(Cl, (n, n), "cov_sqrtm_2")
        mean_out = torch.empty_like(mean) if mean_out is None else mean_out
        Cl_out = torch.empty_like(Cl)
        err = torch.empty((), dtype=self.dtype, device=self.device)
        ref = torch.empty(d, dtype=self.dtype, device=self.device) if want_ref else None
        scal = self._next_scalars()
        zsq, itol = scal[2:4], scal[0:1]
        st = self._stream()
        want_norm = abstol != 0.0 or reltol != 0.0
        args = self._args(t, dt, abstol, reltol)
        if self.world == 1 or self.peer:
            # ONE call per attempt (fhn_2d: one cooperative launch that holds both sweeps, the calibration and, on
            # several GPUs, the halo delivery and both scalar all-reduces over NVLink peer memory)
            flags = (_lib.TDX_STEP_COMM_PACK | _lib.TDX_STEP_COMM_REDUCE) if self.world > 1 else 0
            one = self._args(t, dt, abstol, reltol, flags)
            _lib.check(self.lib.tdx_ek0_attempt_step(self.handle, ctypes.byref(one), _ptr(mean), _ptr(Cl), _ptr(mean_out),
                                                     _ptr(Cl_out), _ptr(err), _ptr(ref), _ptr(zsq[0:1]), _ptr(itol), st))
            return mean_out, Cl_out, err, ref, itol
        if self.world > 1 and self.overlap:
            halo_lo, halo_hi, ready = self._halos_overlapped(dt, mean)
            a_int = self._args(t, dt, abstol, reltol, _lib.TDX_STEP_TILES_INTERIOR)
            _lib.check(self.lib.tdx_ek0_sweep_a(self.handle, ctypes.byref(a_int), _ptr(mean), None, None,
                                                _ptr(zsq[0:1]), st))
            torch.cuda.current_stream(self.device).wait_event(ready)
            a_rim = self._args(t, dt, abstol, reltol, _lib.TDX_STEP_TILES_RIM)
            _lib.check(self.lib.tdx_ek0_sweep_a(self.handle, ctypes.byref(a_rim), _ptr(mean), _ptr(halo_lo),
                                                _ptr(halo_hi), _ptr(zsq[1:2]), st))
            self.all_reduce_(zsq)                       # the calibration needs the global sum before anything else
            a_mid = self._args(t, dt, abstol, reltol, _lib.TDX_STEP_TWO_PARTS)
        else:
            halo_lo, halo_hi = self._halos_for_mean(dt, mean)
            _lib.check(self.lib.tdx_ek0_sweep_a(self.handle, ctypes.byref(args), _ptr(mean), _ptr(halo_lo),
                                                _ptr(halo_hi), _ptr(zsq[0:1]), st))
            self.all_reduce_(zsq[0:1])
            a_mid = args
        _lib.check(self.lib.tdx_ek0_mid(self.handle, ctypes.byref(a_mid), _ptr(Cl), _ptr(zsq), self.d_global,
                                        _ptr(Cl_out), _ptr(err), st))
        _lib.check(self.lib.tdx_ek0_sweep_b(self.handle, ctypes.byref(args), _ptr(mean), _ptr(mean_out), _ptr(ref),
                                            _ptr(halo_lo), _ptr(halo_hi), _ptr(itol), st))
        if want_norm and self.world > 1:
            if self.overlap:
                itol.ready_event = self._all_reduce_on_comm_stream(itol)
            else:
                self.all_reduce_(itol)
        return mean_out, Cl_out, err, ref, itol


def to_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
