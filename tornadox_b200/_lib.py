"""ctypes binding of ``libtornadox_b200.so`` (C ABI: include/tornadox_b200.h).

There is NO fallback: if the shared library is missing or no CUDA device is present, every entry
point that needs compute raises.  torch is used only for device buffers and streams.
"""

import ctypes
import os
import pathlib

import numpy as np

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("TDX_LIB", _PKG / "lib" / "libtornadox_b200.so"))   # TDX_LIB: tuning variants

TDX_F64, TDX_F32 = 0, 1
TDX_STEP_ZERO_JACOBIAN = 1
TDX_STEP_TILES_INTERIOR = 2
TDX_STEP_TILES_RIM = 4
TDX_STEP_TWO_PARTS = 8
TDX_STEP_COMM_HALO = 16
TDX_STEP_COMM_REDUCE = 32
TDX_STEP_COMM_PACK = 64
TDX_COMM_HANDLE_BYTES = 64
TDX_IVP_VANDERPOL, TDX_IVP_LORENZ96, TDX_IVP_BRUSSELATOR, TDX_IVP_FHN2D, TDX_IVP_BURGERS1D, TDX_IVP_WAVE1D = 0, 1, 2, 3, 4, 5
TDX_IVP_THREEBODY, TDX_IVP_PLEIADES, TDX_IVP_EXTERNAL = 6, 7, 8

EXPORTED_SYMBOLS = (
    "tdx_last_error", "tdx_abi_version", "tdx_create", "tdx_destroy", "tdx_set_problem", "tdx_local_dim",
    "tdx_halo_sizes", "tdx_set_diffusion_sqrt", "tdx_prior", "tdx_nordsieck", "tdx_eval_f", "tdx_pack_halo", "tdx_dek1_attempt_step",
    "tdx_ek0_sweep_a", "tdx_ek0_mid", "tdx_ek0_sweep_b", "tdx_ek0_attempt_step", "tdx_launch_count",
    "tdx_debug_read_workspace", "tdx_dek1_attempt_step_host", "tdx_dek1_attempt_step_host_sharded", "tdx_dek1_full_step", "tdx_ek0_full_step", "tdx_propagate_cholesky_factor",
    "tdx_ctl_begin", "tdx_dek1_enqueue_attempts", "tdx_ek0_enqueue_attempts", "tdx_ctl_read",
    "tdx_ctl_begin_ex", "tdx_dek1_run_attempts", "tdx_ek0_run_attempts", "tdx_ctl_poll", "tdx_ctl_history", "tdx_ctl_abort",
    "tdx_predict_state", "tdx_set_external",
    "tdx_comm_init", "tdx_comm_connect", "tdx_comm_pack_halo", "tdx_comm_allreduce", "tdx_comm_error",
)


class ProblemDesc(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("nu", ctypes.c_int32),
        ("grid_rows", ctypes.c_int64),
        ("grid_cols", ctypes.c_int64),
        ("shard_begin", ctypes.c_int64),
        ("shard_end", ctypes.c_int64),
        ("n_params", ctypes.c_int32),
        ("params", ctypes.c_double * 8),
    ]


class StepArgs(ctypes.Structure):
    _fields_ = [
        ("t", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("abstol", ctypes.c_double),
        ("reltol", ctypes.c_double),
        ("flags", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
    ]


class StepRule(ctypes.Structure):
    _fields_ = [
        ("adaptive", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("abstol", ctypes.c_double),
        ("reltol", ctypes.c_double),
        ("min_change", ctypes.c_double),
        ("max_change", ctypes.c_double),
        ("safety_scale", ctypes.c_double),
        ("constant_dt", ctypes.c_double),
    ]


class FullStepResult(ctypes.Structure):
    _fields_ = [
        ("t_new", ctypes.c_double),
        ("dt_accepted", ctypes.c_double),
        ("dt_next", ctypes.c_double),
        ("error_norm", ctypes.c_double),
        ("attempts", ctypes.c_int64),
    ]


TDX_CTL_HISTORY = 256


class CtlState(ctypes.Structure):
    _fields_ = [
        ("t", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("last_norm", ctypes.c_double),
        ("attempts", ctypes.c_int64),
        ("accepted", ctypes.c_int64),
        ("cur", ctypes.c_int32),
        ("done", ctypes.c_int32),
        ("failed", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


TDX_CTL_MAX_RING = 8


class CtlOptions(ctypes.Structure):
    _fields_ = [
        ("nbuf", ctypes.c_int32),
        ("follow", ctypes.c_int32),
        ("stop_at", ctypes.POINTER(ctypes.c_double)),
        ("n_stops", ctypes.c_int64),
    ]


class CtlProgress(ctypes.Structure):
    _fields_ = [("accepted", ctypes.c_int64), ("attempts", ctypes.c_int64)]


class TornadoxB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (building is the job of ``tornadox_b200.build`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TornadoxB200Error(
            f"{LIB_PATH} not found: build it with `python -m tornadox_b200.build` "
            "(tornadox_b200 has no CPU fallback and cannot run without its CUDA library)"
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i64, dbl, cint = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
    pd = ctypes.POINTER(ctypes.c_double)
    lib.tdx_last_error.restype = ctypes.c_char_p
    lib.tdx_last_error.argtypes = []
    lib.tdx_abi_version.restype = cint
    lib.tdx_create.argtypes = [ctypes.POINTER(vp), cint, cint]
    lib.tdx_destroy.argtypes = [vp]
    lib.tdx_set_problem.argtypes = [vp, ctypes.POINTER(ProblemDesc)]
    lib.tdx_local_dim.restype = i64
    lib.tdx_local_dim.argtypes = [vp]
    lib.tdx_halo_sizes.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.tdx_prior.argtypes = [cint, pd, pd]
    lib.tdx_set_diffusion_sqrt.argtypes = [vp, pd]
    lib.tdx_nordsieck.argtypes = [cint, dbl, pd, pd]
    lib.tdx_eval_f.argtypes = [vp, dbl, vp, vp, vp, vp, vp, vp]
    lib.tdx_pack_halo.argtypes = [vp, dbl, vp, vp, vp, vp]
    lib.tdx_predict_state.argtypes = [vp, dbl, vp, vp, vp]
    lib.tdx_set_external.argtypes = [vp, vp, vp]
    lib.tdx_dek1_attempt_step.argtypes = [vp, ctypes.POINTER(StepArgs)] + [vp] * 10
    lib.tdx_dek1_attempt_step_host.argtypes = [vp, ctypes.POINTER(StepArgs)] + [vp] * 6 + [pd, cint]
    full = [vp, ctypes.POINTER(StepRule), dbl, dbl, dbl, ctypes.c_uint32] + [vp] * 6 + [ctypes.POINTER(FullStepResult), vp]
    lib.tdx_dek1_full_step.argtypes = full
    lib.tdx_ek0_full_step.argtypes = full
    lib.tdx_propagate_cholesky_factor.argtypes = [cint, cint, i64, vp, vp, vp, cint, cint, cint, vp]
    lib.tdx_ctl_begin.argtypes = [vp, ctypes.POINTER(StepRule), dbl, dbl, dbl, vp]
    lib.tdx_dek1_enqueue_attempts.argtypes = [vp, cint, ctypes.c_uint32] + [vp] * 7
    lib.tdx_ek0_enqueue_attempts.argtypes = [vp, cint, ctypes.c_uint32] + [vp] * 7
    lib.tdx_ctl_read.argtypes = [vp, ctypes.POINTER(CtlState), pd, pd, vp]
    lib.tdx_ctl_begin_ex.argtypes = [vp, ctypes.POINTER(StepRule), dbl, dbl, dbl, ctypes.POINTER(CtlOptions), vp]
    ring = ctypes.POINTER(vp)
    lib.tdx_dek1_run_attempts.argtypes = [vp, cint, i64, ctypes.c_uint32, ring, ring, cint, ring, ring, vp]
    lib.tdx_ek0_run_attempts.argtypes = [vp, cint, i64, ctypes.c_uint32, ring, ring, cint, ring, ring, vp]
    lib.tdx_ctl_poll.argtypes = [vp, ctypes.POINTER(CtlProgress)]
    lib.tdx_ctl_history.argtypes = [vp, i64, pd, pd, ctypes.POINTER(i64)]
    lib.tdx_ctl_abort.argtypes = [vp]
    lib.tdx_dek1_attempt_step_host_sharded.argtypes = [vp, ctypes.POINTER(StepArgs)] + [vp] * 8 + [pd, cint]
    lib.tdx_ek0_sweep_a.argtypes = [vp, ctypes.POINTER(StepArgs), vp, vp, vp, vp, vp]
    lib.tdx_ek0_mid.argtypes = [vp, ctypes.POINTER(StepArgs), vp, vp, i64, vp, vp, vp]
    lib.tdx_ek0_sweep_b.argtypes = [vp, ctypes.POINTER(StepArgs), vp, vp, vp, vp, vp, vp, vp]
    lib.tdx_ek0_attempt_step.argtypes = [vp, ctypes.POINTER(StepArgs)] + [vp] * 9
    lib.tdx_comm_init.argtypes = [vp, cint, cint, cint, cint, vp]
    lib.tdx_comm_connect.argtypes = [vp, cint, vp]
    lib.tdx_comm_pack_halo.argtypes = [vp, dbl, vp, vp]
    lib.tdx_comm_allreduce.argtypes = [vp, vp, cint, vp]
    lib.tdx_comm_error.argtypes = [vp, ctypes.POINTER(cint), vp]
    lib.tdx_debug_read_workspace.argtypes = [vp, pd, i64, i64]
    lib.tdx_launch_count.restype = i64
    lib.tdx_launch_count.argtypes = [vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ("tdx_abi_version",):
            fn.restype = cint
    _lib = lib
    return lib


def check(rc, exc=None):
    """Map a non-zero status to a Python exception (TDX_ERR_INVALID -> ValueError, like the reference)."""
    if rc == 0:
        return
    msg = load().tdx_last_error().decode()
    if rc == 1:
        raise (exc or ValueError)(msg)
    if rc == 2:
        raise NotImplementedError(msg)
    if rc == 4:
        raise FloatingPointError(msg)
    raise TornadoxB200Error(msg)


def prior(nu):
    """(A, Ql) of the preconditioned IWP(nu) prior as the kernels use them (reference: iwp.py:19-37)."""
    n = nu + 1
    A = np.zeros((n, n))
    Ql = np.zeros((n, n))
    pd = ctypes.POINTER(ctypes.c_double)
    check(load().tdx_prior(nu, A.ctypes.data_as(pd), Ql.ctypes.data_as(pd)))
    return A, Ql


def lapack_diffusion_sqrt(nu):
    """Ql exactly as the reference computes it: LAPACK dpotrf of flip(Hilbert(n)) (iwp.py:36-37)."""
    n = nu + 1
    idx = np.arange(n)
    Q = 1.0 / (2 * nu + 1 - idx[:, None] - idx[None, :])
    return np.ascontiguousarray(np.linalg.cholesky(Q))


def nordsieck(nu, dt):
    """(p, p_inv) Nordsieck preconditioner vectors (reference: iwp.py:65-73)."""
    n = nu + 1
    p = np.zeros(n)
    pinv = np.zeros(n)
    pd = ctypes.POINTER(ctypes.c_double)
    check(load().tdx_nordsieck(nu, float(dt), p.ctypes.data_as(pd), pinv.ctypes.data_as(pd)))
    return p, pinv


def library_is_loaded():
    """True if libtornadox_b200.so is mapped into this process (used by tests / smoke)."""
    if not os.path.exists("/proc/self/maps"):
        return _lib is not None
    with open("/proc/self/maps") as fh:
        return "libtornadox_b200.so" in fh.read()
