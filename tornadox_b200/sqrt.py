"""Square-root transition utilities on the GPU.  Mirrors tornadox/sqrt.py: ``propagate_cholesky_factor`` (:8-12),
``sqrtm_to_cholesky`` (:15-23) and their vmapped versions (:27-30).

Inside the solvers these operations are fused into the step kernels (``stacked_qr`` in csrc/tdx_device.cuh); the functions
here expose the same factorisation as a standalone batched CUDA op (``tdx_propagate_cholesky_factor``: one thread per
(2n x n) stacked matrix, Householder QR in registers with LAPACK's sign convention, i.e. what ``jnp.linalg.qr`` returns on
CPU) for code that calls the reference's helpers directly.  Device tensors in, device tensors out; no CPU fallback.
"""

import ctypes

import torch

from . import _lib


def _launch(s1, s2, batch, n, s1_broadcast, s2_broadcast, rows_given):
    for x in (s1, s2):
        if x is not None and (not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype not in (torch.float64, torch.float32)):
            raise ValueError("square-root factors must be float64 / float32 CUDA tensors")
    if s2 is not None and (s2.dtype != s1.dtype or s2.device != s1.device):
        raise ValueError("S1 and S2 must share dtype and device")
    lib = _lib.load()
    out = torch.empty((batch, n, n), dtype=s1.dtype, device=s1.device)
    dtype = _lib.TDX_F64 if s1.dtype == torch.float64 else _lib.TDX_F32
    with torch.cuda.device(s1.device):
        stream = ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(s1.device.index))
        _lib.check(lib.tdx_propagate_cholesky_factor(
            dtype, n, batch, ctypes.c_void_p(s1.data_ptr()), ctypes.c_void_p(s2.data_ptr()) if s2 is not None else None,
            ctypes.c_void_p(out.data_ptr()), int(s1_broadcast), int(s2_broadcast), int(rows_given), stream))
    return out


def _square(x, name):
    if x.ndim < 2 or x.shape[-1] != x.shape[-2]:
        raise ValueError(f"{name} must be (..., n, n), got {tuple(x.shape)}")
    return x.shape[-1]


def propagate_cholesky_factor(S1, S2):
    """Cholesky factor of S1 S1^T + S2 S2^T for two (n, n) square roots (sqrt.py:8-12)."""
    n = _square(S1, "S1")
    if tuple(S2.shape) != (n, n) or S1.ndim != 2:
        raise ValueError("propagate_cholesky_factor takes two (n, n) matrices; use batched_propagate_cholesky_factor")
    return _launch(S1.contiguous(), S2.contiguous(), 1, n, 0, 0, 0)[0]


def batched_propagate_cholesky_factor(S1, S2):
    """vmap of ``propagate_cholesky_factor`` over the leading axis (sqrt.py:27-29).  Either argument may also be a single
    (n, n) matrix shared by the whole batch (what DiagonalEK1 does with sigma * Ql)."""
    n = _square(S1, "S1")
    if _square(S2, "S2") != n:
        raise ValueError("S1 and S2 must have the same block size")
    b1 = S1.shape[0] if S1.ndim == 3 else None
    b2 = S2.shape[0] if S2.ndim == 3 else None
    if b1 is not None and b2 is not None and b1 != b2:
        raise ValueError("batch sizes differ")
    batch = b1 if b1 is not None else (b2 if b2 is not None else 1)
    return _launch(S1.contiguous(), S2.contiguous(), batch, n, b1 is None, b2 is None, 0)


def sqrtm_to_cholesky(St):
    """Lower factor L = R^T of a 'right' square root St with M = St^T St ... i.e. St = S^T, M = S S^T (sqrt.py:15-23).
    St is (n, n) or (2n, n)."""
    return batched_sqrtm_to_cholesky(St[None])[0]


def batched_sqrtm_to_cholesky(St):
    """vmap of ``sqrtm_to_cholesky`` over the leading axis (sqrt.py:30); St is (b, n, n) or (b, 2n, n)."""
    if St.ndim != 3 or St.shape[1] not in (St.shape[2], 2 * St.shape[2]):
        raise ValueError(f"St must be (b, n, n) or (b, 2n, n), got {tuple(St.shape)}")
    b, m, n = St.shape
    St = St.contiguous()
    if m == n:
        return _launch(St, None, b, n, 0, 0, 1)
    top, bottom = St[:, :n, :].contiguous(), St[:, n:, :].contiguous()
    return _launch(top, bottom, b, n, 0, 0, 1)
