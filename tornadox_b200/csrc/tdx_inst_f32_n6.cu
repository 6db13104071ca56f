// explicit instantiation of the step kernels for dtype=float, n = nu + 1 = 6
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(float, 6, f32_6)
