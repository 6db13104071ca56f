// explicit instantiation of the step kernels for dtype=float, n = nu + 1 = 3
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(float, 3, f32_3)
