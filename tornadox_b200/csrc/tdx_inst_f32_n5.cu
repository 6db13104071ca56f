// explicit instantiation of the step kernels for dtype=float, n = nu + 1 = 5
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(float, 5, f32_5)
