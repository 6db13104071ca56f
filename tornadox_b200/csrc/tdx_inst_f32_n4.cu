// explicit instantiation of the step kernels for dtype=float, n = nu + 1 = 4
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(float, 4, f32_4)
