// explicit instantiation of the step kernels for dtype=float, n = nu + 1 = 2
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(float, 2, f32_2)
