// explicit instantiation of the step kernels for dtype=double, n = nu + 1 = 3
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(double, 3, f64_3)
