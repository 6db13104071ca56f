// explicit instantiation of the step kernels for dtype=double, n = nu + 1 = 5
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(double, 5, f64_5)
