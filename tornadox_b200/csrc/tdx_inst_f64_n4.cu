// explicit instantiation of the step kernels for dtype=double, n = nu + 1 = 4
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(double, 4, f64_4)
