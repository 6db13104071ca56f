// explicit instantiation of the step kernels for dtype=double, n = nu + 1 = 6
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(double, 6, f64_6)
