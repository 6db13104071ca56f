// explicit instantiation of the step kernels for dtype=double, n = nu + 1 = 2
#include "tdx_ops.cuh"
TDX_DEFINE_OPS(double, 2, f64_2)
