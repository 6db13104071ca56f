// f / diag(J) evaluation kernels (tdx_eval_f), independent of the number of derivatives
#include "tdx_ops.cuh"

namespace tdx {

template <typename T>
static int eval_f_launch(const Geo& g, const void* y, void* fy, void* jd, const void* halo_lo, const void* halo_hi,
                         cudaStream_t st) {
    EvalArgs<T> a;
    a.g = g;
    a.y = (const T*)y;
    a.fy = (T*)fy;
    a.jd = (T*)jd;
    a.halo_lo = (const T*)halo_lo;
    a.halo_hi = (const T*)halo_hi;
    TDX_SWITCH_IVP(g.kind, T, {
        using Shape = typename IVP::Shape;
        const long long grid = num_tiles<Shape>(g.rows, g.cols);
        a.tiles_x = (unsigned)((g.cols + Shape::TC - 1) / Shape::TC);
        eval_f_kernel<IVP, T><<<(unsigned)grid, kThreads, 0, st>>>(a);
    })
    return (int)cudaGetLastError();
}

const EvalOps* eval_ops_f64() {
    static const EvalOps o{&eval_f_launch<double>};
    return &o;
}
const EvalOps* eval_ops_f32() {
    static const EvalOps o{&eval_f_launch<float>};
    return &o;
}

}  // namespace tdx
