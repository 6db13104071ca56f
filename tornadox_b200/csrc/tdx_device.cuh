// tornadox_b200 device code: shared building blocks of the fused filter-step kernels (sm_100a).
//
// Reference semantics (paths into the upstream repository, see include/tornadox_b200.h):
//   iwp.py:19-37,65-73   prior constants / Nordsieck preconditioner
//   sqrt.py:8-30         propagate_cholesky_factor == R-factor of a stacked (2n x n) matrix (LAPACK geqrf signs)
//   ek1.py:228-369       DiagonalEK1 attempt_step        ek0.py:152-193  KroneckerEK0 attempt_step
//   ivp.py               vanderpol / lorenz96 / brusselator / fhn_2d right-hand sides and Jacobian diagonals
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace tdx {

constexpr int kMaxN = 7;       // n = nu + 1 <= 7
constexpr int kWarps = 8;      // warps per CTA; every warp owns one 32-wide segment of one field
constexpr int kThreads = kWarps * 32;

enum : int { IVP_VDP = 0, IVP_LORENZ96 = 1, IVP_BRUSS = 2, IVP_FHN2D = 3 };

// Local shard geometry.  Flat local state index = field * npts + row * cols + col.
struct Geo {
    int kind, nf;
    int rows, cols;             // local extent in points
    long long npts, d;          // rows * cols,  nf * npts
    long long grow0, gcol0;     // global coordinates of the local (0, 0)
    long long grows, gcols;     // global extent
    double prm[8];
};

// Per-attempt constants, passed by value (lands in the constant bank).
template <typename T, int N>
struct Consts {
    T p[N];                      // Nordsieck preconditioner        iwp.py:70-72
    T pinv[N];
    T ql[N * (N + 1) / 2];       // lower Cholesky factor of the preconditioned diffusion, packed rows (iwp.py:36-37)
    T dt, abstol, reltol;
    int want_norm;
    unsigned int flags;          // TDX_STEP_*
};

template <typename T, int N>
__device__ __forceinline__ T QL(const Consts<T, N>& c, int i, int j) {
    return c.ql[i * (i + 1) / 2 + j];
}

// A[i][j] = C(nu - i, nu - j) for j >= i: flip(lower Pascal), iwp.py:31-35 (unit upper triangular).
__host__ __device__ constexpr double binom(int n, int k) {
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    return r;
}
template <int N>
__host__ __device__ constexpr double A_entry(int i, int j) {
    return (j >= i) ? binom(N - 1 - i, N - 1 - j) : 0.0;
}

// ---------------------------------------------------------------------------------------------
// mean prediction for one state coordinate:  mp = A (p_inv * raw)      ek1.py:232,262-265 / ek0.py:162-166
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
__device__ __forceinline__ void predict_column(const T (&raw)[N], const Consts<T, N>& c, T (&mp)[N]) {
    T m[N];
#pragma unroll
    for (int k = 0; k < N; ++k) m[k] = c.pinv[k] * raw[k];
#pragma unroll
    for (int r = 0; r < N; ++r) {
        T acc = m[r];
#pragma unroll
        for (int k = r + 1; k < N; ++k) acc = fma((T)A_entry<N>(r, k), m[k], acc);
        mp[r] = acc;
    }
}

// Accessor: predicted, un-preconditioned state p[0] * (A p_inv m)[0] at an arbitrary local flat index,
// recomputed from the mean in global memory (bitwise identical to the owner thread's own value).
template <typename T, int N>
struct MeanAt {
    const T* __restrict__ mean;
    long long d;
    const Consts<T, N>* c;
    __device__ __forceinline__ T operator()(long long idx) const {
        T acc = c->pinv[0] * __ldg(mean + idx);
#pragma unroll
        for (int k = 1; k < N; ++k)
            acc = fma((T)A_entry<N>(0, k), c->pinv[k] * __ldg(mean + (long long)k * d + idx), acc);
        return c->p[0] * acc;
    }
};

// Accessor over a plain state vector (tdx_eval_f).
template <typename T>
struct PlainAt {
    const T* __restrict__ y;
    __device__ __forceinline__ T operator()(long long idx) const { return __ldg(y + idx); }
};

// ---------------------------------------------------------------------------------------------
// Tile mapping.  A CTA covers TR grid rows x (32 * SEGS) columns of all NF fields; warp w owns the
// 32-wide segment (field, tile row, seg).  NF * TR * SEGS == kWarps.
// ---------------------------------------------------------------------------------------------
template <int NF_, int TR_, int SEGS_>
struct TileShape {
    static constexpr int NF = NF_, TR = TR_, SEGS = SEGS_;
    static constexpr int TC = 32 * SEGS_;
    static_assert(NF_ * TR_ * SEGS_ == kWarps, "tile shape must use all warps");
    static constexpr int AT_ELEMS = NF_ * TR_ * TC;
};

struct Coord {
    int field, tr, tcol;   // tile-local
    int row, col;          // shard-local
    bool valid;
    long long flat;        // field * npts + row * cols + col
    long long seg_base;    // flat index of lane 0 of this warp's segment
    int seg_len;           // valid lanes in the segment (0..32)
};

template <class Shape>
__device__ __forceinline__ Coord tile_coord(const Geo& g) {
    Coord c;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (g.cols + Shape::TC - 1) / Shape::TC;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    c.field = w / (Shape::TR * Shape::SEGS);
    const int rem = w - c.field * (Shape::TR * Shape::SEGS);
    c.tr = rem / Shape::SEGS;
    const int seg = rem - c.tr * Shape::SEGS;
    c.tcol = seg * 32 + lane;
    c.row = tile_y * Shape::TR + c.tr;
    const int col0 = tile_x * Shape::TC + seg * 32;
    c.col = col0 + lane;
    const bool row_ok = c.row < g.rows;
    c.valid = row_ok && (c.col < g.cols);
    c.seg_base = (long long)c.field * g.npts + (long long)c.row * g.cols + col0;
    c.flat = c.seg_base + lane;
    int len = g.cols - col0;
    len = len < 0 ? 0 : (len > 32 ? 32 : len);
    c.seg_len = row_ok ? len : 0;
    return c;
}

template <class Shape>
__host__ __device__ inline long long num_tiles(int rows, int cols) {
    return (long long)((cols + Shape::TC - 1) / Shape::TC) * ((rows + Shape::TR - 1) / Shape::TR);
}

template <class Shape>
__device__ __forceinline__ int at_index(int field, int tr, int tcol) {
    return (field * Shape::TR + tr) * Shape::TC + tcol;
}

// ---------------------------------------------------------------------------------------------
// Vector fields.  Each provides
//   Shape                       tile shape
//   prefetch(...)  -> Ext       out-of-tile neighbour values, fetched BEFORE the tile barrier
//   eval(...)                   f_i and (df/dy)_ii from the shared tile + Ext, AFTER the barrier
// halo_lo / halo_hi hold the predicted state of the neighbouring shards (or nullptr: single shard).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Ext {
    T a, b, c, e;
};

// --- vanderpol: ivp.py:28-49.  One point, two fields (y1, y2). ---------------------------------
template <typename T>
struct VanderPol {
    using Shape = TileShape<2, 1, 4>;
    static __host__ void halo_sizes(const Geo&, long long& lo, long long& hi) { lo = hi = 0; }
    template <class At>
    __device__ static __forceinline__ Ext<T> prefetch(const Geo&, const Coord&, const At&, const T*, const T*) {
        return Ext<T>{};
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_at, const Ext<T>&, T own,
                                                T& fx, T& jac) {
        const T mu = (T)g.prm[0];
        const T y1 = s_at[at_index<Shape>(0, 0, c.tcol)];
        const T y2 = s_at[at_index<Shape>(1, 0, c.tcol)];
        if (c.field == 0) {
            fx = y2;
            jac = (T)0;
        } else {
            fx = mu * ((T)1 - y1 * y1) * y2 - y1;
            jac = mu * ((T)1 - y1 * y1);
        }
        (void)own;
    }
};

// --- lorenz96: ivp.py:268-294.  f_i = (y[i+1] - y[i-2]) y[i-1] - y[i] + F, cyclic; diag J = -1. ---
template <typename T>
struct Lorenz96 {
    using Shape = TileShape<1, 1, 8>;
    // lower halo = points (i0-2, i0-1), upper halo = point i0+cols (all cyclic)
    static __host__ void halo_sizes(const Geo& g, long long& lo, long long& hi) {
        const bool sharded = g.cols != g.gcols;
        lo = sharded ? 2 : 0;
        hi = sharded ? 1 : 0;
    }
    template <class At>
    __device__ static __forceinline__ T outside(const Geo& g, const Coord& c, const At& at, const T* halo_lo,
                                                const T* halo_hi, int off) {
        // cyclic neighbour col + off that is not inside the tile
        long long gi = g.gcol0 + c.col + off;
        if (gi < 0) gi += g.gcols;
        if (gi >= g.gcols) gi -= g.gcols;
        const long long li = gi - g.gcol0;
        if (li >= 0 && li < g.cols) return at(li);
        // not owned: one of the exchanged halo values
        if (off < 0) {
            const long long k = (long long)c.col + off;  // -2 or -1
            return halo_lo[2 + k];
        }
        return halo_hi[0];
    }
    template <class At>
    __device__ static __forceinline__ Ext<T> prefetch(const Geo& g, const Coord& c, const At& at, const T* halo_lo,
                                                      const T* halo_hi) {
        Ext<T> e{};
        if (!c.valid) return e;
        if (c.tcol - 2 < 0) e.a = outside(g, c, at, halo_lo, halo_hi, -2);
        if (c.tcol - 1 < 0) e.b = outside(g, c, at, halo_lo, halo_hi, -1);
        if (c.tcol + 1 >= Shape::TC || c.col + 1 >= g.cols) e.c = outside(g, c, at, halo_lo, halo_hi, +1);
        return e;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_at, const Ext<T>& e, T own,
                                                T& fx, T& jac) {
        const T F = (T)g.prm[0];
        const T ym2 = (c.tcol - 2 < 0) ? e.a : s_at[c.tcol - 2];
        const T ym1 = (c.tcol - 1 < 0) ? e.b : s_at[c.tcol - 1];
        const T yp1 = (c.tcol + 1 >= Shape::TC || c.col + 1 >= g.cols) ? e.c : s_at[c.tcol + 1];
        fx = (yp1 - ym2) * ym1 - own + F;
        jac = (T)-1;
    }
};

// --- brusselator: ivp.py:219-265.  y = [u(N); v(N)], c = (N+1)^2/50, boundary pads u=1, v=3. -------
template <typename T>
struct Brusselator {
    using Shape = TileShape<2, 1, 4>;
    static __host__ void halo_sizes(const Geo& g, long long& lo, long long& hi) {
        lo = (g.gcol0 > 0) ? 2 : 0;                  // (u, v) of point i0-1
        hi = (g.gcol0 + g.cols < g.gcols) ? 2 : 0;   // (u, v) of point i0+cols
    }
    template <class At>
    __device__ static __forceinline__ Ext<T> prefetch(const Geo& g, const Coord& c, const At& at, const T* halo_lo,
                                                      const T* halo_hi) {
        Ext<T> e{};
        if (!c.valid) return e;
        const T pad = (c.field == 0) ? (T)1 : (T)3;
        const long long gi = g.gcol0 + c.col;
        if (gi == 0) e.a = pad;
        else if (c.tcol == 0) e.a = (c.col > 0) ? at(c.flat - 1) : halo_lo[c.field];
        if (gi == g.gcols - 1) e.b = pad;
        else if (c.tcol + 1 >= Shape::TC || c.col + 1 >= g.cols)
            e.b = (c.col + 1 < g.cols) ? at(c.flat + 1) : halo_hi[c.field];
        return e;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_at, const Ext<T>& e, T own,
                                                T& fx, T& jac) {
        const T cc = (T)g.prm[0];
        const long long gi = g.gcol0 + c.col;
        const bool ext_l = (gi == 0) || (c.tcol == 0);
        const bool ext_r = (gi == g.gcols - 1) || (c.tcol + 1 >= Shape::TC) || (c.col + 1 >= g.cols);
        const T left = ext_l ? e.a : s_at[at_index<Shape>(c.field, 0, c.tcol - 1)];
        const T right = ext_r ? e.b : s_at[at_index<Shape>(c.field, 0, c.tcol + 1)];
        const T u = s_at[at_index<Shape>(0, 0, c.tcol)];
        const T v = s_at[at_index<Shape>(1, 0, c.tcol)];
        const T conv = (left - (T)2 * own) + right;
        if (c.field == 0) {
            fx = (T)1 + u * u * v - (T)4 * u + cc * conv;
            jac = (T)2 * u * v - (T)4 - (T)2 * cc;
        } else {
            fx = (T)3 * u - u * u * v + cc * conv;
            jac = -(u * u) - (T)2 * cc;
        }
    }
};

// --- fhn_2d: ivp.py:409-505.  y = [u; v] on an (nx, ny) grid, 5-point Laplacian, Neumann by edge
//     replication; u' = a Lap u + u - u^3 - v + k ; v' = (b Lap v + u - v) / tau. --------------------
template <typename T>
struct Fhn2d {
    using Shape = TileShape<2, 4, 1>;
    static __host__ void halo_sizes(const Geo& g, long long& lo, long long& hi) {
        lo = (g.grow0 > 0) ? 2LL * g.cols : 0;                 // row above, both fields
        hi = (g.grow0 + g.rows < g.grows) ? 2LL * g.cols : 0;  // row below
    }
    struct Where {
        bool up_rep, dn_rep, lf_rep, rt_rep;   // replicate own value (domain boundary)
        bool up_ext, dn_ext, lf_ext, rt_ext;   // value comes from outside the tile
    };
    __device__ static __forceinline__ Where where(const Geo& g, const Coord& c) {
        Where w;
        const long long gr = g.grow0 + c.row;
        w.up_rep = (gr == 0);
        w.dn_rep = (gr == g.grows - 1);
        w.lf_rep = (c.col == 0);
        w.rt_rep = (c.col == g.cols - 1);
        w.up_ext = !w.up_rep && (c.tr == 0);
        w.dn_ext = !w.dn_rep && (c.tr + 1 >= Shape::TR || c.row + 1 >= g.rows);
        w.lf_ext = !w.lf_rep && (c.tcol == 0);
        w.rt_ext = !w.rt_rep && (c.tcol + 1 >= Shape::TC);
        return w;
    }
    template <class At>
    __device__ static __forceinline__ Ext<T> prefetch(const Geo& g, const Coord& c, const At& at, const T* halo_lo,
                                                      const T* halo_hi) {
        Ext<T> e{};
        if (!c.valid) return e;
        const Where w = where(g, c);
        if (w.up_ext) e.a = (c.row > 0) ? at(c.flat - g.cols) : halo_lo[(long long)c.field * g.cols + c.col];
        if (w.dn_ext) e.b = (c.row + 1 < g.rows) ? at(c.flat + g.cols) : halo_hi[(long long)c.field * g.cols + c.col];
        if (w.lf_ext) e.c = at(c.flat - 1);
        if (w.rt_ext) e.e = at(c.flat + 1);
        return e;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_at, const Ext<T>& e, T own,
                                                T& fx, T& jac) {
        const T inv_dx2 = (T)g.prm[0], a = (T)g.prm[1], b = (T)g.prm[2], k = (T)g.prm[3], tau = (T)g.prm[4];
        const Where w = where(g, c);
        const T up = w.up_rep ? own : (w.up_ext ? e.a : s_at[at_index<Shape>(c.field, c.tr - 1, c.tcol)]);
        const T dn = w.dn_rep ? own : (w.dn_ext ? e.b : s_at[at_index<Shape>(c.field, c.tr + 1, c.tcol)]);
        const T lf = w.lf_rep ? own : (w.lf_ext ? e.c : s_at[at_index<Shape>(c.field, c.tr, c.tcol - 1)]);
        const T rt = w.rt_rep ? own : (w.rt_ext ? e.e : s_at[at_index<Shape>(c.field, c.tr, c.tcol + 1)]);
        const T lap = (((up + dn) + (lf + rt)) - (T)4 * own) * inv_dx2;
        const int nrep = (int)w.up_rep + (int)w.dn_rep + (int)w.lf_rep + (int)w.rt_rep;
        const T dl = -(T)(4 - nrep) * inv_dx2;   // ivp.py:447-463: 2 (corner), 3 (edge), 4 (interior)
        const T u = s_at[at_index<Shape>(0, c.tr, c.tcol)];
        const T v = s_at[at_index<Shape>(1, c.tr, c.tcol)];
        if (c.field == 0) {
            fx = a * lap + u - u * u * u - v + k;
            jac = a * dl + (T)1 - (T)3 * (u * u);
        } else {
            fx = (b * lap + u - v) / tau;
            jac = (b * dl - (T)1) / tau;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Householder QR of the stacked matrix [ (A sc)^T ; (sigma Ql)^T ]  (2n x n), LAPACK dgeqr2/dlarfg
// convention: beta = -sign(alpha) * ||(alpha, x)||, a column with x == 0 is left untouched.
//
// Storage is by COLUMN of the stacked matrix: column c = [ B[c][0..n-1] ; G[c][0..c] ] where
// B = A sc (row c of B is the top part of column c) and G = sigma * Ql (lower triangular: the bottom
// block (sigma Ql)^T is upper triangular and stays so, there is no fill-in).  On exit the lower
// triangle of B holds L = R^T (the propagated Cholesky factor, sqrt.py:8-23).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T tsqrt(T x);
template <>
__device__ __forceinline__ double tsqrt<double>(double x) { return sqrt(x); }
template <>
__device__ __forceinline__ float tsqrt<float>(float x) { return sqrtf(x); }

template <typename T, int N>
__device__ __forceinline__ void stacked_qr(T (&B)[N][N], T sigma, const Consts<T, N>& c) {
    T G[N][N];   // only j <= i used; materialised lazily (row k of the bottom block is first touched at step k)
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int cc = k; cc < N; ++cc) G[cc][k] = sigma * QL(c, cc, k);
        const T alpha = B[k][k];
        T ns = (T)0;
#pragma unroll
        for (int r = k + 1; r < N; ++r) ns = fma(B[k][r], B[k][r], ns);
#pragma unroll
        for (int r = 0; r <= k; ++r) ns = fma(G[k][r], G[k][r], ns);
        if (ns != (T)0) {
            const T nrm = tsqrt<T>(fma(alpha, alpha, ns));
            const T beta = (alpha >= (T)0 && !(alpha == (T)0 && signbit(alpha))) ? -nrm : nrm;   // -copysign(nrm, alpha)
            const T u0 = alpha - beta;
            const T gamma = (T)1 / (beta * (beta - alpha));
#pragma unroll
            for (int cc = k + 1; cc < N; ++cc) {
                T w = u0 * B[cc][k];
#pragma unroll
                for (int r = k + 1; r < N; ++r) w = fma(B[k][r], B[cc][r], w);
#pragma unroll
                for (int r = 0; r <= k; ++r) w = fma(G[k][r], G[cc][r], w);
                w *= gamma;
                B[cc][k] = fma(-w, u0, B[cc][k]);
#pragma unroll
                for (int r = k + 1; r < N; ++r) B[cc][r] = fma(-w, B[k][r], B[cc][r]);
#pragma unroll
                for (int r = 0; r <= k; ++r) G[cc][r] = fma(-w, G[k][r], G[cc][r]);
            }
            B[k][k] = beta;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA 1-D bulk copies + mbarrier (raw PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Deterministic grid-wide sum of one double per thread: warp shuffle -> smem -> per-CTA partial ->
// last-arriving CTA adds the partials in a fixed order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The final value is sum * mult * (*mult_sqrt_ptr)^2 (the last two default to 1).
__device__ __forceinline__ void grid_sum(double v, double* s_red, double* partials, unsigned int* ticket,
                                         double* out, double mult = 1.0, const double* mult_sqrt_ptr = nullptr) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum(v);
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) b += s_red[i];
        partials[blockIdx.x] = b;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double acc = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) acc += __ldcg(partials + i);
        acc = warp_sum(acc);
        if (lane == 0) s_red[w] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = 0.0;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) b += s_red[i];
            if (mult_sqrt_ptr) {
                const double r = *mult_sqrt_ptr;
                b *= r * r;
            }
            *out = b * mult;
            *ticket = 0u;
        }
    }
}

}  // namespace tdx
