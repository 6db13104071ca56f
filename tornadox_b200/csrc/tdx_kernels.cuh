// tornadox_b200 kernels: fused DiagonalEK1 step, KroneckerEK0 sweeps, halo packing, f evaluation.
// One thread per state coordinate, everything in registers; AoS factor blocks are streamed through
// shared memory with TMA 1-D bulk copies (cp.async.bulk), one 32-block segment per warp.
#pragma once

#include "tdx_device.cuh"

namespace tdx {

#ifndef TDX_MIN_BLOCKS
#define TDX_MIN_BLOCKS 2
#endif

template <typename T>
__device__ __forceinline__ T tabs(T x) { return x < (T)0 ? -x : x; }
template <>
__device__ __forceinline__ double tabs<double>(double x) { return fabs(x); }
template <>
__device__ __forceinline__ float tabs<float>(float x) { return fabsf(x); }

// NaN-propagating maximum (jnp.maximum semantics, ek1.py:251)
template <typename T>
__device__ __forceinline__ T nanmax(T a, T b) { return (a > b || a != a) ? a : b; }

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
struct SmemLayout {
    static constexpr int NN = N * N;
    // odd stride (in elements) => conflict-free one-block-per-lane access; even n^2 gets one pad element
    static constexpr int STRIDE = (NN % 2 == 1) ? NN : NN + 1;
    static constexpr int SC_ELEMS = kWarps * 32 * STRIDE;
    static constexpr size_t BAR_BYTES = 128;                       // kWarps mbarriers, padded
    static constexpr size_t RED_BYTES = 128;                       // kWarps doubles, padded
    __host__ __device__ static constexpr size_t sc_bytes() { return ((size_t)SC_ELEMS * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t at_bytes(int at_elems) { return ((size_t)at_elems * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t total(int at_elems, bool with_sc) {
        return BAR_BYTES + RED_BYTES + at_bytes(at_elems) + (with_sc ? sc_bytes() : 0);
    }
};

// ---------------------------------------------------------------------------------------------
// front end shared by all step kernels: load the mean column, predict, publish the predicted state
// to the tile, evaluate f and diag(J).      ek1.py:232,262-265,302-312 / ek0.py:161-166,142-150
// ---------------------------------------------------------------------------------------------
template <class IVP, typename T, int N>
__device__ __forceinline__ void front_end(const Geo& g, const Consts<T, N>& c, const T* __restrict__ mean_in,
                                          const T* halo_lo, const T* halo_hi, const Coord& co, T* s_at, T& raw0,
                                          T (&mp)[N], T& fx, T& jac) {
    using Shape = typename IVP::Shape;
    T raw[N];
#pragma unroll
    for (int k = 0; k < N; ++k) raw[k] = co.valid ? __ldg(mean_in + (long long)k * g.d + co.flat) : (T)0;
    MeanAt<T, N> at{mean_in, g.d, &c};
    const Ext<T> ext = IVP::prefetch(g, co, at, halo_lo, halo_hi);
    predict_column<T, N>(raw, c, mp);
    raw0 = raw[0];
    const T own = c.p[0] * mp[0];
    s_at[at_index<Shape>(co.field, co.tr, co.tcol)] = own;
    __syncthreads();
    fx = (T)0;
    jac = (T)0;
    if (co.valid) IVP::eval(g, co, s_at, ext, own, fx, jac);
    if (c.flags & 1u) jac = (T)0;   // TDX_STEP_ZERO_JACOBIAN: DiagonalEK0, ek0.py:196-280
}

// ---------------------------------------------------------------------------------------------
// DiagonalEK1 fused attempt
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
struct DekArgs {
    Geo g;
    Consts<T, N> c;
    const T* mean_in;
    const T* sc_in;
    T* mean_out;
    T* sc_out;
    T* err_out;
    T* ref_out;
    const T* halo_lo;
    const T* halo_hi;
    double* partials;
    unsigned int* ticket;
    double* out_sum;
};

template <class IVP, typename T, int N>
__global__ void __launch_bounds__(kThreads, TDX_MIN_BLOCKS) dek1_kernel(const __grid_constant__ DekArgs<T, N> a) {
    using Shape = typename IVP::Shape;
    using L = SmemLayout<T, N>;
    constexpr int NN = L::NN, STRIDE = L::STRIDE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_red = reinterpret_cast<double*>(smem_raw + L::BAR_BYTES);
    T* s_at = reinterpret_cast<T*>(smem_raw + L::BAR_BYTES + L::RED_BYTES);
    T* s_sc = reinterpret_cast<T*>(smem_raw + L::BAR_BYTES + L::RED_BYTES + L::at_bytes(Shape::AT_ELEMS));

    const Geo& g = a.g;
    const Consts<T, N>& c = a.c;
    const Coord co = tile_coord<Shape>(g);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* my_sc = s_sc + w * 32 * STRIDE;
    uint64_t* bar = bars + w;

    // (1) start streaming this warp's 32 factor blocks into shared memory
    const long long sc_off = co.seg_base * NN;
    const uint32_t seg_bytes = (uint32_t)(co.seg_len * NN * sizeof(T));
    const bool full = (STRIDE == NN) && (co.seg_len == 32);
    const bool bulk_in = full && ((reinterpret_cast<uintptr_t>(a.sc_in + sc_off) & 15) == 0);
    const bool bulk_out = full && ((reinterpret_cast<uintptr_t>(a.sc_out + sc_off) & 15) == 0);
    if (bulk_in) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            fence_proxy_async();
            mbar_expect_tx(bar, seg_bytes);
            bulk_g2s(my_sc, a.sc_in + sc_off, seg_bytes, bar);
        }
    } else {
        const int total = co.seg_len * NN;
        for (int e = lane; e < total; e += 32) {
            const int blk = e / NN;
            my_sc[blk * STRIDE + (e - blk * NN)] = __ldg(a.sc_in + sc_off + e);
        }
    }

    // (2) mean column, prediction, vector field
    T raw0, mp[N], fx, jac;
    front_end<IVP, T, N>(g, c, a.mean_in, a.halo_lo, a.halo_hi, co, s_at, raw0, mp, fx, jac);

    if (bulk_in) mbar_wait(bar, 0);
    __syncwarp();

    double ratio_sq = 0.0;
    if (co.valid) {
        const T z = c.p[1] * mp[1] - fx;

        // estimate_error / local diffusion  ek1.py:314-331.  Ql rows 0 and 1 have 1 and 2 nonzeros.
        const T h0 = c.p[1] * QL(c, 1, 0) - jac * (c.p[0] * QL(c, 0, 0));
        const T h1 = c.p[1] * QL(c, 1, 1);
        const T s = h0 * h0 + h1 * h1;
        const T sqs = tsqrt<T>(s);
        const T sigma = tabs<T>(z / sqs);
        const T err = sigma * sqs;

        // B = A (p_inv * sc)   ek1.py:233,269
        T B[N][N];
        const T* blk = my_sc + lane * STRIDE;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            T col[N];
#pragma unroll
            for (int k = 0; k < N; ++k) col[k] = c.pinv[k] * blk[k * N + j];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                T acc = col[i];
#pragma unroll
                for (int k = i + 1; k < N; ++k) acc = fma((T)A_entry<N>(i, k), col[k], acc);
                B[i][j] = acc;
            }
        }

        // predict_cov_sqrtm: QR of [(A sc)^T ; (sigma Ql)^T]   ek1.py:267-270, sqrt.py:8-30
        stacked_qr<T, N>(B, sigma, c);

        // observe_cov_sqrtm  ek1.py:333-349 (L = lower(B); rows 0,1 have 1 and 2 nonzeros)
        const T g0 = c.p[1] * B[1][0] - jac * (c.p[0] * B[0][0]);
        const T g1 = c.p[1] * B[1][1];
        const T s2 = (g0 * g0 + g1 * g1) + (T)1e-16;
        const T inv_s2 = (T)1 / s2;

        // correct_cov_sqrtm / correct_mean  ek1.py:351-369, un-precondition ek1.py:246-247
        T* outblk = my_sc + lane * STRIDE;
        T new_m0 = (T)0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const T l0 = B[r][0];
            const T l1 = (r >= 1) ? B[r][1] : (T)0;
            const T kg = (l0 * g0 + l1 * g1) * inv_s2;
            const T pr = c.p[r];
            outblk[r * N + 0] = pr * (l0 - kg * g0);
            outblk[r * N + 1] = pr * (l1 - kg * g1);
#pragma unroll
            for (int cc = 2; cc < N; ++cc) outblk[r * N + cc] = (r >= cc) ? pr * B[r][cc] : (T)0;
            const T mo = pr * (mp[r] - kg * z);
            a.mean_out[(long long)r * g.d + co.flat] = mo;
            if (r == 0) new_m0 = mo;
        }
        const T ref = nanmax<T>(tabs<T>(raw0), tabs<T>(new_m0));   // ek1.py:249-251
        if (a.err_out) a.err_out[co.flat] = err;
        if (a.ref_out) a.ref_out[co.flat] = ref;
        if (c.want_norm) {
            const double ratio = ((double)c.dt * (double)err) / ((double)c.abstol + (double)c.reltol * (double)ref);
            ratio_sq = ratio * ratio;                              // step.py:101-104
        }
    }

    // (3) stream the updated blocks back
    if (bulk_out) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(a.sc_out + sc_off, my_sc, seg_bytes);
            bulk_commit();
        }
    } else {
        __syncwarp();
        const int total = co.seg_len * NN;
        for (int e = lane; e < total; e += 32) {
            const int blk = e / NN;
            a.sc_out[sc_off + e] = my_sc[blk * STRIDE + (e - blk * NN)];
        }
    }

    if (c.want_norm) grid_sum(ratio_sq, s_red, a.partials, a.ticket, a.out_sum);
    if (bulk_out && lane == 0) bulk_wait_read0();
}

// ---------------------------------------------------------------------------------------------
// KroneckerEK0 sweeps
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
struct Ek0Mid {          // written by ek0_mid_kernel, read by sweep B
    T K[kMaxN];          // gain  ek0.py:136
    double err;          // calibrated scalar error estimate  ek0.py:129
};

template <typename T, int N>
struct Ek0Args {
    Geo g;
    Consts<T, N> c;
    const T* mean_in;
    T* mean_out;
    T* ref_out;
    const T* halo_lo;
    const T* halo_hi;
    const Ek0Mid<T, N>* mid;
    double* partials;
    unsigned int* ticket;
    double* out_sum;
};

// sweep A: z = p1 mp[1] - f(p0 mp[0]);  sum z^2        ek0.py:166-172
template <class IVP, typename T, int N>
__global__ void __launch_bounds__(kThreads) ek0_sweep_a_kernel(const __grid_constant__ Ek0Args<T, N> a) {
    using Shape = typename IVP::Shape;
    using L = SmemLayout<T, N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_red = reinterpret_cast<double*>(smem_raw + L::BAR_BYTES);
    T* s_at = reinterpret_cast<T*>(smem_raw + L::BAR_BYTES + L::RED_BYTES);
    const Coord co = tile_coord<Shape>(a.g);
    T raw0, mp[N], fx, jac;
    front_end<IVP, T, N>(a.g, a.c, a.mean_in, a.halo_lo, a.halo_hi, co, s_at, raw0, mp, fx, jac);
    double zz = 0.0;
    if (co.valid) {
        const T z = a.c.p[1] * mp[1] - fx;
        zz = (double)z * (double)z;
    }
    grid_sum(zz, s_red, a.partials, a.ticket, a.out_sum);
}

// sweep B: m_new = P (mp - K z); ref = |m_new[0]|; sum 1/(abstol + reltol ref)^2     ek0.py:132-140,181-184
template <class IVP, typename T, int N>
__global__ void __launch_bounds__(kThreads) ek0_sweep_b_kernel(const __grid_constant__ Ek0Args<T, N> a) {
    using Shape = typename IVP::Shape;
    using L = SmemLayout<T, N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_red = reinterpret_cast<double*>(smem_raw + L::BAR_BYTES);
    T* s_at = reinterpret_cast<T*>(smem_raw + L::BAR_BYTES + L::RED_BYTES);
    const Coord co = tile_coord<Shape>(a.g);
    T raw0, mp[N], fx, jac;
    front_end<IVP, T, N>(a.g, a.c, a.mean_in, a.halo_lo, a.halo_hi, co, s_at, raw0, mp, fx, jac);
    double acc = 0.0;
    if (co.valid) {
        const T z = a.c.p[1] * mp[1] - fx;
        T m0 = (T)0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const T mo = a.c.p[r] * (mp[r] - a.mid->K[r] * z);
            a.mean_out[(long long)r * a.g.d + co.flat] = mo;
            if (r == 0) m0 = mo;
        }
        const T ref = tabs<T>(m0);
        if (a.ref_out) a.ref_out[co.flat] = ref;
        if (a.c.want_norm) {
            const double tol = (double)a.c.abstol + (double)a.c.reltol * (double)ref;
            acc = 1.0 / (tol * tol);
        }
    }
    // sum_i (dt err / tol_i)^2 = (dt err)^2 sum_i 1 / tol_i^2     step.py:101-104 with a scalar error estimate
    if (a.c.want_norm)
        grid_sum(acc, s_red, a.partials, a.ticket, a.out_sum, (double)a.c.dt * (double)a.c.dt, &a.mid->err);
}

// mid: calibration, ONE stacked QR, gain, factor update; a single thread in fp64.   ek0.py:123-140,175,182
template <typename T, int N>
struct Ek0MidArgs {
    Consts<double, N> c;
    const T* Cl_in;
    const double* z_sq_sum;
    double d_global;
    T* Cl_out;
    T* err_out;
    Ek0Mid<T, N>* mid;
};

template <typename T, int N>
__global__ void ek0_mid_kernel(const __grid_constant__ Ek0MidArgs<T, N> a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Consts<double, N>& c = a.c;
    const double p1 = c.p[1];
    const double Q11 = QL(c, 1, 0) * QL(c, 1, 0) + QL(c, 1, 1) * QL(c, 1, 1);
    const double HQH = p1 * p1 * Q11;
    const double sigma_sq = (*a.z_sq_sum) / HQH / a.d_global;
    const double err = sqrt(sigma_sq * HQH);
    const double sigma = sqrt(sigma_sq);
    double B[N][N];
    for (int j = 0; j < N; ++j) {
        double col[N];
        for (int k = 0; k < N; ++k) col[k] = c.pinv[k] * (double)a.Cl_in[k * N + j];
        for (int i = 0; i < N; ++i) {
            double acc = col[i];
            for (int k = i + 1; k < N; ++k) acc = fma(A_entry<N>(i, k), col[k], acc);
            B[i][j] = acc;
        }
    }
    stacked_qr<double, N>(B, sigma, c);
    const double g0 = p1 * B[1][0], g1 = p1 * B[1][1];   // H Clp = p1 * Clp[1, :]
    const double S = p1 * p1 * (B[1][0] * B[1][0] + B[1][1] * B[1][1]);
    for (int r = 0; r < N; ++r) {
        const double l0 = B[r][0], l1 = (r >= 1) ? B[r][1] : 0.0;
        const double kg = (l0 * g0 + l1 * g1) / S;
        a.mid->K[r] = (T)kg;
        const double pr = c.p[r];
        a.Cl_out[r * N + 0] = (T)(pr * (l0 - kg * g0));
        a.Cl_out[r * N + 1] = (T)(pr * (l1 - kg * g1));
        for (int cc = 2; cc < N; ++cc) a.Cl_out[r * N + cc] = (r >= cc) ? (T)(pr * B[r][cc]) : (T)0;
    }
    *a.err_out = (T)err;
    a.mid->err = err;
}

// ---------------------------------------------------------------------------------------------
// halo packing: predicted state of the shard's first / last rows (fhn) or points (1-d chains)
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
struct PackArgs {
    Geo g;
    Consts<T, N> c;
    const T* mean_in;
    T* edge_first;
    T* edge_last;
    int n_first, n_last;   // points per field taken from the start / end of the shard
};

template <typename T, int N>
__global__ void pack_halo_kernel(const __grid_constant__ PackArgs<T, N> a) {
    MeanAt<T, N> at{a.mean_in, a.g.d, &a.c};
    const int per_field = a.n_first + a.n_last;
    const int total = a.g.nf * per_field;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int field = i / per_field, j = i - field * per_field;
        if (j < a.n_first) {
            a.edge_first[field * a.n_first + j] = at((long long)field * a.g.npts + j);
        } else {
            const int jj = j - a.n_first;
            a.edge_last[field * a.n_last + jj] = at((long long)field * a.g.npts + (a.g.npts - a.n_last + jj));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// f(y) and diag J(y) on a plain state vector
// ---------------------------------------------------------------------------------------------
template <typename T>
struct EvalArgs {
    Geo g;
    const T* y;
    T* fy;
    T* jd;
    const T* halo_lo;
    const T* halo_hi;
};

template <class IVP, typename T>
__global__ void __launch_bounds__(kThreads) eval_f_kernel(const __grid_constant__ EvalArgs<T> a) {
    using Shape = typename IVP::Shape;
    __shared__ T s_at[Shape::AT_ELEMS];
    const Coord co = tile_coord<Shape>(a.g);
    PlainAt<T> at{a.y};
    const T own = co.valid ? __ldg(a.y + co.flat) : (T)0;
    const Ext<T> ext = IVP::prefetch(a.g, co, at, a.halo_lo, a.halo_hi);
    s_at[at_index<Shape>(co.field, co.tr, co.tcol)] = own;
    __syncthreads();
    if (co.valid) {
        T fx, jac;
        IVP::eval(a.g, co, s_at, ext, own, fx, jac);
        a.fy[co.flat] = fx;
        if (a.jd) a.jd[co.flat] = jac;
    }
}

}  // namespace tdx
