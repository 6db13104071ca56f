// KroneckerEK0.attempt_step for fhn_2d as ONE cooperative, persistent launch (ek0.py:152-193).
//
//   phase A   stream the mean once: predict, f, z = p1 mp[1] - f(p0 mp[0]), sum z^2            ek0.py:161-172
//   barrier   the last CTA to arrive adds the partial sums (fixed order), all-reduces them over the ranks
//             (peer memory), runs the calibration + the ONE stacked QR + gain (ek0.py:123-140,175) and releases a flag
//   phase B   stream the mean again IN REVERSE (what phase A touched last is still in the 126 MB L2):
//             m_new = P (mp - K z), reference_state = |m_new[0]|, sum (dt err / tol_i)^2        ek0.py:132-140,181-184
//
// Work decomposition: the shard's grid is cut into strips of W columns x (rows / chunks) rows, one strip per CTA
// (more when the grid is wider than the device).  A CTA walks its strip row by row; thread t owns column c0 + t of
// BOTH fields.  The raw mean rows (2 fields x n derivatives x (W + 2 PADC) columns) arrive through a ring of STAGES
// shared-memory stages filled by TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), issued STAGES rows
// ahead by one thread, so HBM latency is hidden by the ring depth and not by registers.  The 5-point stencil needs the
// predicted state of the row above / below (kept in registers: the thread owns the column) and of the left / right
// columns (a double-buffered shared-memory row, ONE __syncthreads per row, which also frees the consumed stage).
// Rows outside the strip: one extra row is streamed on each side (2 / rows-per-strip redundant reads, L2 hits);
// outside the shard: the neighbour GPU's halo buffer, or edge replication (ivp.py:488-490).
#pragma once

#include <type_traits>

#include "tdx_kernels.cuh"

#ifndef TDX_EK0_STAGES
#define TDX_EK0_STAGES 4     // depth of the TMA ring (rows in flight per CTA)
#endif
#ifndef TDX_EK0_MAXREG
#define TDX_EK0_MAXREG 96    // 2 CTAs x 9 warps: 5 warps share one of the four 16 K-register sub-partitions -> <= 102
#endif

namespace tdx {

template <typename T, int N, int W_>
struct Ek0FusedCfg {
    static constexpr int W = W_;                          // columns per strip == consumer threads per CTA
    static constexpr int THREADS = W_ + 32;               // + one producer warp (which also does the n x n algebra)
    static constexpr int PADC = 16 / (int)sizeof(T);      // columns of left/right overlap (16-byte granularity of TMA)
    static constexpr int WP = W + 2 * PADC;
    static constexpr int STAGES = TDX_EK0_STAGES;
    static constexpr int STAGE_ELEMS = 2 * N * WP;        // [field][derivative][WP]
    static constexpr int AT_ROW = W + 2;                  // [left edge][W own][right edge]
    static constexpr int AT_ELEMS = 2 * 2 * AT_ROW;       // [slot][field][AT_ROW]
    static constexpr size_t BAR_BYTES = 128, RED_BYTES = 512;
    static constexpr size_t BUF_BYTES = 8 * (size_t)W_ * sizeof(double);     // a row block's terms per thread (strips that walk against the summation order)
    __host__ __device__ static constexpr size_t at_bytes() { return ((size_t)AT_ELEMS * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t stage_bytes() { return (size_t)STAGE_ELEMS * sizeof(T); }
    __host__ __device__ static constexpr size_t total() { return BAR_BYTES + RED_BYTES + BUF_BYTES + at_bytes() + STAGES * stage_bytes(); }
    static_assert((STAGE_ELEMS * sizeof(T)) % 16 == 0 && (WP * sizeof(T)) % 16 == 0, "TMA destinations stay 16-byte aligned");
    static_assert(W_ % 32 == 0 && W_ <= 256, "one row partial per consumer warp, at most eight");
};

// Layout of the reduction scratch (doubles) behind the mbarriers of the fused kernels
constexpr int kEk0RedScratch = 0;      // [kRedChunks + 2] final reductions
constexpr int kEk0RowRed = 32;         // [2][8] per-unit warp sums, double-buffered

template <typename T, int N>
struct Ek0FusedArgs {
    Geo g;
    Consts<T, N> c;
    Consts<double, N> cd;            // the calibration / QR run in fp64 whatever the state dtype
    // state buffers: a ring of ``nbuf`` (mean, factor) pairs.  Without a controller: in = slot 0, out = slot 1.
    T* mean_ring[kMaxRing];
    T* Cl_ring[kMaxRing];
    T* err_ring[kMaxRing];           // scalar error estimate of the state in every slot
    T* ref_ring[kMaxRing];           // reference_state vectors (null entries: not written)
    const T* halo_lo;
    const T* halo_hi;
    long long halo_parity_stride;
    Ek0Mid<T, N>* mid;
    unsigned long long* mid_flag;    // raised to mid_seq + attempt once the global sum of z^2 is published
    unsigned long long mid_seq;      // of the kernel's first attempt
    unsigned long long* cl_flag;     // raised to cl_seq + attempt + 1 once the attempt's new factor has been written
    unsigned long long cl_seq;
    double* partials;                // [2][gridDim.x] (unused by the unit-partial reductions; timing instrumentation)
    unsigned int* ticket;            // [2]
    double* z_sq_out;                // global sum of z^2 (diagnostic / API parity with the three-phase path)
    double* out_sum;                 // sum_i (dt err / tol_i)^2
    RedPlan red_a, red_b;            // unit partials of the two phases
    double d_global;
    unsigned int col_blocks, chunks, total_strips;
    int block0, nblocks, row_block;  // fhn: first global row block that touches the shard, number of row blocks of the shard, R
    int use_bulk;
    float keep_frac;                 // share of a strip's rows kept in L2 across the phase turn-arounds (>= 1: no hints)
    CommWait wait;
    CommReduce reduce_a, reduce_b;   // seq of the kernel's first attempt; an attempt uses seq + 2 * attempt
    PackPeer<T> pack;
    // device-resident step controller (kCtl instantiation)
    StepCtl* ctl;
    int max_attempts;
    long long max_accepted;          // stop the launch after this many newly accepted steps (0: no limit)
};

struct StripGeom {
    int c0, wlen, r0, r1, dir;
    int b0;                  // local index of the strip's first row block
};

// The error-norm and z^T z sums of the fhn kernel are formed per ROW BLOCK of RB grid rows (aligned in GLOBAL row numbers)
// and 256-column strip.  RB = 8 (template parameter) for large grids: every thread adds its column's terms of the block's
// rows in walking order -- up in phase A, down in phase B, the same for every strip -- then the lanes go through the shuffle
// tree and the warps through shared memory, once per 8 rows.  A row the grid / shard does not have is simply absent.  Strips
// are whole numbers of row blocks, so a block's partial is a fixed function of the block's coordinates: independent of how
// many strips, CTAs or GPUs the grid is cut into.  Small grids take RB = 1 (one shuffle tree per row) so that there are still
// enough strips to fill the device (ek0_row_block: a function of the GLOBAL grid only, i.e. the same for every number of ranks).
__host__ __device__ inline int ek0_row_block(long long global_rows, int col_blocks) {
    return ((global_rows + 7) / 8) * col_blocks >= 256 ? 8 : 1;
}

// calibration, the single stacked QR, gain, factor update (one thread, fp64)   ek0.py:123-140,175,182
template <typename T, int N>
__device__ __noinline__ void ek0_mid_compute(const Consts<double, N>& c, const T* Cl_in, double z_sq, double d_global,
                                             T* Cl_out, T* err_out, Ek0Mid<T, N>* mid) {
    const double p1 = c.p[1];
    const double Q11 = QL(c, 1, 0) * QL(c, 1, 0) + QL(c, 1, 1) * QL(c, 1, 1);
    const double HQH = p1 * p1 * Q11;
    const double sigma_sq = z_sq / HQH / d_global;
    const double err = sqrt(sigma_sq * HQH);
    const double sigma = sqrt(sigma_sq);
    double B[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double col[N];
#pragma unroll
        for (int k = 0; k < N; ++k) col[k] = c.pinv[k] * (double)__ldcg(Cl_in + k * N + j);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = col[i];
#pragma unroll
            for (int k = i + 1; k < N; ++k) acc = fma(A_entry<N>(i, k), col[k], acc);
            B[i][j] = acc;
        }
    }
    stacked_qr<double, N>(B, sigma, c);
    const double g0 = p1 * B[1][0], g1 = p1 * B[1][1];   // H Clp = p1 * Clp[1, :]
    const double S = p1 * p1 * (B[1][0] * B[1][0] + B[1][1] * B[1][1]);
#pragma unroll
    for (int r = 0; r < N; ++r) {
        const double l0 = B[r][0], l1 = (r >= 1) ? B[r][1] : 0.0;
        const double kg = (l0 * g0 + l1 * g1) / S;
        mid->K[r] = (T)kg;
        const double pr = c.p[r];
        Cl_out[r * N + 0] = (T)(pr * (l0 - kg * g0));
        Cl_out[r * N + 1] = (T)(pr * (l1 - kg * g1));
#pragma unroll
        for (int cc = 2; cc < N; ++cc) Cl_out[r * N + cc] = (r >= cc) ? (T)(pr * B[r][cc]) : (T)0;
    }
    *err_out = (T)err;
    mid->err = err;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int W>
__device__ __forceinline__ StripGeom strip_geom(const Geo& g, unsigned strip, unsigned col_blocks, unsigned chunks, int block0,
                                                int nblocks, int kRowBlock, bool has_lo) {
    StripGeom s;
    const unsigned chunk = strip / col_blocks, cb = strip - chunk * col_blocks;
    s.c0 = (int)cb * W;
    s.wlen = min(W, g.cols - s.c0);
    // whole row blocks (aligned in global rows), clipped to the shard
    const int bl = (int)(((long long)chunk * nblocks) / chunks), bh = (int)(((long long)(chunk + 1) * nblocks) / chunks);
    s.b0 = bl;
    s.r0 = max(0, (int)((long long)(block0 + bl) * kRowBlock - g.grow0));
    s.r1 = min(g.rows, (int)((long long)(block0 + bh) * kRowBlock - g.grow0));
    // Strips walk UP in phase A and DOWN in phase B (the rows phase A read last are the rows phase B reads first: L2 reuse
    // across the turn-around), and that is also the order in which a thread adds the rows of a row block -- a function of the
    // phase only, not of which strip, CTA or GPU the block landed in.  ONE exception to the walk, not to the order of the sums:
    // the strips under the shard's lower edge walk DOWN in phase A so that the neighbour GPU's halo row is needed last, not
    // first (its delivery takes ~10 us from the start of the attempt); they buffer a block's terms and add them in the usual order.
    s.dir = (has_lo && chunk == 0u && chunks > 1u) ? -1 : 1;
    return s;
}

// sigma-independent part of column 1 of the predicted covariance: M1 = (B B^T)[:, 1], B = A (p_inv * Cl).
// The gain K = Clp Clp^T H^T / S (ek0.py:136) only needs column 1 of Clp Clp^T = B B^T + sigma^2 Ql Ql^T, so once
// sigma^2 is known (after phase A) every CTA derives K in a handful of flops; the stacked QR that produces the new
// factor itself (ek0.py:175,138) is off the critical path (service warp of CTA 0, while phase B is streaming).
template <typename T, int N>
__device__ __forceinline__ void ek0_m1_column(const Consts<double, N>& cd, const T* Cl_in, double* m1) {
    double B[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double colv[N];
#pragma unroll
        for (int kd = 0; kd < N; ++kd) colv[kd] = cd.pinv[kd] * (double)__ldcg(Cl_in + kd * N + j);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double accb = colv[i];
#pragma unroll
            for (int kd = i + 1; kd < N; ++kd) accb = fma(A_entry<N>(i, kd), colv[kd], accb);
            B[i][j] = accb;
        }
    }
#pragma unroll
    for (int r = 0; r < N; ++r) {
        double m = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) m = fma(B[r][j], B[1][j], m);
        m1[r] = m;
    }
}

// calibration (ek0.py:123-130) and gain (ek0.py:132-136) from the global sum of z^2; returns the calibrated error estimate
template <typename T, int N>
__device__ __forceinline__ double ek0_gain(const Consts<double, N>& cd, const double* m1, double z_sq, double d_global, T (&K)[N]) {
    const double p1 = cd.p[1];
    const double Q11 = QL(cd, 1, 0) * QL(cd, 1, 0) + QL(cd, 1, 1) * QL(cd, 1, 1);
    const double HQH = p1 * p1 * Q11;
    const double sigma_sq = z_sq / HQH / d_global;
    const double den = p1 * fma(sigma_sq, Q11, m1[1]);
#pragma unroll
    for (int r = 0; r < N; ++r) {
        // (Ql Ql^T)[r][1]: rows 0 and 1 of Ql have 1 and 2 nonzeros
        const double q = (r == 0) ? QL(cd, 0, 0) * QL(cd, 1, 0) : QL(cd, r, 0) * QL(cd, 1, 0) + QL(cd, r, 1) * QL(cd, 1, 1);
        K[r] = (T)(fma(sigma_sq, q, m1[r]) / den);
    }
    return sqrt(sigma_sq * HQH);
}

// ---------------------------------------------------------------------------------------------------------------------
// Pieces shared by the two fused kernels (fhn strips / 1-d tile ranges).  Thread roles: consumers [0, CONS), the producer
// warp [CONS, CONS + 32).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int N>
struct Ek0CtaState {
    T p[N], pinv[N], dt;             // kCtl: the step's vectors rebuilt from the controller block (see StepView)
    Consts<double, N> cd;
    int cur, out;                    // ring slots of the attempt's input / output state (cur < 0: stop)
    int flag, ok;
    unsigned long long epoch0;
    long long pause_at;
    uint64_t m1_bar;                 // the producer warp has published m1[attempt & 1]
    double m1[2][kMaxN + 1];
    T gain[kMaxN];                   // the attempt's gain K (ek0.py:136): phase B reads it from here (broadcast), not from registers
    double timed_out;                // the wait for the global sum of z^2 timed out
};

// top of an attempt, all threads of the CTA: thread 0 waits for the previous attempt's decision and reads the controller
template <typename T, int N, bool kCtl>
__device__ __forceinline__ bool ek0_attempt_begin(const Ek0FusedArgs<T, N>& a, Ek0CtaState<T, N>& sh, int att) {
    if (threadIdx.x == 0) {
        if constexpr (kCtl) {
            StepCtl* ctl = a.ctl;
            bool ok = true;
            if (att == 0) {
                sh.epoch0 = ld_acquire_gpu(&ctl->epoch);
                sh.pause_at = a.max_accepted > 0 ? ctl->accepted + a.max_accepted : (long long)0x7fffffffffffffffLL;
            } else {
                ok = spin_until_gpu(&ctl->epoch, sh.epoch0 + (unsigned long long)att, (unsigned long long)ctl->wait_timeout_ns);
            }
            if (!ok) {
                ctl->failed = TDX_FAIL_GRID_BARRIER;
                ctl->done = 1;
            }
            Consts<double, N> ccd = a.cd;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                ccd.p[k] = ctl->p[k];
                ccd.pinv[k] = ctl->pinv[k];
                sh.p[k] = (T)ctl->p[k];
                sh.pinv[k] = (T)ctl->pinv[k];
            }
            ccd.dt = ctl->dt;
            sh.dt = (T)ctl->dt;
            sh.cd = ccd;
            const int cur = ctl->cur, nbuf = ctl->nbuf;
            sh.cur = (ctl->done || !ok || ctl->accepted >= sh.pause_at) ? -1 : cur;
            sh.out = (cur + 1) % nbuf;
        } else {
            sh.cur = 0;
            sh.out = 1;
        }
    }
    __syncthreads();
    return sh.cur >= 0;
}

// The n x n algebra of an attempt is done by lane 0 of the PRODUCER warp while it has nothing to issue:
//  * after the phase A loads are on their way: M1 of the attempt's input factor (every CTA; the consumers need it for the gain
//    right after the barrier in the middle of the attempt);
//  * after the phase B loads are on their way, in CTA 0 only: the ONE stacked QR for the attempt's output factor
//    (ek0.py:175,138), as soon as the global sum of z^2 is known.  It overlaps the consumers' last rows and the end barrier.
template <typename T, int N>
__device__ __forceinline__ void ek0_service_m1(const Ek0FusedArgs<T, N>& a, Ek0CtaState<T, N>& sh, const Consts<double, N>& cd,
                                               int att, int lane) {
    if (lane == 0) {
        // the factor this attempt starts from was written by the previous attempt's QR (the first attempt's by an earlier launch)
        const bool ok = att == 0 || spin_until_gpu(a.cl_flag, a.cl_seq + (unsigned long long)att, a.wait.timeout_ns);
        ek0_m1_column<T, N>(cd, a.Cl_ring[sh.cur], sh.m1[att & 1]);
        if (!ok) sh.m1[att & 1][1] = nan("");
        mbar_arrive(&sh.m1_bar);
    }
    __syncwarp();
}

template <typename T, int N>
__device__ __forceinline__ void ek0_service_qr(const Ek0FusedArgs<T, N>& a, Ek0CtaState<T, N>& sh, const Consts<double, N>& cd,
                                               int att, int lane) {
    if (lane == 0 && blockIdx.x == 0) {
        const bool ok = spin_until_gpu(a.mid_flag, a.mid_seq + (unsigned long long)att, a.wait.timeout_ns);
        const double z_sq = ok ? __ldcg(&a.mid->z_sq) : nan("");
        TDX_STAMP(a.partials, att, 5);
        ek0_mid_compute<T, N>(cd, a.Cl_ring[sh.cur], z_sq, a.d_global, a.Cl_ring[sh.out], a.err_ring[sh.out], a.mid);
        TDX_STAMP(a.partials, att, 6);
        __threadfence();
        st_release_gpu(a.cl_flag, a.cl_seq + (unsigned long long)att + 1ull);
    }
    __syncwarp();
}

// consumers, after phase A: publish the CTA's arrival; the last CTA reduces the unit partials (over the ranks too) and
// releases the global sum of z^2; everybody waits for it and derives the gain.  Returns the calibrated error estimate.
template <typename T, int N, class CSync>
__device__ __forceinline__ double ek0_mid_barrier(const Ek0FusedArgs<T, N>& a, Ek0CtaState<T, N>& sh, const Consts<double, N>& cd,
                                                  int att, double* s_red, T (&K)[N], CSync csync) {
    const int tc = threadIdx.x;
    if (cta_arrive_last(a.ticket, &sh.flag, csync)) {
        const double z_sq = reduce_units(a.red_a, &a.reduce_a, a.reduce_a.seq + 2ull * (unsigned long long)att,
                                         s_red + kEk0RedScratch, csync);
        if (tc == 0) {
            a.mid->z_sq = z_sq;
            if (a.z_sq_out) *a.z_sq_out = z_sq;
            __threadfence();
            st_release_gpu(a.mid_flag, a.mid_seq + (unsigned long long)att);
        }
    }
    if (tc == 0) {
        sh.timed_out = spin_until_gpu(a.mid_flag, a.mid_seq + (unsigned long long)att, a.wait.timeout_ns) ? 0.0 : 1.0;
#ifdef TDX_DEBUG_WAITS
        if (sh.timed_out != 0.0) a.partials[3004] = (double)(1 + blockIdx.x) + 1e-3 * att;
#endif
        TDX_MBAR_WAIT(&sh.m1_bar, (unsigned)att & 1u, a.partials, 3, att);           // the producer warp's m1 of this attempt
    }
    csync();
    // a dead peer must not hang the GPU and must not go unnoticed: NaN state, NaN norm, the controller / host raises
    const double z_sq_all = (sh.timed_out != 0.0) ? nan("") : __ldcg(&a.mid->z_sq);
    return ek0_gain<T, N>(cd, sh.m1[att & 1], z_sq_all, a.d_global, K);
}

// consumers, after phase B: the last CTA reduces sum_i 1 / tol_i^2, scales it to sum_i (dt err / tol_i)^2 (step.py:101-104 with a
// scalar error estimate), writes it out and -- under the controller -- decides and publishes the decision
template <typename T, int N, bool kCtl, class CSync, class C>
__device__ __forceinline__ void ek0_end_barrier(const Ek0FusedArgs<T, N>& a, Ek0CtaState<T, N>& sh, const C& c, int att,
                                                double err_cal, double* s_red, CSync csync) {
    const int tc = threadIdx.x;
    if (!(c.want_norm || kCtl)) return;
    if (cta_arrive_last(a.ticket + 1, &sh.flag, csync)) {
        double b = 0.0;
        if (c.want_norm || a.reduce_b.world > 0) {
            b = reduce_units(a.red_b, &a.reduce_b, a.reduce_b.seq + 2ull * (unsigned long long)att, s_red + kEk0RedScratch, csync);
            if (tc == 0) {
                b *= err_cal * err_cal;
                b *= (double)c.dt * (double)c.dt;
                if (c.want_norm && a.out_sum) *a.out_sum = b;
            }
        }
        if constexpr (kCtl) {
            if (tc == 0) ctl_finish_attempt(a.ctl, b, a.d_global, sh.epoch0 + (unsigned long long)att + 1ull);
        }
    }
}

// Threads [0, W): consumers, thread t owns column c0 + t of both fields.  Threads [W, W + 32): the producer warp.
template <typename T, int N, int W, bool kCtl = false, int RB = 8>
__global__ void __maxnreg__(TDX_EK0_MAXREG) ek0_fused_fhn_kernel(const __grid_constant__ Ek0FusedArgs<T, N> a) {
    static_assert(RB == 1 || RB == 8, "row blocks of 1 or 8 grid rows");
    using Cfg = Ek0FusedCfg<T, N, W>;
    using CSync = NamedSync<1, W>;                         // the consumers' barrier (the producer warp never joins)
    constexpr int PADC = Cfg::PADC, WP = Cfg::WP, S = Cfg::STAGES, AT_ROW = Cfg::AT_ROW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                                   // [S] data has landed
    uint64_t* empty = full + S;                                                               // [S] consumers are done
    double* s_red = reinterpret_cast<double*>(smem_raw + Cfg::BAR_BYTES);                     // see kEk0Red*
    double* s_buf = reinterpret_cast<double*>(smem_raw + Cfg::BAR_BYTES + Cfg::RED_BYTES);                    // [8][W]
    T* s_at = reinterpret_cast<T*>(smem_raw + Cfg::BAR_BYTES + Cfg::RED_BYTES + Cfg::BUF_BYTES);               // [2][2][AT_ROW]
    T* s_stage = reinterpret_cast<T*>(smem_raw + Cfg::BAR_BYTES + Cfg::RED_BYTES + Cfg::BUF_BYTES + Cfg::at_bytes());
    __shared__ Ek0CtaState<T, N> sh;

    const Geo& g = a.g;
    const int tc = threadIdx.x;
    const bool has_lo = a.halo_lo != nullptr, has_hi = a.halo_hi != nullptr;
    const bool hints = a.keep_frac < 1.0f;
    const int ns = (int)((a.total_strips - blockIdx.x + gridDim.x - 1) / gridDim.x);   // strips of this CTA

    if (tc == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(&sh.m1_bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    unsigned q = 0;                       // producer: ring position
    unsigned cq = 0, cr = 0, phase_bits = 0u;     // consumers: row counter, ring position, parity of each stage's full barrier
    const int max_attempts = kCtl ? a.max_attempts : 1;

#pragma unroll 1
    for (int att = 0; att < max_attempts; ++att) {
        if (!ek0_attempt_begin<T, N, kCtl>(a, sh, att)) break;
        decltype(auto) c = [&]() -> decltype(auto) {
            if constexpr (kCtl) return StepView<T, N>{sh.p, sh.pinv, a.c.ql, sh.dt, a.c.abstol, a.c.reltol, a.c.want_norm, a.c.flags};
            else return (a.c);
        }();
        const Consts<double, N>& cdr = kCtl ? sh.cd : a.cd;
        const T* mean_in = a.mean_ring[sh.cur];
        T* mean_out = a.mean_ring[sh.out];
        T* ref_out = a.ref_ring[sh.out];
        const unsigned long long hseq = a.wait.seq + (unsigned long long)att;
        const T* halo_lo = has_lo ? a.halo_lo + (long long)(hseq & 1ull) * a.halo_parity_stride : nullptr;
        const T* halo_hi = has_hi ? a.halo_hi + (long long)(hseq & 1ull) * a.halo_parity_stride : nullptr;

        if (tc >= W) {
            // ---- producer warp: walks the same (strip, row) enumeration as the consumers, up to S rows ahead ------------
            // Every row inside the shard owns ring slot q % S.  The 2 n row segments of a
            // row are issued by 2 n lanes in one instruction.
            const int lane = tc - W;
            const int seg_f = lane / N, seg_k = lane - seg_f * N;           // (field, derivative) of this lane's segment
            const long long seg_src = (long long)seg_k * g.d + (long long)seg_f * g.npts;
            const int seg_dst = (seg_f * N + seg_k) * WP;
            const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
            if (kCtl) fence_proxy_async_global();      // the mean was written with ordinary stores by the previous attempt
#pragma unroll 1
            for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
                for (int kk = 0; kk < ns; ++kk) {
                    const int k = phase ? ns - 1 - kk : kk;
                    StripGeom s = strip_geom<W>(g, blockIdx.x + (unsigned)k * gridDim.x, a.col_blocks, a.chunks, a.block0, a.nblocks, RB, has_lo);
                    if (phase) s.dir = -s.dir;
                    const int L = s.r1 - s.r0 + 2;
                    const int gs = max(s.c0 - PADC, 0), ge = min(s.c0 + W + PADC, g.cols);
                    const int off = gs - (s.c0 - PADC);
                    const uint32_t bytes = (uint32_t)((ge - gs) * (int)sizeof(T));
                    // L2 residency: the rows phase A reads last are the rows phase B reads first -> keep those, stream the rest
                    const int keep_from = (phase == 0 && hints) ? L - 1 - (int)ceilf(a.keep_frac * (float)(s.r1 - s.r0)) : L;
                    int rr = s.dir > 0 ? s.r0 - 1 : s.r1;
#pragma unroll 1
                    for (int i = 0; i < L; ++i, rr += s.dir) {
                        // Rows beyond the shard take no ring slot: a consumer can get through such a row without waiting for
                        // anything, i.e. ahead of the producer, and two completions of an "empty" barrier in a row would make the
                        // producer's parity wait miss its turn.
                        if (rr < 0 || rr >= g.rows) continue;
                        const unsigned stage = q % S, round = q / S;
                        ++q;
                        if (round > 0) TDX_MBAR_WAIT(empty + stage, (round - 1u) & 1u, a.partials, 2, att);     // the slot's previous row is consumed
                        T* dst = s_stage + (size_t)stage * Cfg::STAGE_ELEMS;
                        const T* src = mean_in + (long long)rr * g.cols + gs;
                        if (a.use_bulk) {
                            if (lane == 0) mbar_expect_tx(full + stage, 2u * N * bytes);
                            __syncwarp();
                            if (lane < 2 * N) {
                                if (hints) bulk_g2s_hint(dst + seg_dst + off, src + seg_src, bytes, full + stage, i >= keep_from ? pol_keep : pol_stream);
                                else bulk_g2s(dst + seg_dst + off, src + seg_src, bytes, full + stage);
                            }
                        } else {
                            // unaligned / odd-sized grids: plain loads by the 32 lanes, then a normal arrival
                            for (int seg = 0; seg < 2 * N; ++seg) {
                                const int f = seg / N, kd = seg - f * N;
                                const T* sp = src + (long long)kd * g.d + (long long)f * g.npts;
                                for (int j = lane; j < ge - gs; j += 32) dst[seg * WP + off + j] = ld_state<kCtl>(sp + j);
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(full + stage);
                        }
                    }
                }
                if (phase == 0) ek0_service_m1<T, N>(a, sh, cdr, att, lane);
            }
            ek0_service_qr<T, N>(a, sh, cdr, att, lane);
            continue;
        }

        // ---- consumers ------------------------------------------------------------------------------------------------
        CSync csync;
        double* s_rowred = s_red + kEk0RowRed;
        // deliver this shard's edge rows to the neighbours before anything else (multi-GPU)
        if (a.pack.enabled)
            for (int part = (int)blockIdx.x; part < 2 * kPackCtas; part += (int)gridDim.x)   // a small grid takes several parts per CTA
                pack_halo_part<T, N, CSync, kCtl>(g, c, mean_in, a.pack, a.pack.seq + (unsigned long long)att, part, csync);

        bool halos_ready = false, poisoned = false;
        int pending_unit = -1;                       // CTA-uniform: the row block whose warp sums wait in s_rowred for the next barrier
        int pending_slot = 0;
        const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
        const T inv_dx2 = (T)g.prm[0], pa = (T)g.prm[1], pb = (T)g.prm[2], pk = (T)g.prm[3], inv_tau = (T)g.prm[5];
        const int wq = tc >> 5, lq = tc & 31;

        auto m_at_from = [&](const T* fld, int idx) -> T {     // predicted, un-preconditioned state of a stage column
            T acc = c.pinv[0] * fld[idx];
#pragma unroll
            for (int k = 1; k < N; ++k) acc = fma((T)A_entry<N>(0, k), c.pinv[k] * fld[k * WP + idx], acc);
            return c.p[0] * acc;
        };

        // one phase over this CTA's strips.  kUpdate = false: phase A (forward), true: phase B (backward).
        // Every finished row block (RB = 8: eight grid rows; RB = 1: every row) leaves ONE partial sum per column block: the
        // threads' block sums through the shuffle tree, the warps through shared memory in warp order (see the note on row blocks
        // above ek0_row_block) -- independent of strip boundaries, CTAs and GPUs.
        auto run_phase = [&](auto update_tag, const T (&K)[N], double* unit) {
            constexpr bool kUpdate = decltype(update_tag)::value;
            constexpr int NK = kUpdate ? N : 2;                 // rows of the predicted mean kept for the row being finalised
            const bool reduces = !kUpdate || c.want_norm || (kCtl && a.reduce_b.world > 0);
            // after a barrier that follows a block's warp sums: the block's partial = the warp sums in warp order
            auto flush_block_partial = [&]() {
                if (pending_unit >= 0) {
                    if (tc == 0) {
                        const double* ws = s_rowred + (pending_slot & 1) * 8;
                        double b = 0.0;
#pragma unroll
                        for (int i = 0; i < W / 32; ++i) b += ws[i];
                        unit[pending_unit] = b;
                    }
                    pending_unit = -1;
                }
            };
            constexpr int kRowBlock = RB;
            double t0 = 0.0;                                    // RB = 8: this thread's sum over the rows of the block so far
            int blk = 0;                                        // local index of the row block being summed
            // a finished block: lanes through the shuffle tree, warps through shared memory; the warp sums of the block before
            // (stored at least one row, i.e. one barrier, ago) are added up first
            auto emit_block = [&](double v, int cb_, int dir_) {
                flush_block_partial();
                const double ws = warp_sum(v);
                pending_slot ^= 1;
                if (lq == 0) s_rowred[(pending_slot & 1) * 8 + wq] = ws;
                pending_unit = blk * (int)a.col_blocks + cb_;
                blk += dir_;
            };
#pragma unroll 1
            for (int kk = 0; kk < ns; ++kk) {
                const int k = kUpdate ? ns - 1 - kk : kk;
                const unsigned strip = blockIdx.x + (unsigned)k * gridDim.x;
                StripGeom s = strip_geom<W>(g, strip, a.col_blocks, a.chunks, a.block0, a.nblocks, RB, has_lo);
                const int cb = (int)(strip % a.col_blocks);
                if (kUpdate) s.dir = -s.dir;
                const int L = s.r1 - s.r0 + 2;
                const int col = s.c0 + tc;
                const bool valid = tc < s.wlen;
                const bool left_inner = (tc == 0) && s.c0 > 0;                            // recompute the column left of the strip
                const bool right_inner = (tc == s.wlen - 1) && (s.c0 + s.wlen < g.cols);  // ... right of the strip
                const bool left_edge = (tc == 0) && s.c0 == 0;                            // grid edge: replicate own value
                const bool right_edge = (tc == s.wlen - 1) && (s.c0 + s.wlen >= g.cols);
                int rr = s.dir > 0 ? s.r0 - 1 : s.r1;
                T at_m1[2] = {}, at_m2[2] = {}, hs_m1[2] = {}, mp_m1[2][NK] = {};
                bool first_replicated = false;
                T* out_ptr = mean_out + ((long long)(rr - s.dir) * g.cols + col);      // row being finalised, field 0
                // the row block of the strip's first finalised row, and that row's leaf number inside it (in walking order)
                int leaf = 0;
                {
                    const long long first = g.grow0 + (s.dir > 0 ? s.r0 : s.r1 - 1);
                    blk = (int)(first / kRowBlock) - a.block0;
                    if (RB == 8) {
                        const int pos = (int)(first & 7);
                        leaf = s.dir > 0 ? pos : 7 - pos;
                        t0 = 0.0;
                    }
                }
                const bool buffered = RB == 8 && reduces && s.dir != (kUpdate ? -1 : 1);
                if (buffered) {
#pragma unroll
                    for (int j8 = 0; j8 < 8; ++j8) s_buf[j8 * W + tc] = 0.0;
                }
                // the rows phase B writes last are the rows the NEXT attempt's phase A reads first -> keep those in L2
                const int keep_from = (kUpdate && hints) ? L - (int)ceilf(a.keep_frac * (float)(s.r1 - s.r0)) : L;
#pragma unroll 1
                for (int i = 0; i < L; ++i, rr += s.dir, ++cq, out_ptr += (long long)s.dir * g.cols) {
                    const unsigned stage = cr % S;
                    const bool in_shard = rr >= 0 && rr < g.rows;
                    T* at_row = s_at + (cq & 1u) * 2 * AT_ROW;
                    T at_cur[2], mp_cur[2][NK];
                    if (in_shard) {
                        ++cr;
                        TDX_MBAR_WAIT(full + stage, (phase_bits >> stage) & 1u, a.partials, 1, att);
                        phase_bits ^= (1u << stage);
                        const T* st = s_stage + (size_t)stage * Cfg::STAGE_ELEMS;
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            T raw[N], mp[N];
#pragma unroll
                            for (int kd = 0; kd < N; ++kd) raw[kd] = st[(f * N + kd) * WP + PADC + tc];
                            predict_column<T, N>(raw, c, mp);
                            at_cur[f] = c.p[0] * mp[0];
#pragma unroll
                            for (int r = 0; r < NK; ++r) mp_cur[f][r] = mp[r];
                        }
                        // columns just outside the strip (inside the grid): recomputed from the stage's overlap columns
                        if (left_inner) {
#pragma unroll
                            for (int f = 0; f < 2; ++f) at_row[f * AT_ROW] = m_at_from(st + f * N * WP, PADC - 1);
                        }
                        if (right_inner) {
#pragma unroll
                            for (int f = 0; f < 2; ++f) at_row[f * AT_ROW + s.wlen + 1] = m_at_from(st + f * N * WP, PADC + s.wlen);
                        }
                    } else {
                        const T* h = rr < 0 ? halo_lo : halo_hi;
                        if (h != nullptr) {
                            if (!halos_ready) {
                                if (!cta_wait_halos(a.wait, hseq, &sh.ok, csync)) poisoned = true;
                                halos_ready = true;
                            }
#pragma unroll
                            for (int f = 0; f < 2; ++f) at_cur[f] = valid ? __ldcg(h + (long long)f * g.cols + col) : (T)0;
                        } else {
                            // edge replication (ivp.py:488-490): the row beyond the grid equals the grid's edge row
#pragma unroll
                            for (int f = 0; f < 2; ++f) at_cur[f] = at_m1[f];
                            if (i == 0) first_replicated = true;
                        }
#pragma unroll
                        for (int f = 0; f < 2; ++f)
#pragma unroll
                            for (int r = 0; r < NK; ++r) mp_cur[f][r] = (T)0;
                    }
                    if (i == 1 && first_replicated) {
                        at_m1[0] = at_cur[0];
                        at_m1[1] = at_cur[1];
                    }
                    if (valid) {
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            at_row[f * AT_ROW + tc + 1] = at_cur[f];
                            if (left_edge) at_row[f * AT_ROW] = at_cur[f];
                            if (right_edge) at_row[f * AT_ROW + s.wlen + 1] = at_cur[f];
                        }
                    }
                    csync();              // publishes the row; every consumer has read the stage -> hand it back
                    if (tc == 0 && in_shard) mbar_arrive(empty + stage);
                    T hs_cur[2];
#pragma unroll
                    for (int f = 0; f < 2; ++f) hs_cur[f] = at_row[f * AT_ROW + tc] + at_row[f * AT_ROW + tc + 2];

                    if (i >= 2) {         // the previous row now has both vertical neighbours
                        double term = 0.0;
                        if (valid) {
                            const T u = at_m1[0], v = at_m1[1];
#pragma unroll
                            for (int f = 0; f < 2; ++f) {
                                const T own = at_m1[f];
                                const T lap = (((at_m2[f] + at_cur[f]) + hs_m1[f]) - (T)4 * own) * inv_dx2;
                                T fx;
                                if (f == 0) fx = pa * lap + u - u * u * u - v + pk;       // ivp.py:441
                                else fx = (pb * lap + u - v) * inv_tau;                   // ivp.py:442
                                const T z = c.p[1] * mp_m1[f][1] - fx;
                                if constexpr (!kUpdate) {
                                    term += (double)z * (double)z;
                                } else {
                                    // all n values first, then the stores: the stores are opaque to the compiler (asm with a
                                    // memory clobber), and loads of p / K between them would be serialised behind each one
                                    T mo[N];
#pragma unroll
                                    for (int r = 0; r < N; ++r) mo[r] = c.p[r] * (mp_m1[f][r] - sh.gain[r] * z);
                                    T* mo_ptr = out_ptr + (long long)f * g.npts;
                                    const T ref = tabs<T>(mo[0]);
                                    T* rp = ref_out ? ref_out + ((out_ptr - mean_out) + (long long)f * g.npts) : nullptr;
                                    if (hints) {
                                        const uint64_t pol_row = i >= keep_from ? pol_keep : pol_stream;
#pragma unroll
                                        for (int r = 0; r < N; ++r) st_global_hint(mo_ptr + (long long)r * g.d, mo[r], pol_row);
                                        if (rp) st_global_hint(rp, ref, pol_stream);
                                    } else {
#pragma unroll
                                        for (int r = 0; r < N; ++r) mo_ptr[(long long)r * g.d] = mo[r];
                                        if (rp) *rp = ref;
                                    }
                                    if (c.want_norm) {
                                        const double tol = (double)c.abstol + (double)c.reltol * (double)ref;
                                        term += 1.0 / (tol * tol);
                                    }
                                }
                            }
                        }
#ifndef TDX_EK0_NO_SUMS
                        if (reduces) {
                            if (RB == 1) {
                                emit_block(term, cb, s.dir);
                            } else {
                                if (!buffered) {
                                    t0 += term;               // rows of a block in walking order: up in phase A, down in phase B
                                    leaf = (leaf + 1) & 7;
                                    if (leaf == 0) {
                                        emit_block(t0, cb, s.dir);
                                        t0 = 0.0;
                                    }
                                } else {
                                    // this strip walks against the summation order: park the term, add at the end of the block
                                    s_buf[(7 - leaf) * W + tc] = term;
                                    leaf = (leaf + 1) & 7;
                                    if (leaf == 0) {
                                        double x = 0.0;
#pragma unroll
                                        for (int j8 = 0; j8 < 8; ++j8) {
                                            x += s_buf[j8 * W + tc];
                                            s_buf[j8 * W + tc] = 0.0;
                                        }
                                        emit_block(x, cb, s.dir);
                                    }
                                }
                            }
                        }
#endif
                    }
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        at_m2[f] = at_m1[f];
                        at_m1[f] = at_cur[f];
                        hs_m1[f] = hs_cur[f];
#pragma unroll
                        for (int r = 0; r < NK; ++r) mp_m1[f][r] = mp_cur[f][r];
                    }
                }
                if (RB == 8 && reduces && leaf != 0) {
                    // the strip ends inside a row block (the grid / the shard ends there): the block's partial is the sum of the
                    // rows there are.  The previous block was emitted at least one row (one barrier) ago.
                    double x = t0;
                    if (buffered) {
                        x = 0.0;
#pragma unroll
                        for (int j8 = 0; j8 < 8; ++j8) x += s_buf[j8 * W + tc];      // rows the block does not have are +0.0
                    }
                    emit_block(x, cb, s.dir);
                    t0 = 0.0;
                    leaf = 0;
                }
            }
            if (reduces) {
                csync();
                flush_block_partial();
                csync();
                if (tc == 0 && poisoned && ns > 0) {     // stale halos: fail the attempt everywhere
                    const StripGeom s0 = strip_geom<W>(g, blockIdx.x, a.col_blocks, a.chunks, a.block0, a.nblocks, RB, has_lo);
                    unit[s0.b0 * (int)a.col_blocks + (int)(blockIdx.x % a.col_blocks)] = nan("");
                }
            }
        };

        T K[N];
#pragma unroll
        for (int r = 0; r < N; ++r) K[r] = (T)0;

#ifdef TDX_TIMELINE
        // development aid (scripts/ek0_cta_spread.py): per CTA [t_start, t_end_A, t_gain_seen, t_end_B, t_fenced] in ns, and its SM
        double* stamps = a.partials + 4096 + 8 * blockIdx.x;
        if (tc == 0) {
            stamps[0] = (double)global_timer_ns();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            stamps[5] = (double)smid;
        }
#endif
        // ---- phase A ------------------------------------------------------------------------------------------------
        if (tc == 0) TDX_STAMP(a.partials, att, 0);
        run_phase(std::false_type{}, K, a.red_a.unit);
#ifdef TDX_TIMELINE
        if (tc == 0) stamps[1] = (double)global_timer_ns();
#endif
        if (tc == 0) TDX_STAMP(a.partials, att, 1);
        const double err_cal = ek0_mid_barrier<T, N>(a, sh, cdr, att, s_red, K, csync);
        if (tc == 0) {
#pragma unroll
            for (int r = 0; r < N; ++r) sh.gain[r] = K[r];
        }
        csync();
        if (tc == 0) TDX_STAMP(a.partials, att, 2);
#ifdef TDX_TIMELINE
        if (tc == 0) stamps[2] = (double)global_timer_ns();
#endif
        // ---- phase B ------------------------------------------------------------------------------------------------
        run_phase(std::true_type{}, K, a.red_b.unit);
#ifdef TDX_TIMELINE
        if (tc == 0) stamps[3] = (double)global_timer_ns();
#endif
        if (tc == 0) TDX_STAMP(a.partials, att, 3);
        if (kCtl) fence_proxy_async_global();          // the next attempt's TMA loads read what this thread just stored
#ifdef TDX_TIMELINE
        if (tc == 0) stamps[4] = (double)global_timer_ns();
#endif
        ek0_end_barrier<T, N, kCtl>(a, sh, c, att, err_cal, s_red, csync);
        if (tc == 0) TDX_STAMP(a.partials, att, 4);
    }
}

// =====================================================================================================================
// The same one-launch attempt for the 1-d vector fields (vanderpol, lorenz96, brusselator, burgers_1d, wave_1d, ...).
//
// Work unit = the 256-coordinate tile of the sweep kernels (tile_coord / ExtTile / IVP::ring_value / IVP::eval are reused as
// they are): thread = one state coordinate.  A CTA owns a CONTIGUOUS range of tiles, walks it forwards in phase A and
// backwards in phase B (L2 reuse), and gets every tile's raw mean columns through the same TMA ring + producer warp as the
// fhn kernel; the few stencil neighbours outside a tile (2-3 per field) are recomputed from the mean through L2.
// =====================================================================================================================
template <typename T, int N>
struct Ek0ChainCfg {
    static constexpr int STAGES = 4;
    static constexpr int THREADS = kThreads + 32;                     // consumers + producer warp (which also does the n x n algebra)
    static constexpr int STAGE_ELEMS = N * kThreads;                 // [derivative][thread]
    static constexpr size_t BAR_BYTES = 128, RED_BYTES = 512;
    static constexpr size_t TERM_BYTES = 2 * kThreads * sizeof(double);   // a tile's per-coordinate terms, double-buffered
    __host__ __device__ static constexpr size_t at_bytes(int at_elems) { return ((size_t)at_elems * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t total(int at_elems) {
        return BAR_BYTES + RED_BYTES + TERM_BYTES + 2 * at_bytes(at_elems) + STAGES * (size_t)STAGE_ELEMS * sizeof(T);
    }
};

template <class IVP, typename T, int N, bool kCtl = false>
__global__ void __maxnreg__(72) ek0_fused_chain_kernel(const __grid_constant__ Ek0FusedArgs<T, N> a) {
    using Shape = typename IVP::Shape;
    using E = ExtTile<IVP>;
    using Cfg = Ek0ChainCfg<T, N>;
    using CSync = NamedSync<1, kThreads>;
    constexpr int S = Cfg::STAGES, NF = Shape::NF, TCOLS = Shape::TC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + S;
    double* s_red = reinterpret_cast<double*>(smem_raw + Cfg::BAR_BYTES);
    double* s_term = reinterpret_cast<double*>(smem_raw + Cfg::BAR_BYTES + Cfg::RED_BYTES);    // [2][kThreads]
    T* s_at0 = reinterpret_cast<T*>(smem_raw + Cfg::BAR_BYTES + Cfg::RED_BYTES + Cfg::TERM_BYTES);
    constexpr size_t AT_STRIDE = Cfg::at_bytes(E::ELEMS) / sizeof(T);
    T* s_stage = reinterpret_cast<T*>(smem_raw + Cfg::BAR_BYTES + Cfg::RED_BYTES + Cfg::TERM_BYTES + 2 * Cfg::at_bytes(E::ELEMS));
    __shared__ Ek0CtaState<T, N> sh;

    const Geo& g = a.g;
    const int tid = threadIdx.x;

    // this CTA's contiguous range of tiles
    const unsigned num_tiles = a.total_strips, tiles_x = a.col_blocks;
    const unsigned t0 = (unsigned)(((unsigned long long)blockIdx.x * num_tiles) / gridDim.x);
    const unsigned t1 = (unsigned)(((unsigned long long)(blockIdx.x + 1) * num_tiles) / gridDim.x);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(&sh.m1_bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    unsigned q = 0, cq = 0, phase_bits = 0u;
    const int max_attempts = kCtl ? a.max_attempts : 1;

#pragma unroll 1
    for (int att = 0; att < max_attempts; ++att) {
        if (!ek0_attempt_begin<T, N, kCtl>(a, sh, att)) break;
        decltype(auto) c = [&]() -> decltype(auto) {
            if constexpr (kCtl) return StepView<T, N>{sh.p, sh.pinv, a.c.ql, sh.dt, a.c.abstol, a.c.reltol, a.c.want_norm, a.c.flags};
            else return (a.c);
        }();
        const Consts<double, N>& cdr = kCtl ? sh.cd : a.cd;
        const T* mean_in = a.mean_ring[sh.cur];
        T* mean_out = a.mean_ring[sh.out];
        T* ref_out = a.ref_ring[sh.out];
        const unsigned long long hseq = a.wait.seq + (unsigned long long)att;
        const T* halo_lo = a.halo_lo ? a.halo_lo + (long long)(hseq & 1ull) * a.halo_parity_stride : nullptr;
        const T* halo_hi = a.halo_hi ? a.halo_hi + (long long)(hseq & 1ull) * a.halo_parity_stride : nullptr;

        if (tid >= kThreads) {
            // ---- producer warp: one tile per ring slot; per tile NF * n contiguous row segments -------------------------
            const int lane = tid - kThreads;
            const int seg_f = lane / N, seg_k = lane - seg_f * N;
            if (kCtl) fence_proxy_async_global();
#pragma unroll 1
            for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
                for (unsigned j = 0; j < t1 - t0; ++j, ++q) {
                    const unsigned tile = phase ? (t1 - 1u - j) : (t0 + j);
                    const unsigned stage = q % S, round = q / S;
                    if (round > 0) mbar_wait(empty + stage, (round - 1u) & 1u);
                    // the tile's columns of field f: [col0, col0 + len) of a chain (one tile row)
                    const long long col0 = (long long)(tile % tiles_x) * TCOLS;
                    long long len = (long long)g.cols - col0;
                    len = len > TCOLS ? TCOLS : len;
                    T* dst = s_stage + (size_t)stage * Cfg::STAGE_ELEMS;
                    const uint32_t bytes = (uint32_t)(len * (long long)sizeof(T));
                    const bool bulk = a.use_bulk && (bytes % 16u == 0u) && (((col0 * (long long)sizeof(T)) & 15) == 0);
                    if (bulk) {
                        if (lane == 0) mbar_expect_tx(full + stage, (uint32_t)(NF * N) * bytes);
                        __syncwarp();
                        if (lane < NF * N)
                            bulk_g2s(dst + seg_k * kThreads + seg_f * TCOLS, mean_in + (long long)seg_k * g.d + (long long)seg_f * g.npts + col0,
                                     bytes, full + stage);
                    } else {
                        for (int seg = 0; seg < NF * N; ++seg) {
                            const int f = seg / N, kd = seg - f * N;
                            const T* sp = mean_in + (long long)kd * g.d + (long long)f * g.npts + col0;
                            for (int jj = lane; jj < (int)len; jj += 32) dst[kd * kThreads + f * TCOLS + jj] = ld_state<kCtl>(sp + jj);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full + stage);
                    }
                }
                if (phase == 0) ek0_service_m1<T, N>(a, sh, cdr, att, lane);
            }
            ek0_service_qr<T, N>(a, sh, cdr, att, lane);
            continue;
        }

        // ---- consumers ------------------------------------------------------------------------------------------------
        CSync csync;
        const int wq = tid >> 5, lq = tid & 31;
        if (a.pack.enabled)
            for (int part = (int)blockIdx.x; part < 2 * kPackCtas; part += (int)gridDim.x)
                pack_halo_part<T, N, CSync, kCtl>(g, c, mean_in, a.pack, a.pack.seq + (unsigned long long)att, part, csync);
        // the shard's end tiles read the neighbours' halo values: they are tiny and on their way since the peers' kernels started
        bool poisoned = false;
        if ((a.halo_lo != nullptr || a.halo_hi != nullptr) && (t0 == 0u || t1 == num_tiles))
            poisoned = !cta_wait_halos(a.wait, hseq, &sh.ok, csync);

        T K[N];
#pragma unroll
        for (int r = 0; r < N; ++r) K[r] = (T)0;
        int pending_unit = -1;

        auto run_phase = [&](auto update_tag, double* unit) {
            constexpr bool kUpdate = decltype(update_tag)::value;
            const bool reduces = !kUpdate || c.want_norm || (kCtl && a.reduce_b.world > 0);
            // one warp (rotating) adds the tile's 256 terms in a fixed order, see dek1_kernel
            auto flush_tile_partial = [&](unsigned cq_prev) {
                if (pending_unit >= 0) {
                    if (wq == (int)(cq_prev & (unsigned)(kWarps - 1))) {
                        const double* v = s_term + (cq_prev & 1u) * kThreads + lq;
                        double x = ((v[0] + v[32]) + (v[64] + v[96])) + ((v[128] + v[160]) + (v[192] + v[224]));
                        x = warp_sum(x);
                        if (lq == 0) unit[pending_unit] = x;
                    }
                    pending_unit = -1;
                }
            };
#pragma unroll 1
            for (unsigned j = 0; j < t1 - t0; ++j, ++cq) {
                const unsigned tile = kUpdate ? (t1 - 1u - j) : (t0 + j);
                const unsigned stage = cq % S;
                const Coord co = tile_coord<Shape>(g, tile, tiles_x);
                FrontRegs<IVP, T, N> fr;
                front_load<IVP, T, N, false, kCtl>(g, c, mean_in, halo_lo, halo_hi, co, fr);     // ring / fill values only
                mbar_wait(full + stage, (phase_bits >> stage) & 1u);
                phase_bits ^= (1u << stage);
                const T* st = s_stage + (size_t)stage * Cfg::STAGE_ELEMS + tid;
#pragma unroll
                for (int kd = 0; kd < N; ++kd) fr.raw[kd] = co.valid ? st[kd * kThreads] : (T)0;
                T mp[N], fx, jac;
                front_eval<IVP, T, N>(g, c, co, fr, s_at0 + (cq & 1u) * AT_STRIDE, mp, fx, jac, csync);
                if (tid == 0) mbar_arrive(empty + stage);       // every consumer has its column in registers (barrier above)
                if (reduces) flush_tile_partial(cq - 1u);
                double term = 0.0;
                if (co.valid) {
                    const T z = c.p[1] * mp[1] - fx;
                    if constexpr (!kUpdate) {
                        term = (double)z * (double)z;
                    } else {
                        T m0 = (T)0;
                        T* mo_ptr = mean_out + co.flat;
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            const T mo = c.p[r] * (mp[r] - K[r] * z);
                            *mo_ptr = mo;
                            mo_ptr += g.d;
                            if (r == 0) m0 = mo;
                        }
                        const T ref = tabs<T>(m0);
                        if (ref_out) ref_out[co.flat] = ref;
                        if (c.want_norm) {
                            const double tol = (double)c.abstol + (double)c.reltol * (double)ref;
                            term = 1.0 / (tol * tol);
                        }
                    }
                }
                if (reduces) {
                    s_term[(cq & 1u) * kThreads + tid] = term;
                    pending_unit = (int)tile;
                }
            }
            if (reduces) {
                csync();
                flush_tile_partial(cq - 1u);
                if (tid == 0 && poisoned && t1 > t0) unit[t0] = nan("");
            }
        };

        // ---- phase A, barrier, gain -----------------------------------------------------------------------------------
        run_phase(std::false_type{}, a.red_a.unit);
        const double err_cal = ek0_mid_barrier<T, N>(a, sh, cdr, att, s_red, K, csync);
        // ---- phase B --------------------------------------------------------------------------------------------------
        run_phase(std::true_type{}, a.red_b.unit);
        if (kCtl) fence_proxy_async_global();
        ek0_end_barrier<T, N, kCtl>(a, sh, c, att, err_cal, s_red, csync);
    }
}

}  // namespace tdx
