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

enum : int { IVP_VDP = 0, IVP_LORENZ96 = 1, IVP_BRUSS = 2, IVP_FHN2D = 3, IVP_BURGERS1D = 4, IVP_WAVE1D = 5, IVP_THREEBODY = 6, IVP_PLEIADES = 7,
              IVP_EXTERNAL = 8 };

// Local shard geometry.  Flat local state index = field * npts + row * cols + col.
struct Geo {
    int kind, nf;
    int rows, cols;             // local extent in points
    long long npts, d;          // rows * cols,  nf * npts
    long long grow0, gcol0;     // global coordinates of the local (0, 0)
    long long grows, gcols;     // global extent
    double prm[8];
    const void* ext_f;          // IVP_EXTERNAL: f(t, y) and diag J at the predicted state, evaluated by the caller (device arrays)
    const void* ext_j;          // (null: zero Jacobian diagonal)
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

// A load of state that another CTA of the SAME kernel may have written (multi-attempt kernels): L2 only.  The read-only path
// (ld.global.nc) is not coherent within a kernel; single-attempt kernels keep it.
template <bool kCoherent, typename T>
__device__ __forceinline__ T ld_state(const T* p) {
    if constexpr (kCoherent) return __ldcg(p);
    else return __ldg(p);
}

// The constants of one attempt as a multi-attempt (controller-driven) kernel sees them: what comes with the launch (Ql,
// tolerances, flags) stays in the constant bank, where every use is an instruction operand; only the step's own vectors
// (rebuilt by thread 0 from the controller block before each attempt) live in shared memory.  A whole Consts in shared
// memory costs the tile loop a load per use, re-issued after every store that might alias it.
template <typename T, int N>
struct StepView {
    const T* p;
    const T* pinv;
    const T* ql;
    T dt, abstol, reltol;
    int want_norm;
    unsigned int flags;
};

template <class C>
__device__ __forceinline__ auto QL(const C& c, int i, int j) {
    return c.ql[i * (i + 1) / 2 + j];
}

// Nordsieck preconditioner vectors  p[k] = |dt|^(nu-k+1/2) / (nu-k)!,  p_inv[k] = 1 / p[k]   (iwp.py:65-73).
// Built from sqrt, multiplications and divisions only (used by the device-resident step controller; the host path
// evaluates the same vectors with libm's pow like NumPy does -- the two agree to a few ulp).
__host__ __device__ inline void nordsieck_vectors(int nu, double dt, double* p, double* pinv) {
    const double adt = dt < 0.0 ? -dt : dt;
    // powers and factorials first (a short chain of multiplications), then all the divisions: they are independent of one
    // another, so the controller's single thread overlaps them instead of waiting for one after the other
    double pw[kMaxN], fact[kMaxN];
    double w = sqrt(adt), f = 1.0;
#pragma unroll
    for (int q = 0; q < kMaxN; ++q) {     // q = nu - k
        if (q > 1) f *= (double)q;
        pw[q] = w;
        fact[q] = f;
        w *= adt;
    }
#pragma unroll
    for (int q = 0; q < kMaxN; ++q) {
        if (q <= nu) {
            const int k = nu - q;
            if (p) p[k] = pw[q] / fact[q];
            if (pinv) pinv[k] = fact[q] / pw[q];
        }
    }
}

// Device-resident step controller (odefilter.py:171-223 + step.py:28-108 in device memory): the step kernels read the
// step to attempt and which state buffer is current from this block, and the last CTA of every attempt decides
// accept / reject, proposes the next dt and advances the buffer ring.  A multi-attempt launch (tdx_*_run_attempts) loops
// over attempts inside ONE persistent cooperative kernel: ``epoch`` is the in-kernel grid barrier that publishes every
// decision to the other CTAs.  The host looks at the block once per launch -- or, for dense output, follows
// ``host_feed`` (mapped pinned memory) while the kernel is running.
constexpr int kCtlHistory = 256;
constexpr int kMaxRing = 8;              // state buffers a controller-driven solve can rotate through

// What the kernel tells the host while it is running, and what the host tells the kernel (mapped pinned memory).
struct HostFeed {
    volatile unsigned long long accepted;      // kernel -> host: accepted states so far (state j lives in ring slot j % nbuf)
    volatile unsigned long long attempts;
    volatile unsigned long long reserved;
    volatile unsigned long long abort;         // host -> kernel: stop at the next attempt boundary
    volatile double hist_t[kCtlHistory];       // time / step of accepted state j at [j % kCtlHistory]
    volatile double hist_dt[kCtlHistory];
    volatile long long hist_att[kCtlHistory];  // attempts made when state j was accepted
};

struct StepCtl {
    double p[kMaxN], pinv[kMaxN];        // Nordsieck vectors of the step to attempt next
    double t, dt, tmax;
    double abstol, reltol, safety_scale, min_change, max_change, constant_dt;
    int adaptive, nu;
    int cur;                              // ring slot that holds the current state
    int done;                             // t reached tmax (or the norm became NaN): further attempts do nothing
    int failed;                           // 1: the error estimate became NaN; >= 2: an in-kernel wait timed out (TDX_FAIL_*)
    int nbuf;                             // ring size (2: plain ping-pong)
    long long attempts, accepted;
    double last_norm;
    unsigned long long epoch;             // attempts whose decision has been published (multi-attempt kernels spin on it)
    // time stops (odefilter.py:355-373): sorted device array, index of the stop the solve is heading for
    const double* stops;
    int n_stops, next_stop;
    // dense output: progress reports for a host that follows the solve (null: nobody follows)
    HostFeed* feed;
    long long wait_timeout_ns;            // in-kernel waits give up after this long and fail the solve
    double hist_t[kCtlHistory];           // times of the accepted states, ring indexed by (accepted - 1) % kCtlHistory
    double hist_dt[kCtlHistory];          // the accepted steps
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kDefaultWaitNs = 4000000000ull;
// why a controller-driven solve stopped (StepCtl::failed / tdx_ctl_state::failed)
enum : int { TDX_FAIL_NAN = 1, TDX_FAIL_POISONED = 2, TDX_FAIL_GRID_BARRIER = 3, TDX_FAIL_HOST_FEED = 4, TDX_FAIL_PEER = 5 };

// development aid (-DTDX_TIMELINE): CTA 0 records when it passes the stations of an attempt (scripts/attempt_timeline.py)
#ifdef TDX_TIMELINE
#define TDX_STAMP(buf, att, k)                                                                                   \
    do {                                                                                                         \
        if (blockIdx.x == 0) (buf)[2048 + ((att) & 63) * 8 + (k)] = (double)global_timer_ns();                   \
    } while (0)
#define TDX_STAMP_ANY(buf, att, k) (buf)[2048 + ((att) & 63) * 8 + (k)] = (double)global_timer_ns()
#else
#define TDX_STAMP(buf, att, k) do { } while (0)
#define TDX_STAMP_ANY(buf, att, k) do { } while (0)
#endif

// odefilter.py:355-373 (_TimeStopper.adjust_dt_to_time_stops) on the controller's stop list
__device__ inline double ctl_adjust_to_stops(StepCtl* ctl, double t, double dt) {
    if (ctl->stops == nullptr) return dt;
    if (ctl->next_stop < ctl->n_stops && t >= ctl->stops[ctl->next_stop]) ctl->next_stop += 1;
    const double loc = ctl->next_stop < ctl->n_stops ? ctl->stops[ctl->next_stop] : INFINITY;
    if (t + dt > loc) dt = loc - t;
    return dt;
}

// after one attempt: ``sum`` = sum_i (dt err_i / tol_i)^2 over ALL ranks (ignored for ConstantSteps)
__device__ inline void ctl_update(StepCtl* ctl, double sum, double d_global) {
    const double t = ctl->t, dt = ctl->dt;
    ctl->attempts += 1;
    bool accepted = true;
    double dt_next;
    if (ctl->adaptive) {
        const double norm = sqrt(sum / d_global);                                   // step.py:101-104
        ctl->last_norm = norm;
        if (norm != norm) {                                                          // odefilter.py:179-219 would spin forever
            if (ctl->failed == 0) ctl->failed = 1;
            ctl->done = 1;
            return;
        }
        accepted = norm < 1.0;                                                       // step.py:90-91
        double factor = ctl->safety_scale * pow(norm != 0.0 ? 1.0 / norm : INFINITY, 1.0 / (double)(ctl->nu + 1));   // step.py:80-88
        factor = fmax(ctl->min_change, fmin(factor, ctl->max_change));
        dt_next = fmin(factor * dt, ctl->tmax - (accepted ? t + dt : t));           // odefilter.py:214-219
    } else {
        if (sum != sum) {                                                            // a poisoned attempt (timed-out wait)
            if (ctl->failed == 0) ctl->failed = TDX_FAIL_POISONED;
            ctl->done = 1;
            return;
        }
        dt_next = fmin(ctl->constant_dt, ctl->tmax - (t + dt));
    }
    if (accepted) {
        const long long slot = ctl->accepted % kCtlHistory;
        ctl->hist_t[slot] = t + dt;
        ctl->hist_dt[slot] = dt;
        ctl->accepted += 1;
        ctl->t = t + dt;
        ctl->cur = (ctl->cur + 1) % ctl->nbuf;
        if (!(ctl->t < ctl->tmax)) ctl->done = 1;                                    // odefilter.py:111 ``while state.t < tmax``
        dt_next = ctl_adjust_to_stops(ctl, ctl->t, dt_next);                         // odefilter.py:134-135
        if (ctl->feed != nullptr) {
            HostFeed* fd = ctl->feed;
            const long long hs = ctl->accepted % kCtlHistory;                        // state j at [j % kCtlHistory]
            fd->hist_t[hs] = t + dt;
            fd->hist_dt[hs] = dt;
            fd->hist_att[hs] = ctl->attempts;
        }
    }
    if (!(dt_next > 0.0)) ctl->done = 1;
    ctl->dt = dt_next;
    nordsieck_vectors(ctl->nu, dt_next, ctl->p, ctl->pinv);
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
template <typename T, int N, class C = Consts<T, N>>
__device__ __forceinline__ void predict_column(const T (&raw)[N], const C& c, T (&mp)[N]) {
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
template <typename T, int N, bool kCoherent = false, class C = Consts<T, N>>
struct MeanAt {
    const T* __restrict__ mean;
    long long d;
    const C* c;
    __device__ __forceinline__ T operator()(long long idx) const {
        T acc = c->pinv[0] * ld_state<kCoherent>(mean + idx);
#pragma unroll
        for (int k = 1; k < N; ++k)
            acc = fma((T)A_entry<N>(0, k), c->pinv[k] * ld_state<kCoherent>(mean + (long long)k * d + idx), acc);
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
__host__ __device__ inline long long num_tiles(int rows, int cols) {
    return (long long)((cols + Shape::TC - 1) / Shape::TC) * ((rows + Shape::TR - 1) / Shape::TR);
}

// coordinates of this thread in tile number ``tile`` (tiles are numbered row-major over the shard)
template <class Shape>
__device__ __forceinline__ Coord tile_coord(const Geo& g, unsigned tile, unsigned tiles_x) {
    Coord c;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
    c.field = w / (Shape::TR * Shape::SEGS);
    const int rem = w - c.field * (Shape::TR * Shape::SEGS);
    c.tr = rem / Shape::SEGS;
    const int seg = rem - c.tr * Shape::SEGS;
    c.tcol = seg * 32 + lane;
    c.row = (int)tile_y * Shape::TR + c.tr;
    const int col0 = (int)tile_x * Shape::TC + seg * 32;
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

// ---------------------------------------------------------------------------------------------
// Vector fields.  The stencil reads the predicted state of neighbouring points.  Every CTA fills an EXTENDED tile
// in shared memory: its own TR x TC points plus a ring of HV rows above/below and HL / HR columns left/right.
// Ring points are computed by the first NF * NRING threads (one point each) from the mean through L2 (recomputed
// with MeanAt: bitwise what the owner computes), from the exchanged shard halos, or from the boundary rule of the
// problem (edge replication = coordinate clamping, constant pads, cyclic wrap).  After ONE barrier the stencil
// is branch-free shared-memory reads.
//   ring_value(g, row0, col0, field, er, ec, at, halo_lo, halo_hi)   value of extended-tile point (er, ec)
//   eval(g, c, s_ext, own, fx, jac)                                  f_i and (df/dy)_ii
// ---------------------------------------------------------------------------------------------
template <class IVP>
struct ExtTile {
    using Shape = typename IVP::Shape;
    static constexpr int HL = IVP::HL, HR = IVP::HR, HV = IVP::HV;
    static constexpr int ER = Shape::TR + 2 * HV, EC = Shape::TC + HL + HR;
    static constexpr int ELEMS = Shape::NF * ER * EC;
    static constexpr int NRING = ER * EC - Shape::TR * Shape::TC;       // ring points per field
    static_assert(Shape::NF * NRING <= kThreads, "one ring point per thread");
    __device__ static __forceinline__ int index(int field, int er, int ec) { return (field * ER + er) * EC + ec; }
    // this thread's own slot
    __device__ static __forceinline__ int own(const Coord& c) { return index(c.field, c.tr + HV, c.tcol + HL); }
    // h-th ring point of a field -> (er, ec)
    __device__ static __forceinline__ void ring(int h, int& er, int& ec) {
        if (HV > 0) {
            if (h < HV * EC) { er = h / EC; ec = h - er * EC; return; }
            h -= HV * EC;
            if (h < HV * EC) { er = h / EC; ec = h - er * EC; er += Shape::TR + HV; return; }
            h -= HV * EC;
        }
        // side columns of the TR middle rows: HL left ones then HR right ones per row
        const int per_row = HL + HR;
        const int r = h / per_row, k = h - r * per_row;
        er = HV + r;
        ec = (k < HL) ? k : (Shape::TC + k);
    }
};

// --- vanderpol: ivp.py:28-49 (prm[1] != 0: vanderpol_julia, ivp.py:52-73).  One point, two fields (y1, y2). ---
template <typename T>
struct VanderPol {
    using Shape = TileShape<2, 1, 4>;
    static constexpr int HL = 0, HR = 0, HV = 0;
    static __host__ void halo_sizes(const Geo&, long long& lo, long long& hi) { lo = hi = 0; }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo&, int, int, int, int, int, const At&, const T*, const T*) {
        return (T)0;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_ext, T own, T& fx, T& jac) {
        using E = ExtTile<VanderPol<T>>;
        const T mu = (T)g.prm[0];
        const T y1 = s_ext[E::index(0, 0, c.tcol)];
        const T y2 = s_ext[E::index(1, 0, c.tcol)];
        const T damp = mu * ((T)1 - y1 * y1);
        if (c.field == 0) {
            fx = y2;
            jac = (T)0;
        } else if (g.prm[1] == 0.0) {
            fx = damp * y2 - y1;
            jac = damp;
        } else {                                   // vanderpol_julia (ivp.py:52-73): mu scales the whole right-hand side
            fx = mu * (((T)1 - y1 * y1) * y2 - y1);
            jac = damp;
        }
        (void)own;
    }
};

// --- lorenz96: ivp.py:268-294.  f_i = (y[i+1] - y[i-2]) y[i-1] - y[i] + F, cyclic; diag J = -1. ---
template <typename T>
struct Lorenz96 {
    using Shape = TileShape<1, 1, 8>;
    static constexpr int HL = 2, HR = 1, HV = 0;
    // lower halo = points (i0-2, i0-1), upper halo = point i0+cols (all cyclic)
    static __host__ void halo_sizes(const Geo& g, long long& lo, long long& hi) {
        const bool sharded = g.cols != g.gcols;
        lo = sharded ? 2 : 0;
        hi = sharded ? 1 : 0;
    }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo& g, int, int col0, int, int, int ec, const At& at,
                                                   const T* halo_lo, const T* halo_hi) {
        const long long col = (long long)col0 + ec - HL;          // shard-local, may be outside [0, cols)
        long long gi = (g.gcol0 + col) % g.gcols;                 // cyclic
        if (gi < 0) gi += g.gcols;
        const long long li = gi - g.gcol0;
        if (li >= 0 && li < g.cols) return at(li);
        if (col < 0 && col >= -2 && halo_lo) return __ldcg(halo_lo + (2 + col));
        if (col == g.cols && halo_hi) return __ldcg(halo_hi);
        return (T)0;                                               // only feeds threads beyond the shard
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_ext, T own, T& fx, T& jac) {
        const T F = (T)g.prm[0];
        const T* p = s_ext + c.tcol;                               // p[HL] is the own point
        fx = (p[3] - p[0]) * p[1] - own + F;
        jac = (T)-1;
    }
};

// --- brusselator: ivp.py:219-265.  y = [u(N); v(N)], c = (N+1)^2/50, boundary pads u=1, v=3. -------
template <typename T>
struct Brusselator {
    using Shape = TileShape<2, 1, 4>;
    static constexpr int HL = 1, HR = 1, HV = 0;
    static __host__ void halo_sizes(const Geo& g, long long& lo, long long& hi) {
        lo = (g.gcol0 > 0) ? 2 : 0;                  // (u, v) of point i0-1
        hi = (g.gcol0 + g.cols < g.gcols) ? 2 : 0;   // (u, v) of point i0+cols
    }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo& g, int, int col0, int field, int, int ec, const At& at,
                                                   const T* halo_lo, const T* halo_hi) {
        const long long col = (long long)col0 + ec - HL;
        const long long gi = g.gcol0 + col;
        if (gi < 0 || gi >= g.gcols) return (field == 0) ? (T)1 : (T)3;      // ivp.py:233-236
        if (col >= 0 && col < g.cols) return at((long long)field * g.npts + col);
        if (col == -1 && halo_lo) return __ldcg(halo_lo + field);
        if (col == g.cols && halo_hi) return __ldcg(halo_hi + field);
        return (T)0;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_ext, T own, T& fx, T& jac) {
        using E = ExtTile<Brusselator<T>>;
        const T cc = (T)g.prm[0];
        const int here = E::index(c.field, 0, c.tcol + HL);
        const T left = s_ext[here - 1], right = s_ext[here + 1];
        const T u = s_ext[E::index(0, 0, c.tcol + HL)];
        const T v = s_ext[E::index(1, 0, c.tcol + HL)];
        const T conv = (left - (T)2 * own) + right;
        const T uu = u * u;
        if (c.field == 0) {
            fx = (T)1 + uu * v - (T)4 * u + cc * conv;
            jac = (T)2 * u * v - (T)4 - (T)2 * cc;
        } else {
            fx = (T)3 * u - uu * v + cc * conv;
            jac = -uu - (T)2 * cc;
        }
    }
};

// --- burgers_1d: ivp.py:76-107.  Interior: -x_i (x_i - x_{i-1}) / dx + nu (x_{i+1} - 2 x_i + x_{i-1}) / dx^2.  BOTH end
//     points get the same "cyclic" value built from x_0, x_1 and x_{N-2}: the ring maps the column left of the mesh to
//     x_{N-2} and the two columns right of it to x_0, x_1, so point 0 is the interior formula and point N-1 re-evaluates
//     point 0's.  diag J (the reference takes diag(jacfwd f), :106): -(2 x_i - x_{i-1}) / dx - 2 nu / dx^2, 0 at N-1.
//     One GPU only (prm[0] = 1/dx, prm[1] = nu/dx^2). -------------------------------------------------------------------
template <typename T>
struct Burgers1d {
    using Shape = TileShape<1, 1, 8>;
    static constexpr int HL = 1, HR = 2, HV = 0;
    static __host__ void halo_sizes(const Geo&, long long& lo, long long& hi) { lo = hi = 0; }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo& g, int, int col0, int, int, int ec, const At& at, const T*,
                                                   const T*) {
        const long long col = (long long)col0 + ec - HL;
        if (col >= 0 && col < g.cols) return at(col);
        if (col == -1) return at(g.cols - 2);
        if (col == g.cols) return at(0);
        if (col == g.cols + 1) return at(1);
        return (T)0;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_ext, T own, T& fx, T& jac) {
        const T inv_dx = (T)g.prm[0], nu_dx2 = (T)g.prm[1];
        const T* p = s_ext + c.tcol;                               // p[HL] is the own point
        const bool last = c.col == g.cols - 1;
        const T xl = p[0];                                         // x_{i-1}; x_{N-2} for i = 0 and for i = N-1
        const T xc = last ? p[2] : own;                            // the last point evaluates x_0's formula
        const T xr = last ? p[3] : p[2];
        fx = -xc * (xc - xl) * inv_dx + nu_dx2 * (xr - (T)2 * xc + xl);
        jac = last ? (T)0 : -((T)2 * xc - xl) * inv_dx - (T)2 * nu_dx2;
    }
};

// --- wave_1d: ivp.py:110-140.  y = [x(N); v(N)]: x' = v and v' = nu (x_{i+1} - 2 x_i + x_{i-1}) / dx^2 in the interior;
//     at the two end points x' = 0 and v' = x.  The Jacobian diagonal vanishes everywhere.  One GPU only
//     (prm[1] = nu/dx^2). ----------------------------------------------------------------------------------------------
template <typename T>
struct Wave1d {
    using Shape = TileShape<2, 1, 4>;
    static constexpr int HL = 1, HR = 1, HV = 0;
    static __host__ void halo_sizes(const Geo&, long long& lo, long long& hi) { lo = hi = 0; }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo& g, int, int col0, int field, int, int ec, const At& at,
                                                   const T*, const T*) {
        const long long col = (long long)col0 + ec - HL;
        if (col >= 0 && col < g.cols) return at((long long)field * g.npts + col);
        return (T)0;                                               // never used: the end points have their own rule
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_ext, T own, T& fx, T& jac) {
        using E = ExtTile<Wave1d<T>>;
        const T nu_dx2 = (T)g.prm[1];
        const bool edge = c.col == 0 || c.col == g.cols - 1;
        const int xi = E::index(0, 0, c.tcol + HL);
        const T x = s_ext[xi], v = s_ext[E::index(1, 0, c.tcol + HL)];
        if (c.field == 0) fx = edge ? (T)0 : v;
        else fx = edge ? x : nu_dx2 * (s_ext[xi + 1] - (T)2 * x + s_ext[xi - 1]);
        jac = (T)0;
        (void)own;
    }
};

template <typename T>
__device__ __forceinline__ T tsqrt(T x);          // specialised further down

// --- threebody (ivp.py:188-216) and pleiades (ivp.py:344-406): small dense systems, every coordinate reads the positions of
//     all bodies.  The flat state is laid out as ONE field of d points (d = 4 / 28 <= one tile), so after the front end's
//     barrier the tile in shared memory IS the predicted state vector and coordinate j sits at index j.  The Jacobian
//     diagonal vanishes for both (the accelerations depend on positions only; threebody's is zero by definition, :204-205).
template <typename T>
struct ThreeBody {
    using Shape = TileShape<1, 1, 8>;
    static constexpr int HL = 0, HR = 0, HV = 0;
    static __host__ void halo_sizes(const Geo&, long long& lo, long long& hi) { lo = hi = 0; }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo&, int, int, int, int, int, const At&, const T*, const T*) {
        return (T)0;
    }
    __device__ static __forceinline__ void eval(const Geo&, const Coord& c, const T* y, T own, T& fx, T& jac) {
        const T mu = (T)0.012277471, mp = (T)1 - mu;
        jac = (T)0;
        (void)own;
        if (c.col < 2) {                          // positions: y' = velocity
            fx = y[c.col + 2];
            return;
        }
        const T a = y[0] + mu, b = y[0] - mp;
        const T s1 = a * a + y[1] * y[1], s2 = b * b + y[1] * y[1];
        const T D1 = s1 * tsqrt<T>(s1), D2 = s2 * tsqrt<T>(s2);            // (...)^(3/2)
        if (c.col == 2) fx = y[0] + (T)2 * y[3] - mp * a / D1 - mu * b / D2;
        else fx = y[1] - (T)2 * y[2] - mp * y[1] / D1 - mu * y[1] / D2;
    }
};

template <typename T>
struct Pleiades {
    using Shape = TileShape<1, 1, 8>;
    static constexpr int HL = 0, HR = 0, HV = 0;
    static __host__ void halo_sizes(const Geo&, long long& lo, long long& hi) { lo = hi = 0; }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo&, int, int, int, int, int, const At&, const T*, const T*) {
        return (T)0;
    }
    // y = [x(7); y(7); x'(7); y'(7)];  x_i'' = sum_{j != i} m_j (x_j - x_i) / r_ij^3, m_j = j + 1 (the reference's 0/0 at
    // j = i is mapped to 0 by nan_to_num, :384-385)
    __device__ static __forceinline__ void eval(const Geo&, const Coord& c, const T* y, T own, T& fx, T& jac) {
        jac = (T)0;
        (void)own;
        if (c.col < 14) {
            fx = y[(c.col + 14) % 28];            // % keeps the threads beyond the state inside the tile
            return;
        }
        const int i = (c.col - 14) % 7;
        const bool is_x = c.col < 21;
        const T xi = y[i], yi = y[7 + i];
        T acc = (T)0;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const T dx = y[j] - xi, dy = y[7 + j] - yi;
            const T s = dx * dx + dy * dy;
            const T r3 = s * tsqrt<T>(s);
            const T term = (T)(j + 1) * (is_x ? dx : dy) / r3;
            acc += (j == i) ? (T)0 : term;
        }
        fx = acc;
    }
};

// --- a vector field the library does not know: the caller evaluated f (and optionally the Jacobian diagonal) at the predicted
//     state E0 P A P^-1 m (tdx_predict_state) with its own code and hands the arrays in (tdx_set_external).  Everything else
//     of the attempt -- prediction, the batched QR, update, error estimate, norm -- is the same fused kernel.
template <typename T>
struct External {
    using Shape = TileShape<1, 1, 8>;
    static constexpr int HL = 0, HR = 0, HV = 0;
    static __host__ void halo_sizes(const Geo&, long long& lo, long long& hi) { lo = hi = 0; }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo&, int, int, int, int, int, const At&, const T*, const T*) {
        return (T)0;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T*, T own, T& fx, T& jac) {
        (void)own;
        fx = __ldg(reinterpret_cast<const T*>(g.ext_f) + c.flat);
        jac = g.ext_j ? __ldg(reinterpret_cast<const T*>(g.ext_j) + c.flat) : (T)0;
    }
};

// --- fhn_2d: ivp.py:409-505.  y = [u; v] on an (nx, ny) grid, 5-point Laplacian, Neumann by edge
//     replication; u' = a Lap u + u - u^3 - v + k ; v' = (b Lap v + u - v) / tau. --------------------
template <typename T>
struct Fhn2d {
    using Shape = TileShape<2, 4, 1>;
    static constexpr int HL = 1, HR = 1, HV = 1;
    static __host__ void halo_sizes(const Geo& g, long long& lo, long long& hi) {
        lo = (g.grow0 > 0) ? 2LL * g.cols : 0;                 // row above, both fields
        hi = (g.grow0 + g.rows < g.grows) ? 2LL * g.cols : 0;  // row below
    }
    template <class At>
    __device__ static __forceinline__ T ring_value(const Geo& g, int row0, int col0, int field, int er, int ec,
                                                   const At& at, const T* halo_lo, const T* halo_hi) {
        // edge replication (ivp.py:488-490) == clamp the coordinates to the global grid
        long long col = (long long)col0 + ec - HL;
        col = col < 0 ? 0 : (col > g.cols - 1 ? g.cols - 1 : col);
        long long gr = g.grow0 + row0 + er - HV;
        gr = gr < 0 ? 0 : (gr > g.grows - 1 ? g.grows - 1 : gr);
        const long long row = gr - g.grow0;
        if (row >= 0 && row < g.rows) return at((long long)field * g.npts + row * g.cols + col);
        if (row < 0) return halo_lo ? __ldcg(halo_lo + (long long)field * g.cols + col) : (T)0;
        return halo_hi ? __ldcg(halo_hi + (long long)field * g.cols + col) : (T)0;
    }
    __device__ static __forceinline__ void eval(const Geo& g, const Coord& c, const T* s_ext, T own, T& fx, T& jac) {
        using E = ExtTile<Fhn2d<T>>;
        const T inv_dx2 = (T)g.prm[0], a = (T)g.prm[1], b = (T)g.prm[2], k = (T)g.prm[3], inv_tau = (T)g.prm[5];
        const int here = E::index(c.field, c.tr + HV, c.tcol + HL);
        const T up = s_ext[here - E::EC], dn = s_ext[here + E::EC], lf = s_ext[here - 1], rt = s_ext[here + 1];
        const T lap = (((up + dn) + (lf + rt)) - (T)4 * own) * inv_dx2;
        const long long gr = g.grow0 + c.row;
        const int nrep = (int)(gr == 0) + (int)(gr == g.grows - 1) + (int)(c.col == 0) + (int)(c.col == g.cols - 1);
        const T dl = -(T)(4 - nrep) * inv_dx2;   // ivp.py:447-463: 2 (corner), 3 (edge), 4 (interior)
        const T u = s_ext[E::index(0, c.tr + HV, c.tcol + HL)];
        const T v = s_ext[E::index(1, c.tr + HV, c.tcol + HL)];
        if (c.field == 0) {
            fx = a * lap + u - u * u * u - v + k;
            jac = a * dl + (T)1 - (T)3 * (u * u);
        } else {
            fx = (b * lap + u - v) * inv_tau;
            jac = (b * dl - (T)1) * inv_tau;
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

template <typename T, int N, class C = Consts<T, N>>
__device__ __forceinline__ void stacked_qr(T (&B)[N][N], T sigma, const C& c) {
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
#ifdef TDX_DEBUG_WAITS
// development aid: an mbarrier wait that gives up after 2 s and leaves a note (code, CTA, attempt) in the debug buffer
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, double* dbg, int code, int att) {
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (global_timer_ns() - t0 > 2000000000ull) {
            dbg[3000 + code] = (double)(1 + blockIdx.x) + 1e-3 * att;
            return;
        }
    }
}
#define TDX_MBAR_WAIT(bar, parity, dbg, code, att) mbar_wait_dbg(bar, parity, dbg, code, att)
#else
#define TDX_MBAR_WAIT(bar, parity, dbg, code, att) mbar_wait(bar, parity)
#endif
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 eviction-priority policies (createpolicy) for streaming data with a known reuse pattern
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void st_global_hint(double* p, double v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_global_hint(float* p, float v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
// Ampere-style asynchronous copies (SASS LDGSTS) of one element with a per-lane destination: used where the shared-memory
// layout is padded and a linear bulk copy cannot be used
// kCoherent (multi-attempt kernels: the source may have been written earlier in the same kernel): a synchronous L2 load
template <bool kCoherent = false>
__device__ __forceinline__ void cp_async_elem(double* smem_dst, const double* gmem_src) {
    if constexpr (kCoherent) *smem_dst = __ldcg(gmem_src);
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
template <bool kCoherent = false>
__device__ __forceinline__ void cp_async_elem(float* smem_dst, const float* gmem_src) {
    if constexpr (kCoherent) *smem_dst = __ldcg(gmem_src);
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// 16 bytes, L2 only (both addresses 16-byte aligned)
__device__ __forceinline__ void cp_async_unit(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// orders generic-proxy and async-proxy (TMA) accesses to global memory: a state buffer written with ordinary stores in one
// attempt is read with cp.async.bulk in the next attempt of the same (multi-attempt) kernel, and vice versa
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Peer-memory exchange (NVLink): per-rank device block, mapped by the neighbours through CUDA IPC.
constexpr int kMaxWorld = 16;
constexpr int kRedChunks = 16;           // fixed global chunks of a reduction (see RedPlan)

struct CommHeader {
    unsigned long long halo_flag_lo[2];              // sequence number of the halo the LOWER neighbour delivered (per parity)
    unsigned long long halo_flag_hi[2];
    unsigned long long red_flag[2][kMaxWorld];       // per parity, per source rank
    double red_val[2][kMaxWorld][kRedChunks];        // per parity, per source rank: its chunk sums in chunk order
    int red_cnt[2][kMaxWorld];                       // how many of them are meaningful
    unsigned long long error;                        // a flag wait timed out
    unsigned int pack_count[2];                      // CTAs of the local pack kernel that have finished (per direction)
    unsigned long long wait_timeout_ns;              // how long a flag wait may take before it fails the attempt
    unsigned long long pad[3];
};
static_assert(sizeof(CommHeader) % 16 == 0, "halo storage behind the header must stay 16-byte aligned");

// What a step kernel has to wait for before it reads halos.  Flags and halo buffers exist once per parity of the halo
// sequence number (a neighbour may already deliver the next attempt's halo while this rank still reads the current one):
// ``flag_*`` point at element [0] of the two-element flag arrays, ``seq`` is the sequence number of the kernel's FIRST attempt.
struct CommWait {
    const unsigned long long* flag_lo;
    const unsigned long long* flag_hi;
    unsigned long long seq;
    unsigned long long* error;
    unsigned long long timeout_ns;
};

struct CommReduce {                                  // fused all-reduce of the kernel's scalar result (world == 0: off)
    CommHeader* local;
    CommHeader* peer[kMaxWorld];
    int rank, world;
    unsigned long long seq;                          // of the kernel's first reduction
    unsigned long long timeout_ns;
};

// spin until *flag >= seq; gives up after the timeout (sets *error, returns false) so that a dead peer cannot hang the GPU.
// The caller poisons the attempt (NaN norm) so that the failure reaches every rank's step controller / host loop.
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long seq,
                                           unsigned long long* error, unsigned long long timeout_ns = kDefaultWaitNs) {
    if (flag == nullptr) return true;
    // relaxed polls, ONE acquire fence once the value is there (see spin_until_gpu)
    if (ld_relaxed_sys(flag) < seq) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_relaxed_sys(flag) < seq) {
            if (global_timer_ns() - t0 > (timeout_ns ? timeout_ns : kDefaultWaitNs)) {
                if (error) *error = 1ull;
                return false;
            }
            __nanosleep(40);
        }
    }
    fence_acq_rel_sys();
    return true;
}
// the same at GPU scope (in-kernel grid barriers)
__device__ __forceinline__ bool spin_until_gpu(const unsigned long long* flag, unsigned long long seq,
                                               unsigned long long timeout_ns = kDefaultWaitNs) {
    // Poll with RELAXED loads and fence once the value is there (an acquire load per poll makes the SM invalidate its L1 every
    // time; measured: no difference for the CTA that shares the SM, kept because it is the cheaper protocol).
    if (ld_relaxed_gpu(flag) >= seq) {
        fence_acq_rel_gpu();
        return true;
    }
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_relaxed_gpu(flag) < seq) {
        // the release is at most a few microseconds away in the common case: poll back to back first
        if (++spins > 64u) {
            if (global_timer_ns() - t0 > (timeout_ns ? timeout_ns : kDefaultWaitNs)) return false;
            __nanosleep(32);
        }
    }
    fence_acq_rel_gpu();
    return true;
}
// Who takes part in a CTA-level helper: the whole CTA (barrier 0), or the first COUNT threads through a named barrier
// (warp-specialised kernels whose producer warp must not be waited for).
struct FullSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
    __device__ __forceinline__ int threads() const { return (int)blockDim.x; }
};
template <int ID, int COUNT>
struct NamedSync {
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
    __device__ __forceinline__ int threads() const { return COUNT; }
};

// One thread per CTA waits for the halos of sequence number ``seq``, the barrier releases the rest.  Returns false
// (CTA-uniform) if a wait timed out: the halo buffers then hold stale data and the caller must poison its result.
template <class Sync = FullSync>
__device__ __forceinline__ bool cta_wait_halos(const CommWait& w, unsigned long long seq, int* s_ok, Sync sync = Sync()) {
    if (w.flag_lo == nullptr && w.flag_hi == nullptr) return true;
    if (threadIdx.x == 0) {
        const int par = (int)(seq & 1ull);
        bool ok = spin_until(w.flag_lo ? w.flag_lo + par : nullptr, seq, w.error, w.timeout_ns);
        ok = spin_until(w.flag_hi ? w.flag_hi + par : nullptr, seq, w.error, w.timeout_ns) && ok;
        *s_ok = ok ? 1 : 0;
    }
    sync();
    const bool ok = *s_ok != 0;
    return ok;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Deterministic, sharding-independent grid-wide sums (step.py:101-104 is ONE norm over all d coordinates).
//
// Every kernel stores one partial per UNIT of work -- a 256-coordinate tile, or one grid row of a 256-column strip -- whose
// value is a fixed function of the unit's coordinates (warp shuffle tree, then the CTA's warps in order), whichever CTA or
// GPU computes it.  The global index range of the units is cut into kRedChunks equal chunks; a chunk sum adds the chunk's unit
// partials in a fixed order (32 strided lanes, shuffle tree), and the result is the sum of the chunk sums in chunk order.
// With several GPUs every rank sends the sums of the chunks (chunk parts) it owns to all peers and every rank adds all of
// them in (rank, chunk) order: when the shard boundaries are chunk boundaries -- 2, 4, 8, 16 equal shards of a grid whose
// unit count divides by 16 -- that is the SAME sequence of additions as on one GPU, so the norm, the accept / reject
// decisions and the proposed steps are bit-identical at 1 / 2 / 4 / 8 GPUs.  Otherwise it is still deterministic and identical
// on all ranks, just not independent of the number of ranks.
// ---------------------------------------------------------------------------------------------
struct RedPlan {
    double* unit;                        // unit partials of this reduction, indexed by LOCAL unit id (null: per-CTA partials)
    unsigned int units_local;
    unsigned long long unit0;            // global id of local unit 0
    unsigned long long units_global;
};

__device__ __forceinline__ unsigned long long chunk_begin(const RedPlan& rp, int c) {
    return ((unsigned long long)c * rp.units_global) / (unsigned long long)kRedChunks;
}

// The CTA's arrival at a grid-wide rendezvous: returns (CTA-uniform) whether this CTA is the LAST of the grid to arrive.
// Everything the CTA's threads wrote to global memory before the call is visible to the last CTA after it.
template <class Sync = FullSync>
__device__ __forceinline__ bool cta_arrive_last(unsigned int* ticket, int* s_flag, Sync sync = Sync()) {
    sync();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        const bool last = (t == gridDim.x - 1);
        if (last) {
            *ticket = 0u;                 // everybody has arrived; the next rendezvous starts after the last CTA's release
            __threadfence();
        }
        *s_flag = last ? 1 : 0;
    }
    sync();
    return *s_flag != 0;
}

// Run by all participating threads of ONE CTA (the last to arrive): chunk sums of the local unit partials, exchange with the
// peers (cr != null && cr->world > 0, reduction sequence number ``seq``), total in (rank, chunk) order.  The total is valid in
// thread 0; a timed-out peer wait makes it NaN.  ``s_red`` needs kRedChunks + 2 doubles.
template <class Sync = FullSync>
__device__ __forceinline__ double reduce_units(const RedPlan& rp, const CommReduce* cr, unsigned long long seq, double* s_red,
                                               Sync sync = Sync()) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = sync.threads() >> 5;
    const unsigned long long lo_all = rp.unit0, hi_all = rp.unit0 + rp.units_local;
    sync();                                                      // s_red may still be read by a previous use
    if (threadIdx.x == 0) s_red[kRedChunks + 1] = 0.0;           // "a peer wait timed out"
    if (nw * 2 >= kRedChunks) {
        // two chunks per warp: issue the loads of both before the first shuffle tree (one L2 latency instead of two)
        double acc[2] = {0.0, 0.0};
        unsigned long long b0[2], b1[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = w + j * nw;
            b0[j] = b1[j] = 0ull;
            if (c < kRedChunks) {
                b0[j] = chunk_begin(rp, c);
                b1[j] = chunk_begin(rp, c + 1);
                b0[j] = b0[j] > lo_all ? b0[j] : lo_all;
                b1[j] = b1[j] < hi_all ? b1[j] : hi_all;
            }
        }
        // lane l adds units l, l + 32, l + 64, ... of the chunk IN THAT ORDER; the loads of kBatch of them (per chunk) are issued
        // before the first addition, so that a chunk of thousands of units costs a few L2 round trips, not one per unit
        constexpr int kBatch = 16;
        const double* base = rp.unit - lo_all;
        unsigned long long u0 = b0[0] + lane, u1 = b0[1] + lane;
        while (u0 + 32ull * (kBatch - 1) < b1[0] && u1 + 32ull * (kBatch - 1) < b1[1]) {
            double v0[kBatch], v1[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                v0[i] = __ldcg(base + u0 + 32 * i);
                v1[i] = __ldcg(base + u1 + 32 * i);
            }
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                acc[0] += v0[i];
                acc[1] += v1[i];
            }
            u0 += 32ull * kBatch;
            u1 += 32ull * kBatch;
        }
        constexpr int kTail = 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned long long u = j == 0 ? u0 : u1;
            double x = acc[j];
            // a warp whose other chunk is empty (sharded runs: a rank owns 16 / G chunks) or shorter still takes deep batches
            while (u + 32ull * (kBatch - 1) < b1[j]) {
                double v[kBatch];
#pragma unroll
                for (int i = 0; i < kBatch; ++i) v[i] = __ldcg(base + u + 32 * i);
#pragma unroll
                for (int i = 0; i < kBatch; ++i) x += v[i];
                u += 32ull * kBatch;
            }
            while (u + 32ull * (kTail - 1) < b1[j]) {
                double v[kTail];
#pragma unroll
                for (int i = 0; i < kTail; ++i) v[i] = __ldcg(base + u + 32 * i);
#pragma unroll
                for (int i = 0; i < kTail; ++i) x += v[i];
                u += 32ull * kTail;
            }
            for (; u < b1[j]; u += 32) x += __ldcg(base + u);
            acc[j] = x;
        }
        acc[0] = warp_sum(acc[0]);
        acc[1] = warp_sum(acc[1]);
        if (lane == 0) {
            if (w < kRedChunks) s_red[w] = acc[0];
            if (w + nw < kRedChunks) s_red[w + nw] = acc[1];
        }
    } else {
        for (int c = w; c < kRedChunks; c += nw) {
            unsigned long long b0 = chunk_begin(rp, c), b1 = chunk_begin(rp, c + 1);
            b0 = b0 > lo_all ? b0 : lo_all;
            b1 = b1 < hi_all ? b1 : hi_all;
            double acc = 0.0;
            for (unsigned long long u = b0 + lane; u < b1; u += 32) acc += __ldcg(rp.unit + (u - lo_all));
            acc = warp_sum(acc);
            if (lane == 0) s_red[c] = acc;
        }
    }
    sync();
    const bool comm = cr != nullptr && cr->world > 0;
    double total = 0.0;
    if (!comm) {
        if (threadIdx.x == 0) {
            for (int c = 0; c < kRedChunks; ++c) {
                const unsigned long long b0 = chunk_begin(rp, c), b1 = chunk_begin(rp, c + 1);
                if ((b0 > lo_all ? b0 : lo_all) < (b1 < hi_all ? b1 : hi_all)) total += s_red[c];
            }
        }
        return total;
    }
    const int parity = (int)(seq & 1ull);
    const int r = threadIdx.x;
    if (r < cr->world) {
        CommHeader* dst = cr->peer[r];
        int cnt = 0;
        for (int c = 0; c < kRedChunks; ++c) {
            const unsigned long long b0 = chunk_begin(rp, c), b1 = chunk_begin(rp, c + 1);
            if ((b0 > lo_all ? b0 : lo_all) < (b1 < hi_all ? b1 : hi_all)) dst->red_val[parity][cr->rank][cnt++] = s_red[c];
        }
        dst->red_cnt[parity][cr->rank] = cnt;
        // the chunk sums, the count and the flag are written by THIS thread: the release store alone orders them (a separate
        // system-scope fence in front of it costs one more NVLink round trip per exchange)
        st_release_sys(&dst->red_flag[parity][cr->rank], seq);
    }
    sync();
    // Warp 0 gathers: lane k waits for rank k's flag and loads rank k's count and ALL its value slots with independent loads (one
    // memory latency per lane instead of one per value: the slots were written by a peer and miss every cache); the sum itself
    // is formed by every lane of the warp in (rank, chunk) order from the other lanes' registers.
    static_assert(kMaxWorld <= 32, "one lane per rank");
    if ((threadIdx.x >> 5) == 0) {
        const int lane = threadIdx.x & 31;
        bool ok = true;
        int cnt = 0;
        double v[kRedChunks];
#pragma unroll
        for (int j = 0; j < kRedChunks; ++j) v[j] = 0.0;
        if (lane < cr->world) {
            ok = spin_until(&cr->local->red_flag[parity][lane], seq, &cr->local->error, cr->timeout_ns);
            cnt = *((volatile int*)&cr->local->red_cnt[parity][lane]);
#pragma unroll
            for (int j = 0; j < kRedChunks; ++j) v[j] = *((volatile double*)&cr->local->red_val[parity][lane][j]);
        }
        const bool failed = __any_sync(0xffffffffu, !ok);
        for (int k = 0; k < cr->world; ++k) {
            const int ck = __shfl_sync(0xffffffffu, cnt, k);
#pragma unroll
            for (int j = 0; j < kRedChunks; ++j) {
                const double x = __shfl_sync(0xffffffffu, v[j], k);
                if (j < ck) total += x;
            }
        }
        if (failed) total = nan("");
    }
    return total;
}

// Legacy per-CTA partial sums (launches that cover only a part of the shard's tiles: TDX_STEP_TILES_INTERIOR / _RIM).
// Returns true in thread 0 of the last CTA, after ``*out`` has been written.
__device__ __forceinline__ bool grid_sum_cta(double v, double* s_red, double* partials, unsigned int* ticket, double* out) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
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
            if (out) *out = b;
            *ticket = 0u;
            return true;
        }
    }
    return false;
}

}  // namespace tdx
