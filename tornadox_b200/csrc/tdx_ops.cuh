// Host-side launch table: one Ops entry per (dtype, n), defined in the per-instantiation translation units.
#pragma once

#include <cstdlib>

#include "tdx_kernels.cuh"
#include "tdx_ek0_fused.cuh"

#ifndef TDX_EK0_W
#define TDX_EK0_W 256      // columns per strip (= threads per CTA) of the fused KroneckerEK0 kernel
#endif

namespace tdx {

struct StepHost {
    double p[kMaxN], pinv[kMaxN];
    double ql[kMaxN * (kMaxN + 1) / 2];   // packed for the actual n
    double dt, abstol, reltol;
    int want_norm;
    unsigned int flags;
};

struct Workspace {
    double* partials;        // >= grid doubles
    unsigned int* ticket;
    unsigned int* tile_counter;        // dynamic hand-out of the DiagonalEK1 tiles (zero between launches)
    void* mid;               // Ek0Mid<T, N>
    int reserve_ctas;        // CTAs left free on the device (room for a concurrent NCCL kernel)
    bool has_lo, has_hi;     // the shard has a lower / upper neighbour (halo buffers exist)
    CommWait wait;           // peer-memory halos: flags the kernel must see before it reads them (null otherwise)
    CommReduce reduce;       // fused all-reduce of the scalar result (world == 0: off)
    PackPeer<double> pack;   // fused halo delivery (pointers are dtype-agnostic; enabled == 0: off)
    long long halo_parity_stride_bytes;   // between the parity-0 and parity-1 halo buffers of the context (0: the caller's halos)
    double* unit_a;          // unit partials of the launch's (first) reduction, >= unit_cap doubles
    double* unit_b;          // ... of the second one (fused KroneckerEK0: phase B)
    long long unit_cap;
    unsigned long long* cl_flag;   // fused KroneckerEK0: "the attempt's new factor has been written"
    unsigned long long cl_seq;
    long long max_accepted;        // controller launches: stop after this many newly accepted steps (0: no limit)
};

// unit partials of a tile-based launch (one unit per tile, numbered row-major over the global tile grid)
template <class Shape>
inline RedPlan tile_red_plan(const Geo& g, double* unit, unsigned tiles_x, unsigned tiles_y) {
    RedPlan rp{};
    rp.unit = unit;
    rp.units_local = tiles_x * tiles_y;
    if (g.kind == IVP_FHN2D) {
        rp.unit0 = (unsigned long long)(g.grow0 / Shape::TR) * tiles_x;
        rp.units_global = (unsigned long long)((g.grows + Shape::TR - 1) / Shape::TR) * tiles_x;
    } else {
        rp.unit0 = (unsigned long long)(g.gcol0 / Shape::TC);
        rp.units_global = (unsigned long long)((g.gcols + Shape::TC - 1) / Shape::TC);
    }
    // a shard that does not start on a tile boundary of the global tiling: still a valid (rank-dependent) order
    if (rp.unit0 + rp.units_local > rp.units_global) rp.units_global = rp.unit0 + rp.units_local;
    return rp;
}

// Which tiles does a launch process?  all / interior (no halo reads) / rim (TDX_STEP_TILES_*).
// Rim = the first / last tile row (2-d grids) or the first / last tile (chains) on sides that have a neighbour.
inline TileRanges tile_ranges(unsigned flags, unsigned tiles_x, unsigned tiles_y, bool has_lo, bool has_hi,
                              bool row_sharded, bool waits) {
    const unsigned total = tiles_x * tiles_y;
    const unsigned want = flags & 6u;
    const unsigned edge = row_sharded ? tiles_x : 1u;
    unsigned lo = has_lo ? edge : 0u, hi = has_hi ? edge : 0u;
    if (lo + hi > total) {   // a single tile (row) touches both halos
        lo = total;
        hi = 0u;
    }
    if (want == 4u) return TileRanges{0u, lo, total - hi, hi, 0u, 0u, 0u};                 // rim only
    if (want == 2u) return TileRanges{lo, total - lo - hi, 0u, 0u, 0u, 0u, 0xffffffffu};   // interior only
    if (!waits) return TileRanges{0u, total, 0u, 0u, 0u, 0u, 0u};                          // everything, natural order
    // everything in one launch, the halo-reading tiles last; the CTA awaits the arrival flags when it reaches them
    return TileRanges{lo, total - lo - hi, 0u, lo, total - hi, hi, total - lo - hi};
}

struct Ops {
    int (*dek1)(const Geo&, const StepHost&, const void* mean_in, const void* sc_in, void* mean_out, void* sc_out,
                void* err_out, void* ref_out, const void* halo_lo, const void* halo_hi, const Workspace&,
                double* out_sum, cudaStream_t);
    int (*ek0_a)(const Geo&, const StepHost&, const void* mean_in, const void* halo_lo, const void* halo_hi,
                 const Workspace&, double* out_sum, cudaStream_t);
    int (*ek0_mid)(const StepHost&, const void* Cl_in, const double* z_sq_sum, double d_global, void* Cl_out,
                   void* err_out, const Workspace&, cudaStream_t);
    int (*ek0_b)(const Geo&, const StepHost&, const void* mean_in, void* mean_out, void* ref_out, const void* halo_lo,
                 const void* halo_hi, const Workspace&, double* out_sum, cudaStream_t);
    // the same attempts driven by the device-resident step controller, up to ``max_attempts`` of them in ONE cooperative
    // launch: in / out = ring[ctl->cur] / ring[(ctl->cur + 1) % nbuf]
    int (*dek1_ctl)(const Geo&, const StepHost&, void* const* mean_ring, void* const* sc_ring, int nbuf, int max_attempts,
                    void* const* err_ring, void* const* ref_ring, const void* halo_lo, const void* halo_hi, const Workspace&, StepCtl* ctl,
                    double d_global, double* out_sum, cudaStream_t);
    int (*ek0_ctl)(const Geo&, const StepHost&, void* const* mean_ring, void* const* Cl_ring, int nbuf, int max_attempts,
                   void* const* err_ring, void* const* ref_ring, const void* halo_lo, const void* halo_hi, const Workspace&,
                   const CommReduce& reduce_b, unsigned long long* mid_flag, unsigned long long mid_seq, StepCtl* ctl,
                   double d_global, double* z_sq_out, double* out_sum, cudaStream_t);
    // whole KroneckerEK0 attempt in one cooperative launch (fhn_2d only; returns kNoFusedKernel otherwise)
    int (*ek0_fused)(const Geo&, const StepHost&, const void* mean_in, const void* Cl_in, void* mean_out, void* Cl_out,
                     void* err_out, void* ref_out, const void* halo_lo, const void* halo_hi, const Workspace&,
                     const CommReduce& reduce_b, unsigned long long* mid_flag, unsigned long long mid_seq,
                     double d_global, double* z_sq_out, double* out_sum, cudaStream_t);
    int (*qr_lower)(const void* s1, const void* s2, void* out, long long batch, int s1_broadcast, int s2_broadcast,
                    int rows_given, cudaStream_t);
    int (*pack)(const Geo&, const StepHost&, const void* mean_in, void* edge_first, void* edge_last, int n_first,
                int n_last, cudaStream_t);
    int (*pack_peer)(const Geo&, const StepHost&, const void* mean_in, const PackPeer<double>& pk, cudaStream_t);
};

constexpr int kNoFusedKernel = -2;

struct EvalOps {
    int (*eval_f)(const Geo&, const void* y, void* fy, void* jd, const void* halo_lo, const void* halo_hi,
                  cudaStream_t);
};

const Ops* find_ops(int dtype, int n);         // nullptr if not instantiated
const EvalOps* find_eval_ops(int dtype);
long long grid_for(const Geo& g);              // CTAs a step kernel launches for this shard

template <typename T, int N>
inline Consts<T, N> make_consts(const StepHost& h) {
    Consts<T, N> c;
    for (int k = 0; k < N; ++k) {
        c.p[k] = (T)h.p[k];
        c.pinv[k] = (T)h.pinv[k];
    }
    for (int k = 0; k < N * (N + 1) / 2; ++k) c.ql[k] = (T)h.ql[k];
    c.dt = (T)h.dt;
    c.abstol = (T)h.abstol;
    c.reltol = (T)h.reltol;
    c.want_norm = h.want_norm;
    c.flags = h.flags;
    return c;
}

#define TDX_CUDA_TRY(expr)                        \
    do {                                          \
        cudaError_t e__ = (expr);                 \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

// grid of a persistent kernel: as many CTAs as fit on the device at once, at most one per tile.
// The occupancy query and the shared-memory opt-in are done once per (kernel instantiation, device).
template <auto kern>
inline cudaError_t persistent_grid(size_t smem, long long tiles, int reserve, unsigned& grid, int threads = kThreads) {
    constexpr int kMaxDevices = 64;
    static int resident_cache[kMaxDevices] = {};     // one static per kernel FUNCTION (non-type template parameter)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int resident = (dev >= 0 && dev < kMaxDevices) ? resident_cache[dev] : 0;
    if (resident == 0) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem)) != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        resident = sms * per_sm;
        if (dev >= 0 && dev < kMaxDevices) resident_cache[dev] = resident;
    }
    if (reserve > 0 && resident - reserve >= 1) resident -= reserve;
    if (const char* cap = std::getenv("TDX_MAX_CTAS")) {     // tests: the sums must not depend on the number of CTAs
        const int m = std::atoi(cap);
        if (m >= 1 && m < resident) resident = m;
    }
    grid = (unsigned)(tiles < resident ? tiles : resident);
    return cudaSuccess;
}

// dispatch a templated launch body over the vector field
#define TDX_SWITCH_IVP(kind, T, ...)                                             \
    switch (kind) {                                                              \
        case IVP_VDP: { using IVP = VanderPol<T>; __VA_ARGS__ } break;           \
        case IVP_LORENZ96: { using IVP = Lorenz96<T>; __VA_ARGS__ } break;       \
        case IVP_BRUSS: { using IVP = Brusselator<T>; __VA_ARGS__ } break;       \
        case IVP_FHN2D: { using IVP = Fhn2d<T>; __VA_ARGS__ } break;             \
        case IVP_BURGERS1D: { using IVP = Burgers1d<T>; __VA_ARGS__ } break;     \
        case IVP_WAVE1D: { using IVP = Wave1d<T>; __VA_ARGS__ } break;           \
        case IVP_THREEBODY: { using IVP = ThreeBody<T>; __VA_ARGS__ } break;     \
        case IVP_PLEIADES: { using IVP = Pleiades<T>; __VA_ARGS__ } break;       \
        case IVP_EXTERNAL: { using IVP = External<T>; __VA_ARGS__ } break;       \
        default: return -1;                                                      \
    }

inline bool force_comm_kernel() {
    static const bool on = std::getenv("TDX_FORCE_COMM_KERNEL") != nullptr;   // experiments only
    return on;
}
inline bool force_cta_partials() {
    static const bool on = std::getenv("TDX_CTA_PARTIALS") != nullptr;        // A/B: per-CTA partial sums instead of per-tile ones
    return on;
}

template <typename T, int N>
struct Launch {
    using L = SmemLayout<T, N>;

    static int dek1_ctl(const Geo& g, const StepHost& h, void* const* mean_ring, void* const* sc_ring, int nbuf, int max_attempts,
                        void* const* err_ring, void* const* ref_ring, const void* halo_lo, const void* halo_hi, const Workspace& ws,
                        StepCtl* ctl, double d_global, double* out_sum, cudaStream_t st) {
        return dek1_impl(g, h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, halo_lo, halo_hi, ws, out_sum, st, ctl,
                         mean_ring, sc_ring, err_ring, ref_ring, nbuf, max_attempts, d_global);
    }

    static int dek1(const Geo& g, const StepHost& h, const void* mean_in, const void* sc_in, void* mean_out,
                    void* sc_out, void* err_out, void* ref_out, const void* halo_lo, const void* halo_hi,
                    const Workspace& ws, double* out_sum, cudaStream_t st) {
        return dek1_impl(g, h, mean_in, sc_in, mean_out, sc_out, err_out, ref_out, halo_lo, halo_hi, ws, out_sum, st, nullptr,
                         nullptr, nullptr, nullptr, nullptr, 0, 1, 0.0);
    }

    static int dek1_impl(const Geo& g, const StepHost& h, const void* mean_in, const void* sc_in, void* mean_out,
                         void* sc_out, void* err_out, void* ref_out, const void* halo_lo, const void* halo_hi,
                         const Workspace& ws, double* out_sum, cudaStream_t st, StepCtl* ctl, void* const* mean_ring,
                         void* const* sc_ring, void* const* err_ring, void* const* ref_ring, int nbuf, int max_attempts,
                         double d_global) {
        DekArgs<T, N> a;
        a.ctl = ctl;
        a.d_global = d_global;
        a.max_attempts = max_attempts;
        a.max_accepted = ws.max_accepted;
        for (int i = 0; i < kMaxRing; ++i) {
            a.mean_ring[i] = (ctl && i < nbuf) ? (T*)mean_ring[i] : nullptr;
            a.sc_ring[i] = (ctl && i < nbuf) ? (T*)sc_ring[i] : nullptr;
            a.err_ring[i] = (ctl && i < nbuf && err_ring) ? (T*)err_ring[i] : nullptr;
            a.ref_ring[i] = (ctl && i < nbuf && ref_ring) ? (T*)ref_ring[i] : nullptr;
        }
        a.g = g;
        a.c = make_consts<T, N>(h);
        a.mean_in = (const T*)mean_in;
        a.sc_in = (const T*)sc_in;
        a.mean_out = (T*)mean_out;
        a.sc_out = (T*)sc_out;
        a.err_out = (T*)err_out;
        a.ref_out = (T*)ref_out;
        a.halo_lo = (const T*)halo_lo;
        a.halo_hi = (const T*)halo_hi;
        a.halo_parity_stride = ws.halo_parity_stride_bytes / (long long)sizeof(T);
        a.partials = ws.partials;
        a.ticket = ws.ticket;
        a.out_sum = out_sum;
        const bool comm = ws.pack.enabled || ws.wait.flag_lo || ws.wait.flag_hi || ws.reduce.world > 0 || force_comm_kernel();
        TDX_SWITCH_IVP(g.kind, T, {
            using Shape = typename IVP::Shape;
            const size_t smem = DekSmem<T, N>::total(ExtTile<IVP>::ELEMS);
            a.tiles_x = (unsigned)((g.cols + Shape::TC - 1) / Shape::TC);
            const unsigned tiles_y = (unsigned)((g.rows + Shape::TR - 1) / Shape::TR);
            a.ranges = tile_ranges(h.flags, a.tiles_x, tiles_y, ws.has_lo, ws.has_hi, g.kind == IVP_FHN2D,
                                   ws.wait.flag_lo != nullptr || ws.wait.flag_hi != nullptr);
            a.num_tiles = a.ranges.n0 + a.ranges.n1 + a.ranges.n2;
            a.wait = ws.wait;
            a.reduce = ws.reduce;
            // a launch over all tiles leaves one partial per tile (sharding-independent summation); a launch over a part of
            // them (interior / rim) falls back to per-CTA partials
            const bool whole = a.num_tiles == a.tiles_x * tiles_y && (long long)a.num_tiles <= ws.unit_cap &&
                               !(force_cta_partials() && !ctl && ws.reduce.world == 0);
            a.red = whole ? tile_red_plan<Shape>(g, ws.unit_a, a.tiles_x, tiles_y) : RedPlan{};
            // tiles beyond a CTA's first two are handed out through a counter (the last CTA of the attempt's rendezvous resets
            // it): a CTA that falls behind its SM mate simply takes fewer tiles.  Needs the rendezvous, i.e. a norm or a controller.
            static const bool static_tiles = std::getenv("TDX_STATIC_TILES") != nullptr;      // A/B experiments
            a.tile_counter = (whole && a.red.unit != nullptr && (h.want_norm || ctl) && !static_tiles) ? ws.tile_counter : nullptr;
            if (a.num_tiles == 0u) {
                if (h.want_norm) TDX_CUDA_TRY(cudaMemsetAsync(out_sum, 0, sizeof(double), st));
                return 0;
            }
            a.pack = typed_pack(ws.pack);
            unsigned grid = 0;
            if (ctl) {
                if (!whole) return (int)cudaErrorInvalidValue;
                // all CTAs wait for one another between the attempts: a cooperative launch guarantees that they are resident
                void* params[] = {(void*)&a};
                if (comm) {
                    constexpr auto kern = dek1_kernel<IVP, T, N, true, true>;
                    TDX_CUDA_TRY(persistent_grid<kern>(smem, a.num_tiles, ws.reserve_ctas, grid));
                    TDX_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), params, smem, st));
                } else {
                    constexpr auto kern = dek1_kernel<IVP, T, N, false, true>;      // one GPU: none of the peer-memory code paths
                    TDX_CUDA_TRY(persistent_grid<kern>(smem, a.num_tiles, ws.reserve_ctas, grid));
                    TDX_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), params, smem, st));
                }
            } else if (comm) {
                constexpr auto kern = dek1_kernel<IVP, T, N, true>;
                TDX_CUDA_TRY(persistent_grid<kern>(smem, a.num_tiles, ws.reserve_ctas, grid));
                kern<<<grid, kThreads, smem, st>>>(a);
            } else {
                constexpr auto kern = dek1_kernel<IVP, T, N, false>;
                TDX_CUDA_TRY(persistent_grid<kern>(smem, a.num_tiles, ws.reserve_ctas, grid));
                kern<<<grid, kThreads, smem, st>>>(a);
            }
        })
        return (int)cudaGetLastError();
    }

    static Ek0Args<T, N> ek0_args(const Geo& g, const StepHost& h, const void* mean_in, void* mean_out, void* ref_out,
                                  const void* halo_lo, const void* halo_hi, const Workspace& ws, double* out_sum) {
        Ek0Args<T, N> a;
        a.g = g;
        a.c = make_consts<T, N>(h);
        a.mean_in = (const T*)mean_in;
        a.mean_out = (T*)mean_out;
        a.ref_out = (T*)ref_out;
        a.halo_lo = (const T*)halo_lo;
        a.halo_hi = (const T*)halo_hi;
        a.halo_parity_stride = ws.halo_parity_stride_bytes / (long long)sizeof(T);
        a.mid = (const Ek0Mid<T, N>*)ws.mid;
        a.partials = ws.partials;
        a.ticket = ws.ticket;
        a.out_sum = out_sum;
        return a;
    }

    template <bool kUpdate>
    static int ek0_sweep(const Geo& g, Ek0Args<T, N>& a, unsigned flags, const Workspace& ws, cudaStream_t st) {
        TDX_SWITCH_IVP(g.kind, T, {
            using Shape = typename IVP::Shape;
            const size_t smem = L::BAR_BYTES + L::RED_BYTES + 2 * L::at_bytes(ExtTile<IVP>::ELEMS);
            constexpr auto kern = ek0_sweep_kernel<IVP, T, N, kUpdate>;
            a.tiles_x = (unsigned)((g.cols + Shape::TC - 1) / Shape::TC);
            const unsigned tiles_y = (unsigned)((g.rows + Shape::TR - 1) / Shape::TR);
            a.ranges = tile_ranges(flags, a.tiles_x, tiles_y, ws.has_lo, ws.has_hi, g.kind == IVP_FHN2D,
                                   ws.wait.flag_lo != nullptr || ws.wait.flag_hi != nullptr);
            a.num_tiles = a.ranges.n0 + a.ranges.n1 + a.ranges.n2;
            a.wait = ws.wait;
            a.reduce = ws.reduce;
            a.reverse = kUpdate ? 1 : 0;
            const bool whole = a.num_tiles == a.tiles_x * tiles_y && (long long)a.num_tiles <= ws.unit_cap;
            a.red = whole ? tile_red_plan<Shape>(g, kUpdate ? ws.unit_b : ws.unit_a, a.tiles_x, tiles_y) : RedPlan{};
            if (a.num_tiles == 0u) {
                if (a.out_sum && (!kUpdate || a.c.want_norm)) TDX_CUDA_TRY(cudaMemsetAsync(a.out_sum, 0, sizeof(double), st));
                return 0;
            }
            unsigned grid = 0;
            TDX_CUDA_TRY(persistent_grid<kern>(smem, a.num_tiles, ws.reserve_ctas, grid));
            a.pack = typed_pack(ws.pack);
            if (kUpdate) a.pack.enabled = 0;
            kern<<<grid, kThreads, smem, st>>>(a);
        })
        return (int)cudaGetLastError();
    }

    static int ek0_ctl(const Geo& g, const StepHost& h, void* const* mean_ring, void* const* Cl_ring, int nbuf, int max_attempts,
                       void* const* err_ring, void* const* ref_ring, const void* halo_lo, const void* halo_hi, const Workspace& ws,
                       const CommReduce& reduce_b, unsigned long long* mid_flag, unsigned long long mid_seq, StepCtl* ctl,
                       double d_global, double* z_sq_out, double* out_sum, cudaStream_t st) {
        return ek0_fused_impl<true>(g, h, mean_ring, Cl_ring, nbuf, max_attempts, err_ring, ref_ring, halo_lo, halo_hi, ws, reduce_b,
                                    mid_flag, mid_seq, d_global, z_sq_out, out_sum, st, ctl);
    }

    static int ek0_fused(const Geo& g, const StepHost& h, const void* mean_in, const void* Cl_in, void* mean_out,
                         void* Cl_out, void* err_out, void* ref_out, const void* halo_lo, const void* halo_hi,
                         const Workspace& ws, const CommReduce& reduce_b, unsigned long long* mid_flag,
                         unsigned long long mid_seq, double d_global, double* z_sq_out, double* out_sum,
                         cudaStream_t st) {
        void* mean_ring[2] = {const_cast<void*>(mean_in), mean_out};
        void* Cl_ring[2] = {const_cast<void*>(Cl_in), Cl_out};
        void* err_ring[2] = {nullptr, err_out};
        void* ref_ring[2] = {nullptr, ref_out};
        return ek0_fused_impl<false>(g, h, mean_ring, Cl_ring, 2, 1, err_ring, ref_ring, halo_lo, halo_hi, ws, reduce_b,
                                     mid_flag, mid_seq, d_global, z_sq_out, out_sum, st, nullptr);
    }

    template <class IVP, bool kCtl>
    static int launch_chain(const Geo& g, Ek0FusedArgs<T, N>& a, const Workspace& ws, cudaStream_t st) {
        using Shape = typename IVP::Shape;
        using CCfg = Ek0ChainCfg<T, N>;
        constexpr auto kern = ek0_fused_chain_kernel<IVP, T, N, kCtl>;
        const size_t smem = CCfg::total(ExtTile<IVP>::ELEMS);
        a.col_blocks = (unsigned)((g.cols + Shape::TC - 1) / Shape::TC);                       // tiles per tile row
        const unsigned tiles_y = (unsigned)((g.rows + Shape::TR - 1) / Shape::TR);
        a.total_strips = a.col_blocks * tiles_y;                                               // tiles
        if ((long long)a.total_strips > ws.unit_cap) return (int)cudaErrorInvalidValue;
        a.red_a = tile_red_plan<Shape>(g, ws.unit_a, a.col_blocks, tiles_y);
        a.red_b = tile_red_plan<Shape>(g, ws.unit_b, a.col_blocks, tiles_y);
        unsigned grid = 0;
        TDX_CUDA_TRY(persistent_grid<kern>(smem, (long long)a.total_strips, 0, grid, CCfg::THREADS));
        void* params[] = {(void*)&a};
        TDX_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(CCfg::THREADS), params, smem, st));
        return (int)cudaGetLastError();
    }

    template <bool kCtl>
    static int ek0_fused_impl(const Geo& g, const StepHost& h, void* const* mean_ring, void* const* Cl_ring, int nbuf,
                              int max_attempts, void* const* err_ring, void* const* ref_ring, const void* halo_lo, const void* halo_hi,
                              const Workspace& ws, const CommReduce& reduce_b, unsigned long long* mid_flag,
                              unsigned long long mid_seq, double d_global, double* z_sq_out, double* out_sum,
                              cudaStream_t st, StepCtl* ctl) {
        constexpr int W = TDX_EK0_W;
        using Cfg = Ek0FusedCfg<T, N, W>;
        Ek0FusedArgs<T, N> a;
        a.ctl = ctl;
        a.max_attempts = max_attempts;
        a.max_accepted = ws.max_accepted;
        bool aligned16 = true;
        for (int i = 0; i < kMaxRing; ++i) {
            a.mean_ring[i] = i < nbuf ? (T*)mean_ring[i] : nullptr;
            a.Cl_ring[i] = i < nbuf ? (T*)Cl_ring[i] : nullptr;
            a.err_ring[i] = (i < nbuf && err_ring) ? (T*)err_ring[i] : nullptr;
            a.ref_ring[i] = (i < nbuf && ref_ring) ? (T*)ref_ring[i] : nullptr;
            if (i < nbuf && (reinterpret_cast<uintptr_t>(mean_ring[i]) & 15) != 0) aligned16 = false;
        }
        a.g = g;
        a.c = make_consts<T, N>(h);
        a.cd = make_consts<double, N>(h);
        a.halo_lo = (const T*)halo_lo;
        a.halo_hi = (const T*)halo_hi;
        a.halo_parity_stride = ws.halo_parity_stride_bytes / (long long)sizeof(T);
        a.mid = (Ek0Mid<T, N>*)ws.mid;
        a.mid_flag = mid_flag;
        a.mid_seq = mid_seq;
        a.cl_flag = ws.cl_flag;
        a.cl_seq = ws.cl_seq;
        a.partials = ws.partials;
        a.ticket = ws.ticket;
        a.z_sq_out = z_sq_out;
        a.out_sum = out_sum;
        a.d_global = d_global;
        a.wait = ws.wait;
        a.reduce_a = ws.reduce;
        a.reduce_b = reduce_b;
        a.pack = typed_pack(ws.pack);
        if (g.kind != IVP_FHN2D) {
            // 1-d vector fields: the tile-based one-launch kernel
            a.chunks = 0;
            a.keep_frac = 2.0f;
            a.use_bulk = (g.cols % (16 / (int)sizeof(T)) == 0) && aligned16 ? 1 : 0;
            switch (g.kind) {
                case IVP_VDP: return launch_chain<VanderPol<T>, kCtl>(g, a, ws, st);
                case IVP_LORENZ96: return launch_chain<Lorenz96<T>, kCtl>(g, a, ws, st);
                case IVP_BRUSS: return launch_chain<Brusselator<T>, kCtl>(g, a, ws, st);
                case IVP_BURGERS1D: return launch_chain<Burgers1d<T>, kCtl>(g, a, ws, st);
                case IVP_WAVE1D: return launch_chain<Wave1d<T>, kCtl>(g, a, ws, st);
                case IVP_THREEBODY: return launch_chain<ThreeBody<T>, kCtl>(g, a, ws, st);
                case IVP_PLEIADES: return launch_chain<Pleiades<T>, kCtl>(g, a, ws, st);
                case IVP_EXTERNAL: return launch_chain<External<T>, kCtl>(g, a, ws, st);
                default: return kNoFusedKernel;
            }
        }
        constexpr int PADC = Cfg::PADC;
        a.use_bulk = (g.cols % PADC == 0) && aligned16 ? 1 : 0;
        a.col_blocks = (unsigned)((g.cols + W - 1) / W);
        // strips are whole row blocks (8 grid rows, or single rows on small grids; aligned in global row numbers)
        static const int forced_rb = std::getenv("TDX_EK0_RB") ? std::atoi(std::getenv("TDX_EK0_RB")) : 0;   // A/B of the unit size (one GPU)
        const int rb = forced_rb ? forced_rb : ek0_row_block(g.grows, (int)a.col_blocks);
        if (rb == 8) return launch_fhn<kCtl, 8>(g, a, ws, st);
        return launch_fhn<kCtl, 1>(g, a, ws, st);
    }

    template <bool kCtl, int RB>
    static int launch_fhn(const Geo& g, Ek0FusedArgs<T, N>& a, const Workspace& ws, cudaStream_t st) {
        constexpr int W = TDX_EK0_W;
        using Cfg = Ek0FusedCfg<T, N, W>;
        constexpr auto kern = ek0_fused_fhn_kernel<T, N, W, kCtl, RB>;
        unsigned resident = 0;
        TDX_CUDA_TRY(persistent_grid<kern>(Cfg::total(), 1LL << 40, 0, resident, Cfg::THREADS));
        a.row_block = RB;
        a.block0 = (int)(g.grow0 / RB);
        a.nblocks = (int)((g.grow0 + g.rows + RB - 1) / RB) - a.block0;
        unsigned chunks = resident / a.col_blocks;
        if (chunks < 1u) chunks = 1u;
        if (chunks > (unsigned)a.nblocks) chunks = (unsigned)a.nblocks;
        a.chunks = chunks;
        a.total_strips = a.col_blocks * chunks;
        const unsigned grid = a.total_strips < resident ? a.total_strips : resident;
        // one unit partial per (row block, column block) and phase
        const long long units = (long long)a.nblocks * a.col_blocks;
        if (units > ws.unit_cap) return (int)cudaErrorInvalidValue;
        a.red_a = RedPlan{ws.unit_a, (unsigned)units, (unsigned long long)a.block0 * a.col_blocks,
                          (unsigned long long)((g.grows + RB - 1) / RB) * a.col_blocks};
        a.red_b = a.red_a;
        a.red_b.unit = ws.unit_b;
        // L2 residency hints: keep as many rows of every strip as fit the budget; everything fits -> no hints at all
        {
            static const double budget_mb = std::getenv("TDX_EK0_L2_KEEP_MB") ? std::atof(std::getenv("TDX_EK0_L2_KEEP_MB")) : 80.0;
            const double mean_mb = (double)g.d * N * sizeof(T) / 1.0e6;
            a.keep_frac = (budget_mb <= 0.0) ? 2.0f : (float)(budget_mb / mean_mb);
        }
        // the CTAs wait for one another at the in-kernel barriers: a cooperative launch guarantees that all of them are resident
        // (measured: no launch-latency penalty against a plain launch of the same grid)
        void* params[] = {(void*)&a};
        TDX_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(Cfg::THREADS), params, Cfg::total(), st));
        return (int)cudaGetLastError();
    }

    static int ek0_a(const Geo& g, const StepHost& h, const void* mean_in, const void* halo_lo, const void* halo_hi,
                     const Workspace& ws, double* out_sum, cudaStream_t st) {
        Ek0Args<T, N> a = ek0_args(g, h, mean_in, nullptr, nullptr, halo_lo, halo_hi, ws, out_sum);
        return ek0_sweep<false>(g, a, h.flags, ws, st);
    }

    static int ek0_b(const Geo& g, const StepHost& h, const void* mean_in, void* mean_out, void* ref_out,
                     const void* halo_lo, const void* halo_hi, const Workspace& ws, double* out_sum, cudaStream_t st) {
        Ek0Args<T, N> a = ek0_args(g, h, mean_in, mean_out, ref_out, halo_lo, halo_hi, ws, out_sum);
        return ek0_sweep<true>(g, a, 0u, ws, st);
    }

    static int ek0_mid(const StepHost& h, const void* Cl_in, const double* z_sq_sum, double d_global, void* Cl_out,
                       void* err_out, const Workspace& ws, cudaStream_t st) {
        Ek0MidArgs<T, N> a;
        a.c = make_consts<double, N>(h);
        a.Cl_in = (const T*)Cl_in;
        a.z_sq_sum = z_sq_sum;
        a.d_global = d_global;
        a.Cl_out = (T*)Cl_out;
        a.err_out = (T*)err_out;
        a.mid = (Ek0Mid<T, N>*)ws.mid;
        ek0_mid_kernel<T, N><<<1, 32, 0, st>>>(a);
        return (int)cudaGetLastError();
    }

    static int qr_lower(const void* s1, const void* s2, void* out, long long batch, int s1_broadcast, int s2_broadcast,
                        int rows_given, cudaStream_t st) {
        QrArgs<T> a{(const T*)s1, (const T*)s2, (T*)out, batch, s1_broadcast, s2_broadcast, rows_given};
        if (batch <= 0) return 0;
        const unsigned blocks = (unsigned)((batch + 127) / 128);
        if (s2) batched_qr_lower_kernel<T, N, 2 * N><<<blocks, 128, 0, st>>>(a);
        else batched_qr_lower_kernel<T, N, N><<<blocks, 128, 0, st>>>(a);
        return (int)cudaGetLastError();
    }

    static int pack(const Geo& g, const StepHost& h, const void* mean_in, void* edge_first, void* edge_last,
                    int n_first, int n_last, cudaStream_t st) {
        PackArgs<T, N> a;
        a.g = g;
        a.c = make_consts<T, N>(h);
        a.mean_in = (const T*)mean_in;
        a.edge_first = (T*)edge_first;
        a.edge_last = (T*)edge_last;
        a.n_first = n_first;
        a.n_last = n_last;
        const int total = g.nf * (n_first + n_last);
        if (total <= 0) return 0;
        const int blocks = (total + 255) / 256;
        pack_halo_kernel<T, N><<<blocks, 256, 0, st>>>(a);
        return (int)cudaGetLastError();
    }

    static PackPeer<T> typed_pack(const PackPeer<double>& pk) {
        PackPeer<T> out;
        out.dst_lower = (T*)pk.dst_lower;
        out.dst_upper = (T*)pk.dst_upper;
        out.flag_lower = pk.flag_lower;
        out.flag_upper = pk.flag_upper;
        out.done_count = pk.done_count;
        out.seq = pk.seq;
        out.parity_stride = pk.parity_stride / (long long)sizeof(T);   // the host sets it in bytes
        out.n_first = pk.n_first;
        out.n_last = pk.n_last;
        out.enabled = pk.enabled;
        return out;
    }

    static int pack_peer(const Geo& g, const StepHost& h, const void* mean_in, const PackPeer<double>& pk,
                         cudaStream_t st) {
        PackPeerArgs<T, N> a;
        a.g = g;
        a.c = make_consts<T, N>(h);
        a.mean_in = (const T*)mean_in;
        a.pk = typed_pack(pk);
        pack_halo_peer_kernel<T, N><<<2 * kPackCtas, 1024, 0, st>>>(a);
        return (int)cudaGetLastError();
    }

    static Ops ops() {
        Ops o{};
        o.dek1 = &dek1;
        o.ek0_a = &ek0_a;
        o.ek0_mid = &ek0_mid;
        o.ek0_b = &ek0_b;
        o.ek0_fused = &ek0_fused;
        o.qr_lower = &qr_lower;
        o.dek1_ctl = &dek1_ctl;
        o.ek0_ctl = &ek0_ctl;
        o.pack = &pack;
        o.pack_peer = &pack_peer;
        return o;
    }
};

}  // namespace tdx

#define TDX_DEFINE_OPS(T, N, NAME)                 \
    namespace tdx {                                \
    const Ops* ops_##NAME() {                      \
        static const Ops o = Launch<T, N>::ops();  \
        return &o;                                 \
    }                                              \
    }
