// tornadox_b200 kernels: fused DiagonalEK1 step, KroneckerEK0 sweeps, halo packing, f evaluation.
// One thread per state coordinate, everything in registers; AoS factor blocks are streamed through
// shared memory with TMA 1-D bulk copies (cp.async.bulk), one 32-block segment per warp.
#pragma once

#include "tdx_device.cuh"

namespace tdx {

#ifndef TDX_MIN_BLOCKS_N2
#define TDX_MIN_BLOCKS_N2 4
#endif
#ifndef TDX_MIN_BLOCKS_N3
#define TDX_MIN_BLOCKS_N3 4
#endif
#ifndef TDX_MIN_BLOCKS_N4
#define TDX_MIN_BLOCKS_N4 2
#endif
#ifndef TDX_MIN_BLOCKS
#define TDX_MIN_BLOCKS 2
#endif
#ifndef TDX_PREFETCH_MEAN
#define TDX_PREFETCH_MEAN 0   // 1: hold the next tile's mean column in registers during the QR (spills at 128 regs)
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
    // One (n, n) block per lane.  Odd n: the block stride is n^2 elements (odd => conflict-free element accesses).
    // Even n: a block is a whole number of 16-byte units; it is accessed with 128-bit loads / stores only and its stride is
    // padded to an ODD number of units (=> conflict-free within every quarter-warp phase of a 128-bit access).
#ifdef TDX_PAD_ELEMENTWISE
    static constexpr bool VEC = false;
#else
    static constexpr bool VEC = (N % 2 == 0);
#endif
    static constexpr int PER_UNIT = 16 / (int)sizeof(T);            // elements per 16-byte unit
    static constexpr int UNITS = NN / PER_UNIT;                     // units per block (VEC only)
    static constexpr int STRIDE = VEC ? ((UNITS | 1) * PER_UNIT) : ((NN % 2 == 1) ? NN : NN + 1);
    static constexpr int SC_ELEMS = kWarps * 32 * STRIDE;
    static constexpr size_t BAR_BYTES = 128;                       // kWarps mbarriers, padded
    static constexpr size_t RED_BYTES = 256;                       // 2 x kWarps per-tile warp sums / the chunk sums of the final reduction
    __host__ __device__ static constexpr size_t sc_bytes() { return ((size_t)SC_ELEMS * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t at_bytes(int at_elems) { return ((size_t)at_elems * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t total(int at_elems, bool with_sc) {
        return BAR_BYTES + RED_BYTES + at_bytes(at_elems) + (with_sc ? sc_bytes() : 0);
    }
};

// 128-bit shared-memory access to a lane's block (even n, see SmemLayout): the whole block to / from registers
template <typename T, int NN>
__device__ __forceinline__ void block_load_vec(const T* blk, T (&v)[NN]) {
    constexpr int PER = 16 / (int)sizeof(T);
    const uint4* src = reinterpret_cast<const uint4*>(blk);
#pragma unroll
    for (int u = 0; u < NN / PER; ++u) {
        const uint4 q = src[u];
        if constexpr (sizeof(T) == 8) {
            v[2 * u] = (T)__hiloint2double((int)q.y, (int)q.x);
            v[2 * u + 1] = (T)__hiloint2double((int)q.w, (int)q.z);
        } else {
            v[4 * u] = (T)__uint_as_float(q.x);
            v[4 * u + 1] = (T)__uint_as_float(q.y);
            v[4 * u + 2] = (T)__uint_as_float(q.z);
            v[4 * u + 3] = (T)__uint_as_float(q.w);
        }
    }
}
template <typename T, int NN>
__device__ __forceinline__ void block_store_vec(T* blk, const T (&v)[NN]) {
    constexpr int PER = 16 / (int)sizeof(T);
    uint4* dst = reinterpret_cast<uint4*>(blk);
#pragma unroll
    for (int u = 0; u < NN / PER; ++u) {
        uint4 q;
        if constexpr (sizeof(T) == 8) {
            q.x = (unsigned)__double2loint((double)v[2 * u]);
            q.y = (unsigned)__double2hiint((double)v[2 * u]);
            q.z = (unsigned)__double2loint((double)v[2 * u + 1]);
            q.w = (unsigned)__double2hiint((double)v[2 * u + 1]);
        } else {
            q.x = __float_as_uint((float)v[4 * u]);
            q.y = __float_as_uint((float)v[4 * u + 1]);
            q.z = __float_as_uint((float)v[4 * u + 2]);
            q.w = __float_as_uint((float)v[4 * u + 3]);
        }
        dst[u] = q;
    }
}

// A launch processes the tiles [b0, b0 + n0) followed by [b1, b1 + n1) (interior / rim split for multi-GPU overlap).
// Tiles from virtual index ``wait_from`` on read the halo buffers: a CTA waits for the arrival flags when it gets there.
struct TileRanges {
    unsigned int b0, n0, b1, n1, b2, n2;
    unsigned int wait_from;
    __device__ __forceinline__ unsigned tile(unsigned v) const {
        return v < n0 ? b0 + v : (v < n0 + n1 ? b1 + (v - n0) : b2 + (v - n0 - n1));
    }
};

// ---------------------------------------------------------------------------------------------
// front end shared by all step kernels, in two halves so that a persistent kernel can issue the loads of
// the NEXT tile before it computes the current one.
//   front_load : this thread's mean column + the out-of-tile stencil neighbours (global / L2 loads)
//   front_eval : predict, publish the predicted state to the tile, barrier, evaluate f and diag(J)
// ek1.py:232,262-265,302-312 / ek0.py:161-166,142-150
// ---------------------------------------------------------------------------------------------
template <class IVP, typename T, int N>
struct FrontRegs {
    T raw[N];     // this thread's mean column
    T ring;       // the ring point of the extended tile this thread is responsible for (if any)
    T fill;       // threads beyond the shard (ragged tiles) publish what their slot means to the valid neighbours:
                  // the replicated / padded / wrapped value the boundary rule prescribes
};

// tile-local origin (shard-local row/col of the tile's first point) from a thread's coordinates
__device__ __forceinline__ int tile_row0(const Coord& c) { return c.row - c.tr; }
__device__ __forceinline__ int tile_col0(const Coord& c) { return c.col - c.tcol; }

// kOwn = false: only the ring / fill values (the caller gets the thread's own column from somewhere else, e.g. a TMA stage)
// kCoherent: the mean may have been written earlier in the same kernel (multi-attempt kernels): no read-only-path loads
template <class IVP, typename T, int N, bool kOwn = true, bool kCoherent = false, class C = Consts<T, N>>
__device__ __forceinline__ void front_load(const Geo& g, const C& c, const T* __restrict__ mean_in,
                                           const T* halo_lo, const T* halo_hi, const Coord& co,
                                           FrontRegs<IVP, T, N>& fr) {
    using E = ExtTile<IVP>;
    if (kOwn) {
        const T* src = mean_in + co.flat;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            fr.raw[k] = co.valid ? ld_state<kCoherent>(src) : (T)0;
            src += g.d;
        }
    }
    fr.ring = (T)0;
    fr.fill = (T)0;
    MeanAt<T, N, kCoherent, C> at{mean_in, g.d, &c};
    if (E::NRING > 0 && (int)threadIdx.x < IVP::Shape::NF * E::NRING) {
        const int field = threadIdx.x / (E::NRING > 0 ? E::NRING : 1);
        int er, ec;
        E::ring(threadIdx.x - field * E::NRING, er, ec);
        fr.ring = IVP::ring_value(g, tile_row0(co), tile_col0(co), field, er, ec, at, halo_lo, halo_hi);
    }
    if (E::NRING > 0 && !co.valid)
        fr.fill = IVP::ring_value(g, tile_row0(co), tile_col0(co), co.field, co.tr + E::HV, co.tcol + E::HL, at, halo_lo,
                                  halo_hi);
}

template <class IVP, typename T, int N, class Sync = FullSync, class C = Consts<T, N>>
__device__ __forceinline__ void front_eval(const Geo& g, const C& c, const Coord& co,
                                           const FrontRegs<IVP, T, N>& fr, T* s_ext, T (&mp)[N], T& fx, T& jac,
                                           Sync sync = Sync()) {
    using E = ExtTile<IVP>;
    predict_column<T, N, C>(fr.raw, c, mp);
    const T own = co.valid ? c.p[0] * mp[0] : fr.fill;
    s_ext[E::own(co)] = own;
    if (E::NRING > 0 && (int)threadIdx.x < IVP::Shape::NF * E::NRING) {
        const int field = threadIdx.x / (E::NRING > 0 ? E::NRING : 1);
        int er, ec;
        E::ring(threadIdx.x - field * E::NRING, er, ec);
        s_ext[E::index(field, er, ec)] = fr.ring;
    }
    sync();
    fx = (T)0;
    jac = (T)0;
    if (co.valid) IVP::eval(g, co, s_ext, own, fx, jac);
    if (c.flags & 1u) jac = (T)0;   // TDX_STEP_ZERO_JACOBIAN: DiagonalEK0, ek0.py:196-280
}

// ---------------------------------------------------------------------------------------------
// Peer-memory halo delivery: 2 * kPackCtas CTAs (kPackCtas per direction) store the predicted state of the shard's
// first rows/points into the LOWER neighbour's hi-halo buffer and of the last ones into the UPPER neighbour's
// lo-halo buffer; the last CTA of each direction to finish raises that neighbour's arrival flag to ``seq``.
// Runs either as its own launch (pack_halo_peer_kernel) or fused at the start of a step kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kPackCtas = 4;

template <typename T>
struct PackPeer {
    T* dst_lower;                    // lower neighbour's halo_hi buffer of parity 0 (mapped), or null
    T* dst_upper;                    // upper neighbour's halo_lo buffer of parity 0 (mapped), or null
    unsigned long long* flag_lower;  // lower neighbour's halo_flag_hi[0]
    unsigned long long* flag_upper;  // upper neighbour's halo_flag_lo[0]
    unsigned int* done_count;        // local CommHeader::pack_count
    unsigned long long seq;          // halo sequence number of the kernel's first attempt
    long long parity_stride;         // elements between the parity-0 and the parity-1 halo buffer
    int n_first, n_last;
    int enabled;
};

template <typename T, int N, class Sync = FullSync, bool kCoherent = false, class C = Consts<T, N>>
__device__ __forceinline__ void pack_halo_part(const Geo& g, const C& c, const T* mean_in,
                                               const PackPeer<T>& pk, unsigned long long seq, int part_index,
                                               Sync sync = Sync()) {
    MeanAt<T, N, kCoherent, C> at{mean_in, g.d, &c};
    const bool lower = part_index < kPackCtas;
    const int part = lower ? part_index : part_index - kPackCtas;
    const int par = (int)(seq & 1ull);
    T* dst = lower ? pk.dst_lower : pk.dst_upper;
    if (dst == nullptr) return;      // CTA-uniform
    dst += (long long)par * pk.parity_stride;
    const int per_field = lower ? pk.n_first : pk.n_last;
    const long long start = lower ? 0 : (g.npts - pk.n_last);
    const int total = g.nf * per_field;
    const int nthr = sync.threads();
    for (int i = part * nthr + threadIdx.x; i < total; i += kPackCtas * nthr) {
        const int field = i / per_field, j = i - field * per_field;
        dst[i] = at((long long)field * g.npts + start + j);
    }
    sync();
    if (threadIdx.x == 0) {
        // one system-scope fence per CTA: it is cumulative over the stores the barrier above ordered before it
        __threadfence_system();
        unsigned int* cnt = pk.done_count + (lower ? 0 : 1);
        const unsigned int prev = atomicAdd(cnt, 1u);
        if (prev == kPackCtas - 1) {
            *cnt = 0u;
            __threadfence();   // the other CTAs' fences happened before their atomicAdd; order our read of the count
            st_release_sys((lower ? pk.flag_lower : pk.flag_upper) + par, seq);
        }
    }
}

template <typename T, int N>
struct PackPeerArgs {
    Geo g;
    Consts<T, N> c;
    const T* mean_in;
    PackPeer<T> pk;
};

template <typename T, int N>
__global__ void __launch_bounds__(1024) pack_halo_peer_kernel(const __grid_constant__ PackPeerArgs<T, N> a) {
    pack_halo_part<T, N>(a.g, a.c, a.mean_in, a.pk, a.pk.seq, (int)blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// DiagonalEK1 fused attempt: persistent CTAs, two shared-memory slots per warp for the factor blocks.
// While a warp computes tile i out of slot i&1, the TMA engine fills the other slot with tile i+1 and
// drains tile i-1's results; the only CTA-wide barrier per tile is the one that publishes the stencil tile.
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
    const T* halo_lo;                  // the caller's halo buffers, or the parity-0 buffers of the context's peer-written ones
    const T* halo_hi;
    long long halo_parity_stride;      // elements between the parity-0 and parity-1 buffers (0: caller-provided halos)
    double* partials;                  // per-CTA partial sums (launches over a part of the tiles only)
    unsigned int* ticket;
    unsigned int* tile_counter;        // non-null: a CTA's third, fourth, ... tile come from this counter (zero at launch, reset by the last CTA)
    double* out_sum;
    RedPlan red;                       // per-tile partials + the sharding-independent summation order (unit == null: per-CTA partials)
    unsigned int num_tiles, tiles_x;   // num_tiles = number of tiles THIS launch processes (all ranges)
    TileRanges ranges;
    CommWait wait;                     // peer-memory halos: flags to wait for (null: nothing to wait for)
    CommReduce reduce;                 // fused all-reduce of the norm sum (world == 0: off)
    PackPeer<T> pack;                  // fused halo delivery (enabled == 0: off)
    // device-resident step controller (kCtl instantiation): dt, Nordsieck vectors and the current ring slot come from ``ctl``;
    // the kernel runs up to ``max_attempts`` attempts, the last CTA of each publishing the decision to the others
    StepCtl* ctl;
    T* mean_ring[kMaxRing];
    T* sc_ring[kMaxRing];
    T* err_ring[kMaxRing];             // per-slot error_estimate / reference_state vectors (null entries: not written)
    T* ref_ring[kMaxRing];
    int max_attempts;
    long long max_accepted;            // stop the launch after this many newly accepted steps (0: no limit)
    double d_global;
};

template <typename T, int N>
struct DekSmem {
    using L = SmemLayout<T, N>;
    static constexpr int SLOTS = 2;           // stencil tiles: always double-buffered (small)
    // factor blocks: double-buffered while two CTAs per SM still fit (n <= 5 in fp64); beyond that ONE slot per warp keeps
    // two CTAs resident (their phases interleave) instead of one CTA with a prefetch slot (2.6x slower at n = 6)
    // resident CTAs per SM the kernel is compiled for (register cap 64 / 85 / 128 per thread): the small blocks of n <= 4 leave
    // registers for a third and fourth CTA, which is what hides the front end's load latency when there are few factor bytes
    // per coordinate to wait for (8.4 M dimensions, fp64: nu = 1 39 -> 53 %, nu = 2 63 -> 81 %, nu = 3 78 -> 85 % of the roofline)
    static constexpr int MIN_BLOCKS = (N <= 2) ? TDX_MIN_BLOCKS_N2 : (N == 3) ? TDX_MIN_BLOCKS_N3 : (N == 4) ? TDX_MIN_BLOCKS_N4 : TDX_MIN_BLOCKS;
    static constexpr size_t TERM_BYTES = 2 * kThreads * sizeof(double);   // the tile's per-coordinate norm terms, double-buffered
    static constexpr int SC_SLOTS = (MIN_BLOCKS * (L::BAR_BYTES + L::RED_BYTES + TERM_BYTES + 2 * 3328 + 2 * L::sc_bytes() + 1024) <= 225 * 1024) ? 2 : 1;
    __host__ __device__ static constexpr size_t total(int at_elems) {
        return L::BAR_BYTES + L::RED_BYTES + TERM_BYTES + SLOTS * L::at_bytes(at_elems) + SC_SLOTS * L::sc_bytes();
    }
};

// What the last CTA of a controller-driven attempt does once the global norm sum is known: decide (ctl_update), tell the
// host feed, publish the decision (epoch) to the other CTAs.
__device__ inline void ctl_finish_attempt(StepCtl* ctl, double total, double d_global, unsigned long long epoch_next) {
    ctl_update(ctl, total, d_global);
    HostFeed* fd = ctl->feed;
    if (fd != nullptr) {
        // a host that follows the solve: progress goes to mapped pinned memory (the launch itself is bounded by max_accepted,
        // so the ring never wraps onto a state the host has not copied yet)
        __threadfence_system();
        fd->attempts = (unsigned long long)ctl->attempts;
        fd->accepted = (unsigned long long)ctl->accepted;
        if (fd->abort != 0ull) ctl->done = 1;
    }
    __threadfence();
    st_release_gpu(&ctl->epoch, epoch_next);
}

// kComm = false: single-GPU instantiation without any of the peer-memory code paths (halo delivery, flag waits,
// fused all-reduce); the register allocation of this kernel sits at the 128-register cap and is sensitive to them.
// kCtl = true: the step to attempt and the in / out buffers are read from the device-resident controller block, the
// last CTA updates it (accept / reject, next dt), and the kernel goes on with the next attempt (up to max_attempts; launched
// cooperatively so that all CTAs are resident) -- no host involvement between attempts.
template <class IVP, typename T, int N, bool kComm, bool kCtl = false>
__global__ void __launch_bounds__(kThreads, DekSmem<T, N>::MIN_BLOCKS) dek1_kernel(const __grid_constant__ DekArgs<T, N> a) {
    using Shape = typename IVP::Shape;
    using L = SmemLayout<T, N>;
    constexpr int NN = L::NN, STRIDE = L::STRIDE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                       // [2][kWarps]
    double* s_red = reinterpret_cast<double*>(smem_raw + L::BAR_BYTES);          // final reduction scratch
    double* s_term = reinterpret_cast<double*>(smem_raw + L::BAR_BYTES + L::RED_BYTES);   // [2][kThreads] norm terms of a tile
    constexpr size_t kHead = L::BAR_BYTES + L::RED_BYTES + DekSmem<T, N>::TERM_BYTES;
    T* s_at0 = reinterpret_cast<T*>(smem_raw + kHead);                            // [2][AT_ELEMS]
    T* s_sc0 = reinterpret_cast<T*>(smem_raw + kHead + 2 * L::at_bytes(ExtTile<IVP>::ELEMS));
    constexpr size_t AT_STRIDE = L::at_bytes(ExtTile<IVP>::ELEMS) / sizeof(T);
    constexpr size_t SC_STRIDE = L::sc_bytes() / sizeof(T);
    __shared__ int s_flag, s_ok;
    __shared__ unsigned int s_next[2];            // dynamic tile hand-out: the tile after next, fetched one iteration ahead
    __shared__ int s_pending[2];                  // the tile whose 256 terms wait in s_term[slot] for the next barrier (-1: none)
    __shared__ T s_p[N], s_pinv[N], s_dt;         // kCtl: the step's vectors, rebuilt from the controller block before each attempt
    __shared__ int s_cur, s_out;
    __shared__ unsigned long long s_epoch0;
    __shared__ long long s_pause_at;

    const Geo& g = a.g;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T* mean_in = a.mean_in;
    const T* sc_in = a.sc_in;
    T* mean_out = a.mean_out;
    T* sc_out = a.sc_out;
    T* err_out = a.err_out;
    T* ref_out = a.ref_out;

    if (lane == 0) {
        mbar_init(bars + w, 1);
        mbar_init(bars + kWarps + w, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();

    unsigned phase_bits = 0u;          // parity of each slot's mbarrier
    constexpr int SC_SLOTS = DekSmem<T, N>::SC_SLOTS;
    const int max_attempts = kCtl ? a.max_attempts : 1;

#pragma unroll 1
    for (int att = 0; att < max_attempts; ++att) {
        if constexpr (kCtl) {
            if (threadIdx.x == 0) {
                StepCtl* ctl = a.ctl;
                bool ok = true;
                if (att == 0) {
                    s_epoch0 = ld_acquire_gpu(&ctl->epoch);
                    s_pause_at = a.max_accepted > 0 ? ctl->accepted + a.max_accepted : (long long)0x7fffffffffffffffLL;
                } else {
                    ok = spin_until_gpu(&ctl->epoch, s_epoch0 + (unsigned long long)att, (unsigned long long)ctl->wait_timeout_ns);
                }
                if (!ok) {                        // the grid barrier timed out (a CTA or a peer rank is gone): fail the solve
                    ctl->failed = TDX_FAIL_GRID_BARRIER;
                    ctl->done = 1;
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {     // Ql, tolerances and flags come with the launch, the step with the controller
                    s_p[k] = (T)ctl->p[k];
                    s_pinv[k] = (T)ctl->pinv[k];
                }
                s_dt = (T)ctl->dt;
                const int cur = ctl->cur, nbuf = ctl->nbuf;
                s_cur = (ctl->done || !ok || ctl->accepted >= s_pause_at) ? -1 : cur;
                s_out = (cur + 1) % nbuf;
            }
            __syncthreads();
            if (s_cur < 0) break;                 // the solve is over (or a barrier timed out): nothing left to do
            if (att > 0) fence_proxy_async_global();   // TMA loads below read what other CTAs stored in the previous attempt
            mean_in = a.mean_ring[s_cur];
            sc_in = a.sc_ring[s_cur];
            mean_out = a.mean_ring[s_out];
            sc_out = a.sc_ring[s_out];
            err_out = a.err_ring[s_out];
            ref_out = a.ref_ring[s_out];
        }
        decltype(auto) c = [&]() -> decltype(auto) {
            if constexpr (kCtl) return StepView<T, N>{s_p, s_pinv, a.c.ql, s_dt, a.c.abstol, a.c.reltol, a.c.want_norm, a.c.flags};
            else return (a.c);
        }();
        const unsigned long long hseq = a.wait.seq + (unsigned long long)att;          // halo sequence number of this attempt
        // halo buffers of this attempt's parity: formed where they are used (two long-lived pointers less in a kernel that sits
        // at its register cap at every order)
        const bool halo_odd = (hseq & 1ull) != 0ull;

        // issue the streaming load of one tile's 32 factor blocks of this warp into a slot; returns whether the
        // bulk (TMA) path was used (warp-uniform), in which case the data lands on the slot's mbarrier
        auto stage_in = [&](const Coord& co, int slot) -> bool {
            T* dst = s_sc0 + slot * SC_STRIDE + w * 32 * STRIDE;
            const long long off = co.seg_base * NN;
            const bool bulk = (STRIDE == NN) && (co.seg_len == 32) && ((reinterpret_cast<uintptr_t>(sc_in + off) & 15) == 0);
            if (bulk) {
                if (lane == 0) {
                    uint64_t* bar = bars + slot * kWarps + w;
                    mbar_expect_tx(bar, (uint32_t)(32 * NN * sizeof(T)));
                    bulk_g2s(dst, sc_in + off, (uint32_t)(32 * NN * sizeof(T)), bar);
                }
            } else {
                // padded layout: element-wise asynchronous copies (LDGSTS), one group per tile; they land while the current
                // tile is being factorised
                if (L::VEC && (reinterpret_cast<uintptr_t>(sc_in) & 15) == 0) {      // 16-byte units, L1 bypassed
                    const int total = co.seg_len * L::UNITS;
                    for (int u = lane; u < total; u += 32) {
                        const int blk = u / L::UNITS;
                        cp_async_unit(dst + blk * STRIDE + (u - blk * L::UNITS) * L::PER_UNIT, sc_in + off + (long long)u * L::PER_UNIT);
                    }
                } else {
                    const int total = co.seg_len * NN;
                    for (int e = lane; e < total; e += 32) {
                        const int blk = e / NN;
                        cp_async_elem<kCtl>(dst + blk * STRIDE + (e - blk * NN), sc_in + off + e);
                    }
                }
                cp_async_commit();
            }
            return bulk;
        };

        // The lowest block indices deliver this shard's halos before they start on their tiles.  A CTA's first two tiles are
        // fixed ((G-1-b) and (G-1-b)+G in the sharded instantiations, so that the halo-delivering CTAs start with the tiles
        // handed out last); the others come from the tile counter, i.e. a CTA that starts late simply takes fewer of them.
        if (kComm && a.pack.enabled)
            for (int part = (int)blockIdx.x; part < 2 * kPackCtas; part += (int)gridDim.x)   // a small grid takes several parts per CTA
                pack_halo_part<T, N, FullSync, kCtl>(g, c, mean_in, a.pack, a.pack.seq + (unsigned long long)att, part);
        if (threadIdx.x == 0) TDX_STAMP(a.partials, att, 0);
#ifdef TDX_TIMELINE
        if (threadIdx.x == 0) a.partials[5000 + 2 * blockIdx.x] = (double)global_timer_ns();     // this CTA starts on its tiles
#endif
        bool halos_ready = false;          // CTA-uniform: the arrival flags have been seen
        bool poisoned = false;             // CTA-uniform: a halo wait timed out
        unsigned tile = kComm ? (gridDim.x - 1u - blockIdx.x) : blockIdx.x;
        const bool unit_mode = a.red.unit != nullptr;
        const bool dyn_tiles = a.tile_counter != nullptr;
        // launches over a part of the tiles (interior / rim pairs) keep per-CTA partials: the per-thread running sum lives in
        // this thread's own s_term slot, not in two registers of every instantiation
        if (!unit_mode) s_term[threadIdx.x] = 0.0;
        if (threadIdx.x == 0) s_pending[0] = s_pending[1] = -1;      // (read only after a later barrier)
        bool cur_bulk = false;
        if (tile < a.num_tiles) cur_bulk = stage_in(tile_coord<Shape>(g, a.ranges.tile(tile), a.tiles_x), 0);

        // After a barrier that follows the tile's terms: ONE warp (a different one from tile to tile, so that no warp is always
        // the one the next barrier waits for) adds the 256 terms in a fixed order -- lane l the eight warps' terms of lane l in a
        // tree, then the shuffle tree over the lanes.  A fixed function of the tile's coordinates, whoever computes it.
        auto flush_tile_partial = [&](unsigned it_prev) {
            if (w == (int)(it_prev & (unsigned)(kWarps - 1))) {
                const int pending = s_pending[it_prev & 1u];
                if (pending >= 0) {
                    const double* v = s_term + (it_prev & 1u) * kThreads + lane;
                    double x = ((v[0] + v[32]) + (v[64] + v[96])) + ((v[128] + v[160]) + (v[192] + v[224]));
                    x = warp_sum(x);
                    if (lane == 0) {
                        a.red.unit[pending] = x;
                        s_pending[it_prev & 1u] = -1;
                    }
                }
            }
        };
        static_assert(kWarps == 8, "the tile partial adds eight warps' terms");

        unsigned it = 0;
        for (; tile < a.num_tiles; ++it) {
            const int slot = (SC_SLOTS == 2) ? (int)(it & 1u) : 0;
            unsigned fetched = 0u;         // thread 0: the ticket for the tile after next (in flight while this tile is computed)
            if (dyn_tiles && threadIdx.x == 0) fetched = 2u * gridDim.x + atomicAdd(a.tile_counter, 1u);
#ifdef TDX_TIMELINE
            if (threadIdx.x == 0 && (it == 28u || it == 56u || it == 84u))
                a.partials[5600 + 4 * blockIdx.x + (it / 28u - 1u)] = (double)global_timer_ns();
#endif
            T* my_sc = s_sc0 + slot * SC_STRIDE + w * 32 * STRIDE;
            T* s_at = s_at0 + (it & 1u) * AT_STRIDE;
            if (kComm && !halos_ready && tile >= a.ranges.wait_from) {
                if (!cta_wait_halos(a.wait, hseq, &s_ok)) poisoned = true;
                halos_ready = true;
            }

            // (1) mean column, prediction and vector field of the current tile
            const unsigned tile_id = a.ranges.tile(tile);
            const Coord cur = tile_coord<Shape>(g, tile_id, a.tiles_x);
            T mp[N], fx, jac, raw0;
            {
                FrontRegs<IVP, T, N> fr;
                const long long hoff = halo_odd ? a.halo_parity_stride : 0ll;
                front_load<IVP, T, N, true, kCtl>(g, c, mean_in, a.halo_lo ? a.halo_lo + hoff : nullptr,
                                                  a.halo_hi ? a.halo_hi + hoff : nullptr, cur, fr);
                front_eval<IVP, T, N>(g, c, cur, fr, s_at, mp, fx, jac);
                raw0 = fr.raw[0];
            }
            if (unit_mode) flush_tile_partial(it - 1u);   // the barrier inside front_eval ordered the previous tile's warp sums

            // (2) start streaming the next tile's factor blocks into the other slot.  The first two tiles of a CTA are fixed;
            // the others were fetched from the counter by thread 0 one iteration ago (published by the barrier above).
            const unsigned next = (!dyn_tiles || it == 0u) ? tile + gridDim.x : s_next[it & 1u];
            bool next_bulk = false;
            if (SC_SLOTS == 2 && next < a.num_tiles) {
                if (lane == 0) bulk_wait_read0();      // the other slot's previous results have left shared memory
                __syncwarp();
                next_bulk = stage_in(tile_coord<Shape>(g, a.ranges.tile(next), a.tiles_x), slot ^ 1);
            }

            // (3) wait for the current tile's blocks
            if (cur_bulk) {
                mbar_wait(bars + slot * kWarps + w, (phase_bits >> slot) & 1u);
                phase_bits ^= (1u << slot);
            } else if (SC_SLOTS == 2 && next < a.num_tiles && !next_bulk) {
                cp_async_wait_group<1>();  // everything but the next tile's group, which was committed just above
            } else {
                cp_async_wait_group<0>();
            }
            __syncwarp();

            const long long sc_off = cur.seg_base * NN;
            const bool bulk_out = (STRIDE == NN) && (cur.seg_len == 32) &&
                                  ((reinterpret_cast<uintptr_t>(sc_out + sc_off) & 15) == 0);
            double ratio_sq = 0.0;
            if (cur.valid) {
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
                T* blk = my_sc + lane * STRIDE;
                T blkv[L::VEC ? NN : 1];           // even n: the block travels through registers, 128 bits per access
                if constexpr (L::VEC) block_load_vec<T, NN>(blk, blkv);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    T col[N];
#pragma unroll
                    for (int k = 0; k < N; ++k) col[k] = c.pinv[k] * (L::VEC ? blkv[L::VEC ? k * N + j : 0] : blk[k * N + j]);
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
                T new_m0 = (T)0;
                T* mo_ptr = mean_out + cur.flat;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const T l0 = B[r][0];
                    const T l1 = (r >= 1) ? B[r][1] : (T)0;
                    const T kg = (l0 * g0 + l1 * g1) * inv_s2;
                    const T pr = c.p[r];
                    T* row = L::VEC ? (blkv + (L::VEC ? r * N : 0)) : (blk + r * N);
                    row[0] = pr * (l0 - kg * g0);
                    row[1] = pr * (l1 - kg * g1);
#pragma unroll
                    for (int cc = 2; cc < N; ++cc) row[cc] = (r >= cc) ? pr * B[r][cc] : (T)0;
                    const T mo = pr * (mp[r] - kg * z);
                    *mo_ptr = mo;
                    mo_ptr += g.d;
                    if (r == 0) new_m0 = mo;
                }
                if constexpr (L::VEC) block_store_vec<T, NN>(blk, blkv);
                const T ref = nanmax<T>(tabs<T>(raw0), tabs<T>(new_m0));   // ek1.py:249-251
                if (err_out) err_out[cur.flat] = err;
                if (ref_out) ref_out[cur.flat] = ref;
                if (c.want_norm) {
                    const double ratio = ((double)c.dt * (double)err) / ((double)c.abstol + (double)c.reltol * (double)ref);
                    ratio_sq = ratio * ratio;                              // step.py:101-104
                }
            }
            if (!unit_mode) s_term[threadIdx.x] += ratio_sq;

            // (4) stream the updated blocks back
            if (bulk_out) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    bulk_s2g(sc_out + sc_off, my_sc, (uint32_t)(32 * NN * sizeof(T)));
                    bulk_commit();
                }
            } else {
                __syncwarp();
                if (L::VEC && (reinterpret_cast<uintptr_t>(sc_out) & 15) == 0) {
                    const int total = cur.seg_len * L::UNITS;
                    uint4* out_units = reinterpret_cast<uint4*>(sc_out + sc_off);
                    for (int u = lane; u < total; u += 32) {
                        const int b = u / L::UNITS;
                        out_units[u] = *reinterpret_cast<const uint4*>(my_sc + b * STRIDE + (u - b * L::UNITS) * L::PER_UNIT);
                    }
                } else {
                    const int total = cur.seg_len * NN;
                    for (int e = lane; e < total; e += 32) {
                        const int b = e / NN;
                        sc_out[sc_off + e] = my_sc[b * STRIDE + (e - b * NN)];
                    }
                }
            }
            if (SC_SLOTS == 1 && next < a.num_tiles) {
                // single slot: refill it as soon as this tile's results have left (the other resident CTA covers the gap)
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
                next_bulk = stage_in(tile_coord<Shape>(g, a.ranges.tile(next), a.tiles_x), 0);
            }
            if (unit_mode && (c.want_norm || (kCtl && kComm && a.reduce.world > 0))) {
                // the coordinate's share of the norm: one shared-memory store now, the tile's sum after the next barrier
                s_term[(it & 1u) * kThreads + threadIdx.x] = ratio_sq;
                if (threadIdx.x == 0) s_pending[it & 1u] = (int)tile_id;
            }
            if (dyn_tiles && threadIdx.x == 0) s_next[(it + 1u) & 1u] = fetched;
            tile = next;
            cur_bulk = next_bulk;
        }

#if defined(TDX_DEK1_TIMING) || defined(TDX_TIMELINE)
        if (threadIdx.x == 0) {      // development aid: when did this CTA run out of tiles? (scripts/dek1_cta_spread.py)
            a.partials[1024 + 2 * blockIdx.x] = (double)global_timer_ns();
#ifdef TDX_TIMELINE
            a.partials[5000 + 2 * blockIdx.x + 1] = (double)it;
#endif
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            a.partials[1024 + 2 * blockIdx.x + 1] = (double)smid;
        }
#endif
        if (threadIdx.x == 0) TDX_STAMP(a.partials, att, 1);
        if (!(c.want_norm || kCtl)) break;            // ConstantSteps without a controller: nothing to reduce

        // ---- end of the attempt: everything this CTA wrote must be visible before it reports in ------------------------
        if constexpr (kCtl) {
            if (lane == 0) bulk_wait0();              // the factor blocks have LANDED in global memory (not just left shared memory)
#ifdef TDX_TIMELINE
            if (threadIdx.x == 0) a.partials[3072 + 4 * blockIdx.x + 0] = (double)global_timer_ns();
#endif
            fence_proxy_async_global();               // ... and ordinary stores are ordered before the next attempt's TMA loads
#ifdef TDX_TIMELINE
            if (threadIdx.x == 0) a.partials[3072 + 4 * blockIdx.x + 1] = (double)global_timer_ns();
#endif
        }
        __syncthreads();
        if (unit_mode) flush_tile_partial(it - 1u);
        __syncthreads();
#ifdef TDX_TIMELINE
        if (threadIdx.x == 0) a.partials[3072 + 4 * blockIdx.x + 2] = (double)global_timer_ns();
#endif
        if (threadIdx.x == 0 && poisoned) {           // stale halos: the attempt must not be accepted anywhere
            if (unit_mode) a.red.unit[a.ranges.tile(kComm ? (gridDim.x - 1u - blockIdx.x) : blockIdx.x)] = nan("");
            else s_term[threadIdx.x] = nan("");
        }
        double total = 0.0;
        bool fin = false;                             // thread 0 of the last CTA
        if (unit_mode) {
            const bool last_cta = cta_arrive_last(a.ticket, &s_flag);
#ifdef TDX_TIMELINE
            if (kCtl && threadIdx.x == 0) a.partials[3072 + 4 * blockIdx.x + 3] = (double)global_timer_ns();
#endif
            if (last_cta) {
                if (threadIdx.x == 0 && dyn_tiles) *a.tile_counter = 0u;      // everybody is past its last fetch
                if (threadIdx.x == 0) TDX_STAMP_ANY(a.partials, att, 3);
                // ConstantSteps on one GPU: nothing to add up, the rendezvous alone orders the attempts
                if (c.want_norm || (kComm && a.reduce.world > 0))
                    total = reduce_units(a.red, kComm ? &a.reduce : nullptr, a.reduce.seq + (unsigned long long)att, s_red);
                fin = threadIdx.x == 0;
                if (fin) TDX_STAMP_ANY(a.partials, att, 4);
                if (fin && c.want_norm && a.out_sum) *a.out_sum = total;
            }
        } else {
            fin = grid_sum_cta(s_term[threadIdx.x], s_red, a.partials, a.ticket, c.want_norm ? a.out_sum : nullptr);
        }
        if constexpr (kCtl) {
            if (fin) {
                ctl_finish_attempt(a.ctl, total, a.d_global, s_epoch0 + (unsigned long long)att + 1ull);
                TDX_STAMP_ANY(a.partials, att, 5);
            }
        }
        if (threadIdx.x == 0) TDX_STAMP(a.partials, att, 2);
    }
    if (lane == 0) bulk_wait_read0();
}

// ---------------------------------------------------------------------------------------------
// KroneckerEK0 sweeps
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
struct Ek0Mid {          // written by ek0_mid_kernel, read by sweep B
    T K[kMaxN];          // gain  ek0.py:136
    double err;          // calibrated scalar error estimate  ek0.py:129
    double z_sq;         // fused kernel: the global sum of z^2 (every CTA derives the gain from it)
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
    long long halo_parity_stride;
    const Ek0Mid<T, N>* mid;
    double* partials;
    unsigned int* ticket;
    double* out_sum;
    RedPlan red;
    unsigned int num_tiles, tiles_x;
    TileRanges ranges;
    CommWait wait;
    CommReduce reduce;
    PackPeer<T> pack;
    int reverse;             // walk the tiles backwards (sweep B re-reads what sweep A touched last: L2 reuse)
};

// Both sweeps: persistent CTAs, the loads of tile i+1 are in flight while tile i is computed.
//   sweep A (kUpdate = false): z = p1 mp[1] - f(p0 mp[0]);  sum z^2                          ek0.py:166-172
//   sweep B (kUpdate = true) : m_new = P (mp - K z); ref = |m_new[0]|; sum (dt err / tol)^2  ek0.py:132-140,181-184
template <class IVP, typename T, int N, bool kUpdate>
__global__ void __launch_bounds__(kThreads, 3) ek0_sweep_kernel(const __grid_constant__ Ek0Args<T, N> a) {
    using Shape = typename IVP::Shape;
    using L = SmemLayout<T, N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_red = reinterpret_cast<double*>(smem_raw + L::BAR_BYTES);
    T* s_at0 = reinterpret_cast<T*>(smem_raw + L::BAR_BYTES + L::RED_BYTES);
    constexpr size_t AT_STRIDE = L::at_bytes(ExtTile<IVP>::ELEMS) / sizeof(T);
    __shared__ int s_flag, s_ok;
    const Geo& g = a.g;
    const Consts<T, N>& c = a.c;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long hseq = a.wait.seq;
    const T* halo_lo = a.halo_lo ? a.halo_lo + (long long)(hseq & 1ull) * a.halo_parity_stride : nullptr;
    const T* halo_hi = a.halo_hi ? a.halo_hi + (long long)(hseq & 1ull) * a.halo_parity_stride : nullptr;

    T K[N];
    if (kUpdate) {
#pragma unroll
        for (int r = 0; r < N; ++r) K[r] = a.mid->K[r];
    }
    auto tile_of = [&](unsigned i) { return a.ranges.tile(a.reverse ? (a.num_tiles - 1u - i) : i); };

    // virtual index of the i-th tile this launch walks (reversed for sweep B) and whether it reads halos
    auto vidx = [&](unsigned i) { return a.reverse ? (a.num_tiles - 1u - i) : i; };
    bool halos_ready = false, poisoned = false;
    auto wait_if_needed = [&](unsigned i) {
        if (!halos_ready && vidx(i) >= a.ranges.wait_from) {
            if (!cta_wait_halos(a.wait, hseq, &s_ok)) poisoned = true;
            halos_ready = true;
        }
    };

    if (a.pack.enabled)
        for (int part = (int)blockIdx.x; part < 2 * kPackCtas; part += (int)gridDim.x)
            pack_halo_part<T, N>(g, c, a.mean_in, a.pack, a.pack.seq, part);
    unsigned i = gridDim.x - 1u - blockIdx.x;
    const unsigned first_tile = i < a.num_tiles ? tile_of(i) : 0u;
    double cta_acc = 0.0;
    int pending_unit = -1;
    const bool reduces = !kUpdate || c.want_norm;
    Coord co{};
    unsigned co_tile = 0u;
    FrontRegs<IVP, T, N> fr;
    if (i < a.num_tiles) {
        wait_if_needed(i);
        co_tile = tile_of(i);
        co = tile_coord<Shape>(g, co_tile, a.tiles_x);
        front_load<IVP, T, N>(g, c, a.mean_in, halo_lo, halo_hi, co, fr);
    }
    auto flush_tile_partial = [&](unsigned it_prev) {
        if (threadIdx.x == 0 && pending_unit >= 0) {
            const double* ws = s_red + (it_prev & 1u) * kWarps;
            double b = 0.0;
#pragma unroll
            for (int k = 0; k < kWarps; ++k) b += ws[k];
            if (a.red.unit) a.red.unit[pending_unit] = b;
            else cta_acc += b;
            pending_unit = -1;
        }
    };
    unsigned it = 0;
    for (; i < a.num_tiles; ++it) {
        T mp[N], fx, jac;
        front_eval<IVP, T, N>(g, c, co, fr, s_at0 + (it & 1) * AT_STRIDE, mp, fx, jac);
        flush_tile_partial(it - 1u);
        const Coord cur = co;
        const unsigned cur_tile = co_tile;
        i += gridDim.x;
        if (i < a.num_tiles) {
            wait_if_needed(i);
            co_tile = tile_of(i);
            co = tile_coord<Shape>(g, co_tile, a.tiles_x);
            front_load<IVP, T, N>(g, c, a.mean_in, halo_lo, halo_hi, co, fr);
        }
        double term = 0.0;
        if (cur.valid) {
            const T z = c.p[1] * mp[1] - fx;
            if (!kUpdate) {
                term = (double)z * (double)z;
            } else {
                T m0 = (T)0;
                T* mo_ptr = a.mean_out + cur.flat;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const T mo = c.p[r] * (mp[r] - K[r] * z);
                    *mo_ptr = mo;
                    mo_ptr += g.d;
                    if (r == 0) m0 = mo;
                }
                const T ref = tabs<T>(m0);
                if (a.ref_out) a.ref_out[cur.flat] = ref;
                if (c.want_norm) {
                    const double tol = (double)c.abstol + (double)c.reltol * (double)ref;
                    term = 1.0 / (tol * tol);
                }
            }
        }
        if (reduces) {
            const double ws = warp_sum(term);
            if (lane == 0) s_red[(it & 1u) * kWarps + w] = ws;
            if (threadIdx.x == 0) pending_unit = (int)cur_tile;
        }
    }
    if (!reduces) return;
    __syncthreads();
    flush_tile_partial(it - 1u);
    if (threadIdx.x == 0 && poisoned) {
        if (a.red.unit) a.red.unit[first_tile] = nan("");
        else cta_acc = nan("");
    }
    // sweep B:  sum_i (dt err / tol_i)^2 = (dt err)^2 sum_i 1 / tol_i^2     step.py:101-104 with a scalar error estimate
    const double scale = kUpdate ? ((double)c.dt * (double)c.dt) : 1.0;
    if (a.red.unit != nullptr) {
        if (cta_arrive_last(a.ticket, &s_flag)) {
            double total = reduce_units(a.red, &a.reduce, a.reduce.seq, s_red);
            if (threadIdx.x == 0) {
                if (kUpdate) {
                    const double r = a.mid->err;
                    total *= r * r;
                    total *= scale;
                }
                *a.out_sum = total;
            }
        }
    } else {
        // a launch over a part of the tiles (interior / rim): per-CTA partials; the caller adds the launches' sums
        if (kUpdate && threadIdx.x == 0) {
            const double r = a.mid->err;
            cta_acc *= r * r;
            cta_acc *= scale;
        }
        grid_sum_cta(cta_acc, s_red, a.partials, a.ticket, a.out_sum);
    }
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
    const double z_sq = a.z_sq_sum[0] + ((c.flags & 8u) ? a.z_sq_sum[1] : 0.0);   // TDX_STEP_TWO_PARTS
    const double sigma_sq = z_sq / HQH / a.d_global;
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
// sqrt.propagate_cholesky_factor / sqrtm_to_cholesky as a standalone batched op (sqrt.py:8-30): the lower factor
// L = R^T of the QR decomposition of a 'right' square root St (M x N per batch element, M = N or 2 N), LAPACK
// dgeqr2 / dlarfg convention like stacked_qr.  One thread per matrix, everything in registers; the inputs are read
// as S1[b] (N x N) and, for M = 2 N, S2[b] (N x N) with St = [S1^T ; S2^T]  (``transposed`` = 0: the rows of St are given).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct QrArgs {
    const T* s1;        // (batch, N, N)
    const T* s2;        // (batch, N, N) or null (M = N)
    T* out;             // (batch, N, N) lower triangular
    long long batch;
    int s1_broadcast, s2_broadcast;   // the same matrix for every batch element
    int rows_given;     // 1: s1 / s2 hold St itself (rows of the stacked matrix) instead of S1, S2
};

template <typename T, int N, int M>
__global__ void __launch_bounds__(128) batched_qr_lower_kernel(const __grid_constant__ QrArgs<T> a) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    // C[c][r] = St[r][c]: column c of the stacked matrix
    T C[N][M];
    const T* p1 = a.s1 + (a.s1_broadcast ? 0 : b * N * N);
    const T* p2 = (M > N) ? a.s2 + (a.s2_broadcast ? 0 : b * N * N) : nullptr;
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = 0; r < M; ++r) {
            const T* src = (r < N) ? p1 : p2;
            const int rr = (r < N) ? r : r - N;
            C[c][r] = a.rows_given ? src[rr * N + c] : src[c * N + rr];     // St = S^T: St[r][c] = S[c][r]
        }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const T alpha = C[k][k];
        T ns = (T)0;
#pragma unroll
        for (int r = k + 1; r < M; ++r) ns = fma(C[k][r], C[k][r], ns);
        if (ns != (T)0) {
            const T nrm = tsqrt<T>(fma(alpha, alpha, ns));
            const T beta = (alpha >= (T)0 && !(alpha == (T)0 && signbit(alpha))) ? -nrm : nrm;
            const T u0 = alpha - beta;
            const T gamma = (T)1 / (beta * (beta - alpha));
#pragma unroll
            for (int cc = k + 1; cc < N; ++cc) {
                T w = u0 * C[cc][k];
#pragma unroll
                for (int r = k + 1; r < M; ++r) w = fma(C[k][r], C[cc][r], w);
                w *= gamma;
                C[cc][k] = fma(-w, u0, C[cc][k]);
#pragma unroll
                for (int r = k + 1; r < M; ++r) C[cc][r] = fma(-w, C[k][r], C[cc][r]);
            }
            C[k][k] = beta;
        }
    }
    T* o = a.out + b * N * N;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) o[i * N + j] = (j <= i) ? C[i][j] : (T)0;   // L[i][j] = R[j][i] = C[i][j]
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
    unsigned int tiles_x;
};

template <class IVP, typename T>
__global__ void __launch_bounds__(kThreads) eval_f_kernel(const __grid_constant__ EvalArgs<T> a) {
    using Shape = typename IVP::Shape;
    using E = ExtTile<IVP>;
    __shared__ T s_ext[E::ELEMS];
    const Coord co = tile_coord<Shape>(a.g, blockIdx.x, a.tiles_x);
    PlainAt<T> at{a.y};
    const T own = co.valid ? __ldg(a.y + co.flat)
                           : (E::NRING > 0 ? IVP::ring_value(a.g, tile_row0(co), tile_col0(co), co.field, co.tr + E::HV,
                                                             co.tcol + E::HL, at, a.halo_lo, a.halo_hi)
                                           : (T)0);
    s_ext[E::own(co)] = own;
    if (E::NRING > 0 && (int)threadIdx.x < Shape::NF * E::NRING) {
        const int field = threadIdx.x / (E::NRING > 0 ? E::NRING : 1);
        int er, ec;
        E::ring(threadIdx.x - field * E::NRING, er, ec);
        s_ext[E::index(field, er, ec)] =
            IVP::ring_value(a.g, tile_row0(co), tile_col0(co), field, er, ec, at, a.halo_lo, a.halo_hi);
    }
    __syncthreads();
    if (co.valid) {
        T fx, jac;
        IVP::eval(a.g, co, s_ext, own, fx, jac);
        a.fy[co.flat] = fx;
        if (a.jd) a.jd[co.flat] = jac;
    }
}

}  // namespace tdx
