// tornadox_b200 C ABI (include/tornadox_b200.h): context, argument validation, dispatch to the kernels.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/tornadox_b200.h"
#include "tdx_ops.cuh"

namespace tdx {

// per-instantiation launch tables (tdx_inst_*.cu)
#define TDX_DECL(NAME) const Ops* ops_##NAME();
TDX_DECL(f64_2) TDX_DECL(f64_3) TDX_DECL(f64_4) TDX_DECL(f64_5) TDX_DECL(f64_6)
TDX_DECL(f32_2) TDX_DECL(f32_3) TDX_DECL(f32_4) TDX_DECL(f32_5) TDX_DECL(f32_6)
#undef TDX_DECL
const EvalOps* eval_ops_f64();
const EvalOps* eval_ops_f32();

const Ops* find_ops(int dtype, int n) {
    if (dtype == TDX_F64) {
        switch (n) {
            case 2: return ops_f64_2();
            case 3: return ops_f64_3();
            case 4: return ops_f64_4();
            case 5: return ops_f64_5();
            case 6: return ops_f64_6();
        }
    } else if (dtype == TDX_F32) {
        switch (n) {
            case 2: return ops_f32_2();
            case 3: return ops_f32_3();
            case 4: return ops_f32_4();
            case 5: return ops_f32_5();
            case 6: return ops_f32_6();
        }
    }
    return nullptr;
}

const EvalOps* find_eval_ops(int dtype) {
    if (dtype == TDX_F64) return eval_ops_f64();
    if (dtype == TDX_F32) return eval_ops_f32();
    return nullptr;
}

long long grid_for(const Geo& g) {
    switch (g.kind) {
        case IVP_VDP: return num_tiles<VanderPol<double>::Shape>(g.rows, g.cols);
        case IVP_LORENZ96: return num_tiles<Lorenz96<double>::Shape>(g.rows, g.cols);
        case IVP_BRUSS: return num_tiles<Brusselator<double>::Shape>(g.rows, g.cols);
        case IVP_FHN2D: return num_tiles<Fhn2d<double>::Shape>(g.rows, g.cols);
        case IVP_BURGERS1D: return num_tiles<Burgers1d<double>::Shape>(g.rows, g.cols);
        case IVP_WAVE1D: return num_tiles<Wave1d<double>::Shape>(g.rows, g.cols);
        case IVP_THREEBODY: return num_tiles<ThreeBody<double>::Shape>(g.rows, g.cols);
        case IVP_PLEIADES: return num_tiles<Pleiades<double>::Shape>(g.rows, g.cols);
        case IVP_EXTERNAL: return num_tiles<External<double>::Shape>(g.rows, g.cols);
    }
    return 0;
}

// Deterministic all-reduce (sum in rank order) of up to 4 doubles through the peers' reduction slots.
struct AllReduceArgs {
    CommHeader* local;
    CommHeader* peer[kMaxWorld];
    double* vals;
    int count, rank, world;
    unsigned long long seq;
    unsigned long long timeout_ns;
};

__global__ void comm_allreduce_kernel(const __grid_constant__ AllReduceArgs a) {
    const int r = threadIdx.x;
    const int parity = (int)(a.seq & 1ull);
    __shared__ int s_bad;
    if (r == 0) s_bad = 0;
    __syncthreads();
    if (r < a.world) {
        CommHeader* dst = a.peer[r];
        for (int j = 0; j < a.count; ++j) dst->red_val[parity][a.rank][j] = a.vals[j];
        dst->red_cnt[parity][a.rank] = a.count;
        __threadfence_system();
        st_release_sys(&dst->red_flag[parity][a.rank], a.seq);
    }
    __syncthreads();
    if (r < a.world && !spin_until(&a.local->red_flag[parity][r], a.seq, &a.local->error, a.timeout_ns)) s_bad = 1;
    __syncthreads();
    if (r < a.count) {
        double acc = 0.0;
        for (int k = 0; k < a.world; ++k) acc += *((volatile double*)&a.local->red_val[parity][k][r]);
        a.vals[r] = s_bad ? nan("") : acc;     // a dead peer must not go unnoticed
    }
}

// ---- prior constants on the host (iwp.py:19-37, 65-73) -------------------------------------------

// Ql = chol(Q), Q[i][j] = 1 / (2 nu + 1 - i - j).  Computed in long double and rounded, so the result is
// the correctly rounded factor up to ~1 ulp (LAPACK dpotrf in the reference agrees to cond(Q) * eps).
static void prior_host(int nu, double* A, double* Ql) {
    const int n = nu + 1;
    long double L[kMaxN][kMaxN] = {};
    for (int j = 0; j < n; ++j) {
        long double s = 1.0L / (long double)(2 * nu + 1 - 2 * j);
        for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
        L[j][j] = sqrtl(s);
        for (int i = j + 1; i < n; ++i) {
            long double t = 1.0L / (long double)(2 * nu + 1 - i - j);
            for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
            L[i][j] = t / L[j][j];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (A) A[i * n + j] = (j >= i) ? binom(nu - i, nu - j) : 0.0;
            if (Ql) Ql[i * n + j] = (j <= i) ? (double)L[i][j] : 0.0;
        }
}

// p[k] = |dt|^(nu-k+1/2) / (nu-k)!   (iwp.py:65-73) with libm's pow, as NumPy / the oracle evaluate it.  (The device-resident
// step controller builds its vectors with nordsieck_vectors(): sqrt and products, a few ulp from these.)
static void nordsieck_host(int nu, double dt, double* p, double* pinv) {
    const double adt = std::fabs(dt);
    for (int k = 0; k <= nu; ++k) {
        const int q = nu - k;
        double fact = 1.0;
        for (int i = 2; i <= q; ++i) fact *= (double)i;
        const double e = (double)q + 0.5;
        if (p) p[k] = std::pow(adt, e) / fact;
        if (pinv) pinv[k] = std::pow(adt, -e) * fact;
    }
}

}  // namespace tdx

using namespace tdx;

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

static int cuda_fail(int cuda_code, const char* where) {
    if (cuda_code == 0) return TDX_OK;
    if (cuda_code < 0) return fail(TDX_ERR_INVALID, std::string(where) + ": unknown vector field kind");
    g_last_error = std::string(where) + ": " + cudaGetErrorString((cudaError_t)cuda_code);
    return TDX_ERR_CUDA;
}

struct HostPipe;

struct tdx_ctx {
    StepCtl* ctl = nullptr;          // device-resident step controller block
    StepCtl* ctl_host = nullptr;     // pinned staging copy
    tdx_steprule ctl_rule{};
    double* host_norm = nullptr;     // mapped pinned: [0] reduced norm sum written by the kernel epilogue (full-step loop)
    double* host_norm_dev = nullptr; // its device alias
    double* zsq_scratch = nullptr;   // device scalar for the EK0 z^2 sum inside the full-step loop
    HostPipe* pipe = nullptr;        // chunked host-buffer pipeline (tdx_dek1_attempt_step_host), built on first use
    int device = 0;
    int dtype = TDX_F64;
    bool has_problem = false;
    tdx_problem_desc desc{};
    Geo geo{};
    int n = 0;
    const Ops* ops = nullptr;
    const EvalOps* eval_ops = nullptr;
    double ql_packed[kMaxN * (kMaxN + 1) / 2] = {};
    // workspace
    double* partials = nullptr;
    long long partials_cap = 0;
    unsigned int* ticket = nullptr;   // [8]: [0..1] arrival tickets (the rim launch of an interior/rim pair uses the second), [4..5] tile counters
    void* mid = nullptr;
    long long halo_lo = 0, halo_hi = 0;
    int64_t launches = 0;
    // peer-memory exchange
    unsigned char* comm_local = nullptr;
    unsigned char* comm_peer[kMaxWorld] = {};
    bool comm_opened[kMaxWorld] = {};
    int rank = 0, world = 1, lower = -1, upper = -1;
    unsigned long long halo_seq = 0, red_seq = 0;
    unsigned long long mid_seq = 0;  // fused EK0 attempts so far (the in-kernel barrier flag counts them)
    unsigned long long cl_seq = 0;   // fused EK0: "new factor written" flag, one count per attempt
    long long halo_cap = 0;          // elements per parity and side
    // unit partials of the sharding-independent reductions: [2][unit_cap]
    double* units = nullptr;
    long long unit_cap = 0;
    unsigned long long timeout_ns = 0;   // in-kernel waits (peers, grid barriers); TDX_WAIT_TIMEOUT_MS, default 4000
    // device-resident controller extras
    double* ctl_stops = nullptr;     // device copy of the time stops
    long long ctl_stops_cap = 0;
    HostFeed* feed = nullptr;        // mapped pinned
    HostFeed* feed_dev = nullptr;
    int ctl_nbuf = 2;
};

static unsigned long long env_timeout_ns() {
    const char* v = std::getenv("TDX_WAIT_TIMEOUT_MS");
    double ms = v ? std::atof(v) : 4000.0;
    if (!(ms > 0.0)) ms = 4000.0;
    return (unsigned long long)(ms * 1.0e6);
}

static size_t comm_bytes(const tdx_ctx* ctx) {
    return sizeof(CommHeader) + 4 * (size_t)ctx->halo_cap * (ctx->dtype == TDX_F64 ? 8 : 4) + 64;
}
// halo storage behind the header: lo[parity] at (0, 1), hi[parity] at (2, 3)
static unsigned char* comm_halo(const tdx_ctx* ctx, unsigned char* base, int hi, int parity) {
    const size_t es = ctx->dtype == TDX_F64 ? 8 : 4;
    return base + sizeof(CommHeader) + (size_t)(2 * hi + parity) * (size_t)ctx->halo_cap * es;
}

static size_t esize(int dtype) { return dtype == TDX_F64 ? 8 : 4; }

extern "C" {

const char* tdx_last_error(void) { return g_last_error.c_str(); }
int tdx_abi_version(void) { return TDX_ABI_VERSION; }

int tdx_prior(int nu, double* A_out, double* Ql_out) {
    if (nu < 1 || nu + 1 > kMaxN) return fail(TDX_ERR_INVALID, "tdx_prior: nu out of range");
    prior_host(nu, A_out, Ql_out);
    return TDX_OK;
}

int tdx_propagate_cholesky_factor(int dtype, int n, int64_t batch, const void* S1, const void* S2, void* out,
                                  int s1_broadcast, int s2_broadcast, int rows_given, void* stream) {
    const Ops* ops = find_ops(dtype, n);
    if (!ops) return fail(TDX_ERR_UNSUPPORTED, "tdx_propagate_cholesky_factor: no kernel instantiated for this (n, dtype)");
    if (!S1 || !out || batch < 0) return fail(TDX_ERR_INVALID, "tdx_propagate_cholesky_factor: null buffer or negative batch");
    const int code = ops->qr_lower(S1, S2, out, (long long)batch, s1_broadcast, s2_broadcast, rows_given, (cudaStream_t)stream);
    if (code == 0) return TDX_OK;
    g_last_error = std::string("tdx_propagate_cholesky_factor: ") + cudaGetErrorString((cudaError_t)code);
    return TDX_ERR_CUDA;
}

int tdx_nordsieck(int nu, double dt, double* p_out, double* p_inv_out) {
    if (nu < 0 || nu + 1 > kMaxN) return fail(TDX_ERR_INVALID, "tdx_nordsieck: nu out of range");
    nordsieck_host(nu, dt, p_out, p_inv_out);
    return TDX_OK;
}

int tdx_create(tdx_ctx** out, int device, int dtype) {
    if (!out) return fail(TDX_ERR_INVALID, "tdx_create: null output pointer");
    if (dtype != TDX_F64 && dtype != TDX_F32) return fail(TDX_ERR_INVALID, "tdx_create: dtype must be TDX_F64 or TDX_F32");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(TDX_ERR_CUDA, std::string("tdx_create: no CUDA device available (") + cudaGetErrorString(e) +
                                      "); tornadox_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(TDX_ERR_INVALID, "tdx_create: device ordinal out of range");
    tdx_ctx* ctx = new tdx_ctx();
    ctx->device = device;
    ctx->dtype = dtype;
    ctx->eval_ops = find_eval_ops(dtype);
    ctx->timeout_ns = env_timeout_ns();
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    e = cudaMalloc(&ctx->ticket, 8 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(ctx->ticket, 0, 8 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->mid, 256);
    if (e == cudaSuccess) e = cudaMemset(ctx->mid, 0, 256);
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        delete ctx;
        return cuda_fail((int)e, "tdx_create");
    }
    *out = ctx;
    return TDX_OK;
}

static void destroy_pipe(tdx_ctx* ctx);

int tdx_destroy(tdx_ctx* ctx) {
    if (!ctx) return TDX_OK;
    destroy_pipe(ctx);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    if (ctx->ctl_host) cudaFreeHost(ctx->ctl_host);
    cudaFree(ctx->ctl);
    if (ctx->host_norm) cudaFreeHost(ctx->host_norm);
    cudaFree(ctx->zsq_scratch);
    cudaFree(ctx->partials);
    cudaFree(ctx->units);
    cudaFree(ctx->ctl_stops);
    if (ctx->feed) cudaFreeHost(ctx->feed);
    cudaFree(ctx->ticket);
    cudaFree(ctx->mid);
    for (int r = 0; r < kMaxWorld; ++r)
        if (ctx->comm_opened[r]) cudaIpcCloseMemHandle(ctx->comm_peer[r]);
    cudaFree(ctx->comm_local);
    cudaSetDevice(prev);
    delete ctx;
    return TDX_OK;
}

int tdx_set_problem(tdx_ctx* ctx, const tdx_problem_desc* d) {
    if (!ctx || !d) return fail(TDX_ERR_INVALID, "tdx_set_problem: null argument");
    if (d->nu < 1 || d->nu + 1 > kMaxN) return fail(TDX_ERR_INVALID, "tdx_set_problem: num_derivatives must be in [1, 6]");
    const Ops* ops = find_ops(ctx->dtype, d->nu + 1);
    if (!ops) return fail(TDX_ERR_UNSUPPORTED, "tdx_set_problem: no kernel instantiated for this (num_derivatives, dtype)");
    if (d->grid_rows < 1 || d->grid_cols < 1) return fail(TDX_ERR_INVALID, "tdx_set_problem: empty grid");
    if (d->shard_begin < 0 || d->shard_end <= d->shard_begin)
        return fail(TDX_ERR_INVALID, "tdx_set_problem: empty or negative shard range");
    Geo g{};
    g.kind = d->kind;
    g.grows = d->grid_rows;
    g.gcols = d->grid_cols;
    for (int i = 0; i < 8; ++i) g.prm[i] = 0.0;
    switch (d->kind) {
        case TDX_IVP_VANDERPOL:
            if (d->grid_rows != 1 || d->grid_cols != 1 || d->shard_begin != 0 || d->shard_end != 1)
                return fail(TDX_ERR_INVALID, "tdx_set_problem: vanderpol is a single point (d = 2) and does not shard");
            if (d->n_params < 1) return fail(TDX_ERR_INVALID, "tdx_set_problem: vanderpol needs params = [mu]");
            g.nf = 2;
            g.rows = 1; g.cols = 1; g.grow0 = 0; g.gcol0 = 0;
            g.prm[0] = d->params[0];
            g.prm[1] = (d->n_params >= 2 && d->params[1] != 0.0) ? 1.0 : 0.0;     // vanderpol_julia's scaling (ivp.py:52-73)
            break;
        case TDX_IVP_THREEBODY:
        case TDX_IVP_PLEIADES: {
            const long long dim = d->kind == TDX_IVP_THREEBODY ? 4 : 28;
            if (d->grid_rows != 1 || d->grid_cols != dim || d->shard_begin != 0 || d->shard_end != dim)
                return fail(TDX_ERR_INVALID, "tdx_set_problem: threebody (d = 4) / pleiades (d = 28) are laid out as one field of d "
                                             "points (grid_rows = 1, grid_cols = d) and do not shard");
            g.nf = 1;
            g.rows = 1; g.cols = (int)dim; g.grow0 = 0; g.gcol0 = 0;
            break;
        }
        case TDX_IVP_EXTERNAL:
            if (d->grid_rows != 1 || d->grid_cols < 1 || d->shard_begin != 0 || d->shard_end != d->grid_cols)
                return fail(TDX_ERR_INVALID, "tdx_set_problem: an external vector field is laid out as one field of d points "
                                             "(grid_rows = 1, grid_cols = d) and runs on one GPU");
            if (d->grid_cols > 2147483647LL) return fail(TDX_ERR_INVALID, "tdx_set_problem: state too large");
            g.nf = 1;
            g.rows = 1; g.cols = (int)d->grid_cols; g.grow0 = 0; g.gcol0 = 0;
            break;
        case TDX_IVP_LORENZ96:
            if (d->grid_rows != 1) return fail(TDX_ERR_INVALID, "tdx_set_problem: lorenz96 is one-dimensional (grid_rows = 1)");
            if (d->shard_end > d->grid_cols) return fail(TDX_ERR_INVALID, "tdx_set_problem: shard exceeds the chain");
            if (d->n_params < 1) return fail(TDX_ERR_INVALID, "tdx_set_problem: lorenz96 needs params = [forcing]");
            if ((d->shard_end - d->shard_begin) != d->grid_cols && (d->shard_end - d->shard_begin) < 2)
                return fail(TDX_ERR_INVALID, "tdx_set_problem: a lorenz96 shard needs at least 2 points");
            g.nf = 1;
            g.rows = 1; g.cols = (int)(d->shard_end - d->shard_begin); g.grow0 = 0; g.gcol0 = d->shard_begin;
            g.prm[0] = d->params[0];
            break;
        case TDX_IVP_BRUSSELATOR:
            if (d->grid_rows != 1) return fail(TDX_ERR_INVALID, "tdx_set_problem: brusselator is one-dimensional (grid_rows = 1)");
            if (d->shard_end > d->grid_cols) return fail(TDX_ERR_INVALID, "tdx_set_problem: shard exceeds the chain");
            g.nf = 2;
            g.rows = 1; g.cols = (int)(d->shard_end - d->shard_begin); g.grow0 = 0; g.gcol0 = d->shard_begin;
            {   // ivp.py:223-224  const = alpha * (N + 1) ** 2, alpha = 1/50
                const double alpha = 1.0 / 50.0;
                const double np1 = (double)(d->grid_cols + 1);
                g.prm[0] = alpha * (np1 * np1);
            }
            break;
        case TDX_IVP_BURGERS1D:
        case TDX_IVP_WAVE1D:
            if (d->grid_rows != 1) return fail(TDX_ERR_INVALID, "tdx_set_problem: burgers_1d / wave_1d are one-dimensional (grid_rows = 1)");
            if (d->grid_cols < 3) return fail(TDX_ERR_INVALID, "tdx_set_problem: burgers_1d / wave_1d need at least 3 mesh points");
            if (d->shard_begin != 0 || d->shard_end != d->grid_cols)
                return fail(TDX_ERR_UNSUPPORTED, "tdx_set_problem: burgers_1d / wave_1d run on one GPU only (their end-point "
                                                 "rules couple the two ends of the mesh)");
            if (d->n_params < 2) return fail(TDX_ERR_INVALID, "tdx_set_problem: burgers_1d / wave_1d need params = [dx, diffusion_param]");
            g.nf = d->kind == TDX_IVP_WAVE1D ? 2 : 1;
            g.rows = 1; g.cols = (int)d->grid_cols; g.grow0 = 0; g.gcol0 = 0;
            g.prm[0] = 1.0 / d->params[0];                                  // 1 / dx
            g.prm[1] = d->params[1] / (d->params[0] * d->params[0]);        // diffusion_param / dx^2  (ivp.py:88, 126)
            break;
        case TDX_IVP_FHN2D:
            if (d->grid_rows != d->grid_cols)
                return fail(TDX_ERR_INVALID,
                            "The diagonal Jacobian only works for quadratic spatial grids.");   // ivp.py:427-430
            if (d->grid_rows < 2) return fail(TDX_ERR_INVALID, "tdx_set_problem: fhn_2d needs at least a 2 x 2 grid");
            if (d->shard_end > d->grid_rows) return fail(TDX_ERR_INVALID, "tdx_set_problem: shard exceeds the grid");
            if (d->n_params < 5) return fail(TDX_ERR_INVALID, "tdx_set_problem: fhn_2d needs params = [dx, a, b, k, tau]");
            g.nf = 2;
            g.rows = (int)(d->shard_end - d->shard_begin); g.cols = (int)d->grid_cols;
            g.grow0 = d->shard_begin; g.gcol0 = 0;
            {
                const double dx = d->params[0];
                g.prm[0] = 1.0 / (dx * dx);   // ivp.py:493-494
                g.prm[1] = d->params[1]; g.prm[2] = d->params[2]; g.prm[3] = d->params[3]; g.prm[4] = d->params[4];
                g.prm[5] = 1.0 / d->params[4];   // 1 / tau (multiplied, not divided, in-kernel)
            }
            break;
        default:
            return fail(TDX_ERR_UNSUPPORTED,
                        "tdx_set_problem: unknown vector field kind; vanderpol, lorenz96, brusselator, fhn_2d, burgers_1d, wave_1d, "
                        "threebody and pleiades are evaluated in-kernel, TDX_IVP_EXTERNAL takes the caller's f / diag J arrays");
    }
    g.npts = (long long)g.rows * g.cols;
    g.d = (long long)g.nf * g.npts;

    long long lo = 0, hi = 0;
    switch (g.kind) {
        case IVP_VDP: VanderPol<double>::halo_sizes(g, lo, hi); break;
        case IVP_LORENZ96: Lorenz96<double>::halo_sizes(g, lo, hi); break;
        case IVP_BRUSS: Brusselator<double>::halo_sizes(g, lo, hi); break;
        case IVP_FHN2D: Fhn2d<double>::halo_sizes(g, lo, hi); break;
        default: break;   // burgers_1d, wave_1d, threebody, pleiades: never sharded
    }

    long long grid = grid_for(g);
    if (grid < 4096) grid = 4096;   // the fused EK0 kernel stores two partial sums per resident CTA
    if (grid > 2147483647LL) return fail(TDX_ERR_INVALID, "tdx_set_problem: shard too large for one launch");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaSuccess;
    if (grid > ctx->partials_cap) {
        cudaFree(ctx->partials);
        ctx->partials = nullptr;
        e = cudaMalloc(&ctx->partials, 2 * sizeof(double) * (size_t)grid);
        ctx->partials_cap = (e == cudaSuccess) ? grid : 0;
    }
    {
        // one partial per 256-coordinate tile, or per (grid row, 256-column block) of the fused fhn_2d kernel
        long long units = grid_for(g);
        if (g.kind == IVP_FHN2D) units = std::max(units, (long long)g.rows * ((g.cols + TDX_EK0_W - 1) / TDX_EK0_W));
        units = (units + 63) / 64 * 64;
        if (e == cudaSuccess && units > ctx->unit_cap) {
            cudaFree(ctx->units);
            ctx->units = nullptr;
            e = cudaMalloc(&ctx->units, 2 * sizeof(double) * (size_t)units);
            ctx->unit_cap = (e == cudaSuccess) ? units : 0;
        }
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail((int)e, "tdx_set_problem");

    ctx->desc = *d;
    ctx->geo = g;
    ctx->n = d->nu + 1;
    ctx->ops = ops;
    ctx->halo_lo = lo;
    ctx->halo_hi = hi;
    double Ql[kMaxN * kMaxN];
    prior_host(d->nu, nullptr, Ql);
    for (int i = 0; i < ctx->n; ++i)
        for (int j = 0; j <= i; ++j) ctx->ql_packed[i * (i + 1) / 2 + j] = Ql[i * ctx->n + j];
    ctx->has_problem = true;
    return TDX_OK;
}

int tdx_set_diffusion_sqrt(tdx_ctx* ctx, const double* Ql) {
    if (!ctx || !ctx->has_problem) return fail(TDX_ERR_INVALID, "tdx_set_diffusion_sqrt: no problem set");
    if (!Ql) return fail(TDX_ERR_INVALID, "tdx_set_diffusion_sqrt: null factor");
    for (int i = 0; i < ctx->n; ++i)
        for (int j = 0; j < ctx->n; ++j) {
            const double v = Ql[i * ctx->n + j];
            if (!std::isfinite(v) || (j > i && v != 0.0))
                return fail(TDX_ERR_INVALID, "tdx_set_diffusion_sqrt: factor must be finite and lower triangular");
            if (j <= i) ctx->ql_packed[i * (i + 1) / 2 + j] = v;
        }
    return TDX_OK;
}

int64_t tdx_local_dim(const tdx_ctx* ctx) { return (ctx && ctx->has_problem) ? (int64_t)ctx->geo.d : -1; }

int tdx_halo_sizes(const tdx_ctx* ctx, int64_t* lo, int64_t* hi) {
    if (!ctx || !ctx->has_problem) return fail(TDX_ERR_INVALID, "tdx_halo_sizes: no problem set");
    if (lo) *lo = ctx->halo_lo;
    if (hi) *hi = ctx->halo_hi;
    return TDX_OK;
}

int tdx_set_external(tdx_ctx* ctx, const void* fx, const void* jdiag) {
    if (!ctx || !ctx->has_problem) return fail(TDX_ERR_INVALID, "tdx_set_external: no problem set");
    if (ctx->geo.kind != IVP_EXTERNAL) return fail(TDX_ERR_INVALID, "tdx_set_external: the context's vector field is evaluated in-kernel");
    if (!fx) return fail(TDX_ERR_INVALID, "tdx_set_external: null f array");
    ctx->geo.ext_f = fx;
    ctx->geo.ext_j = jdiag;
    return TDX_OK;
}

int64_t tdx_launch_count(const tdx_ctx* ctx) { return ctx ? ctx->launches : 0; }

int tdx_debug_read_workspace(tdx_ctx* ctx, double* host_out, int64_t offset, int64_t count) {
    if (!ctx || !ctx->partials || !host_out || offset < 0 || count < 0 || offset + count > 2 * ctx->partials_cap)
        return fail(TDX_ERR_INVALID, "tdx_debug_read_workspace: bad range");
    cudaError_t e = cudaMemcpy(host_out, ctx->partials + offset, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? TDX_OK : cuda_fail((int)e, "tdx_debug_read_workspace");
}

}  // extern "C"

// ---- helpers ---------------------------------------------------------------------------------------

static int check_ready(const tdx_ctx* ctx, const char* where) {
    if (!ctx) return fail(TDX_ERR_INVALID, std::string(where) + ": null context");
    if (!ctx->has_problem) return fail(TDX_ERR_INVALID, std::string(where) + ": tdx_set_problem has not been called");
    return TDX_OK;
}

// step calls on an external vector field need the caller's f array first
static int check_external(const tdx_ctx* ctx, const char* where) {
    if (ctx->geo.kind == IVP_EXTERNAL && !ctx->geo.ext_f)
        return fail(TDX_ERR_INVALID, std::string(where) + ": external vector field: call tdx_predict_state, evaluate f there and "
                                                          "hand it in with tdx_set_external before every attempt");
    return TDX_OK;
}

static int check_halos(const tdx_ctx* ctx, const void* lo, const void* hi, const char* where, unsigned flags = 0u) {
    if ((flags & 6u) == TDX_STEP_TILES_INTERIOR) return TDX_OK;   // interior tiles never read the halo buffers
    if (ctx->halo_lo > 0 && !lo) return fail(TDX_ERR_INVALID, std::string(where) + ": this shard needs halo_lo");
    if (ctx->halo_hi > 0 && !hi) return fail(TDX_ERR_INVALID, std::string(where) + ": this shard needs halo_hi");
    return TDX_OK;
}

static StepHost make_step(const tdx_ctx* ctx, const tdx_step_args* a) {
    StepHost h{};
    nordsieck_host(ctx->n - 1, a->dt, h.p, h.pinv);
    std::memcpy(h.ql, ctx->ql_packed, sizeof(h.ql));
    h.dt = a->dt;
    h.abstol = a->abstol;
    h.reltol = a->reltol;
    h.want_norm = (a->abstol != 0.0 || a->reltol != 0.0) ? 1 : 0;
    h.flags = a->flags;
    return h;
}

static void edge_counts(const Geo& g, int& n_first, int& n_last) {
    n_first = n_last = 0;
    switch (g.kind) {
        case IVP_FHN2D: n_first = n_last = g.cols; break;
        case IVP_BRUSS: n_first = n_last = 1; break;
        case IVP_LORENZ96: n_first = 1; n_last = 2; break;
        default: break;
    }
}

// where this shard's edge state goes (parity-0 buffers / flags; the kernels add the parity of the sequence number)
static PackPeer<double> pack_descriptor(const tdx_ctx* ctx) {
    PackPeer<double> pk{};
    edge_counts(ctx->geo, pk.n_first, pk.n_last);
    if (ctx->lower >= 0) {   // our first rows are the lower neighbour's UPPER halo
        unsigned char* base = ctx->comm_peer[ctx->lower];
        pk.dst_lower = reinterpret_cast<double*>(comm_halo(ctx, base, 1, 0));
        pk.flag_lower = &reinterpret_cast<CommHeader*>(base)->halo_flag_hi[0];
    }
    if (ctx->upper >= 0) {
        unsigned char* base = ctx->comm_peer[ctx->upper];
        pk.dst_upper = reinterpret_cast<double*>(comm_halo(ctx, base, 0, 0));
        pk.flag_upper = &reinterpret_cast<CommHeader*>(base)->halo_flag_lo[0];
    }
    pk.done_count = reinterpret_cast<CommHeader*>(ctx->comm_local)->pack_count;
    pk.seq = ctx->halo_seq;
    pk.parity_stride = (long long)((size_t)ctx->halo_cap * esize(ctx->dtype));     // bytes (Launch::typed_pack converts)
    pk.enabled = 1;
    return pk;
}

static bool uses_comm_halos(const tdx_ctx* ctx, unsigned flags) {
    return (flags & (TDX_STEP_COMM_HALO | TDX_STEP_COMM_PACK)) && ctx->comm_local;
}

static Workspace make_ws(const tdx_ctx* ctx, unsigned flags = 0u) {
    const bool second = (flags & 6u) == TDX_STEP_TILES_RIM;
    Workspace ws{};
    ws.partials = ctx->partials + (second ? ctx->partials_cap : 0);
    ws.ticket = ctx->ticket + (second ? 1 : 0);
    ws.tile_counter = ctx->ticket + 4 + (second ? 1 : 0);
    ws.mid = ctx->mid;
    ws.reserve_ctas = 0;
    ws.has_lo = ctx->halo_lo > 0;
    ws.has_hi = ctx->halo_hi > 0;
    ws.wait = CommWait{nullptr, nullptr, 0ull, nullptr, ctx->timeout_ns};
    ws.unit_a = ctx->units;
    ws.unit_b = ctx->units ? ctx->units + ctx->unit_cap : nullptr;
    ws.unit_cap = ctx->unit_cap;
    ws.cl_flag = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(ctx->mid) + 136);
    ws.cl_seq = ctx->cl_seq;
    if ((flags & TDX_STEP_COMM_PACK) && ctx->comm_local) ws.pack = pack_descriptor(ctx);
    if ((flags & TDX_STEP_COMM_REDUCE) && ctx->comm_local) {
        ws.reduce.local = reinterpret_cast<CommHeader*>(ctx->comm_local);
        for (int r = 0; r < ctx->world; ++r) ws.reduce.peer[r] = reinterpret_cast<CommHeader*>(ctx->comm_peer[r]);
        ws.reduce.rank = ctx->rank;
        ws.reduce.world = ctx->world;
        ws.reduce.seq = ctx->red_seq;      // the caller has already advanced it for this launch
        ws.reduce.timeout_ns = ctx->timeout_ns;
    }
    if (uses_comm_halos(ctx, flags)) {
        ws.halo_parity_stride_bytes = (long long)((size_t)ctx->halo_cap * esize(ctx->dtype));
        if ((flags & 6u) != TDX_STEP_TILES_INTERIOR) {
            CommHeader* hdr = reinterpret_cast<CommHeader*>(ctx->comm_local);
            ws.wait.flag_lo = ctx->halo_lo > 0 ? &hdr->halo_flag_lo[0] : nullptr;
            ws.wait.flag_hi = ctx->halo_hi > 0 ? &hdr->halo_flag_hi[0] : nullptr;
            ws.wait.error = &hdr->error;
        }
    }
    ws.wait.seq = ctx->halo_seq;
    return ws;
}

// halo pointers of a step call: the caller's, or (TDX_STEP_COMM_HALO / _PACK) the context's peer-written buffers of parity 0
static void resolve_halos(const tdx_ctx* ctx, unsigned flags, const void*& lo, const void*& hi) {
    if (!uses_comm_halos(ctx, flags)) return;
    lo = ctx->halo_lo > 0 ? comm_halo(ctx, ctx->comm_local, 0, 0) : nullptr;
    hi = ctx->halo_hi > 0 ? comm_halo(ctx, ctx->comm_local, 1, 0) : nullptr;
}

// A launch that delivers the halos itself (TDX_STEP_COMM_PACK) starts a new halo sequence number.  Split into a check (no
// side effects; call it with the other argument checks) and the advance (immediately before the launch), so that a rank
// whose arguments are rejected does not get out of step with its neighbours.
static int comm_pack_check(const tdx_ctx* ctx, unsigned flags, const char* where) {
    if (!(flags & TDX_STEP_COMM_PACK)) return TDX_OK;
    if (!ctx->comm_local) return fail(TDX_ERR_INVALID, std::string(where) + ": TDX_STEP_COMM_PACK needs tdx_comm_init");
    if ((flags & 6u) == TDX_STEP_TILES_INTERIOR || (flags & 6u) == TDX_STEP_TILES_RIM)
        return fail(TDX_ERR_INVALID, std::string(where) + ": TDX_STEP_COMM_PACK cannot be combined with a partial tile range");
    if ((ctx->lower >= 0 && !ctx->comm_peer[ctx->lower]) || (ctx->upper >= 0 && !ctx->comm_peer[ctx->upper]))
        return fail(TDX_ERR_INVALID, std::string(where) + ": neighbours are not connected");
    return TDX_OK;
}
static void comm_pack_advance(tdx_ctx* ctx, unsigned flags, unsigned long long count = 1ull) {
    if (flags & TDX_STEP_COMM_PACK) ctx->halo_seq += count;
}

// A launch that all-reduces its scalar in the epilogue consumes reduction sequence numbers (same split).
static int comm_reduce_check(const tdx_ctx* ctx, unsigned flags, bool produces_sum, const char* where) {
    if (!(flags & TDX_STEP_COMM_REDUCE) || !produces_sum) return TDX_OK;
    if (!ctx->comm_local) return fail(TDX_ERR_INVALID, std::string(where) + ": TDX_STEP_COMM_REDUCE needs tdx_comm_init");
    if ((flags & 6u) == TDX_STEP_TILES_INTERIOR || (flags & 6u) == TDX_STEP_TILES_RIM)
        return fail(TDX_ERR_INVALID, std::string(where) + ": TDX_STEP_COMM_REDUCE cannot be combined with a partial tile range");
    for (int r = 0; r < ctx->world; ++r)
        if (!ctx->comm_peer[r]) return fail(TDX_ERR_INVALID, std::string(where) + ": not all peers are connected");
    return TDX_OK;
}
static void comm_reduce_advance(tdx_ctx* ctx, unsigned flags, bool produces_sum, unsigned long long count = 1ull) {
    if ((flags & TDX_STEP_COMM_REDUCE) && produces_sum) ctx->red_seq += count;
}

static int check_dt(const tdx_step_args* a, const char* where) {
    if (!a) return fail(TDX_ERR_INVALID, std::string(where) + ": null step arguments");
    if (!(a->dt > 0.0) || !std::isfinite(a->dt))
        return fail(TDX_ERR_INVALID, std::string(where) + ": Invalid step size: dt must be positive and finite");   // odefilter.py:221
    return TDX_OK;
}

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

extern "C" {

int tdx_eval_f(tdx_ctx* ctx, double t, const void* y, void* fy, void* jdiag, const void* halo_lo, const void* halo_hi,
               void* stream) {
    (void)t;
    int rc = check_ready(ctx, "tdx_eval_f");
    if (rc) return rc;
    if (!y || !fy) return fail(TDX_ERR_INVALID, "tdx_eval_f: null buffer");
    if (ctx->geo.kind == IVP_EXTERNAL) return fail(TDX_ERR_UNSUPPORTED, "tdx_eval_f: an external vector field is evaluated by the caller");
    if ((rc = check_halos(ctx, halo_lo, halo_hi, "tdx_eval_f"))) return rc;
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    return cuda_fail(ctx->eval_ops->eval_f(ctx->geo, y, fy, jdiag, halo_lo, halo_hi, (cudaStream_t)stream), "tdx_eval_f");
}

int tdx_pack_halo(tdx_ctx* ctx, double dt, const void* mean_in, void* edge_first, void* edge_last, void* stream) {
    int rc = check_ready(ctx, "tdx_pack_halo");
    if (rc) return rc;
    tdx_step_args a{0.0, dt, 0.0, 0.0, 0u, 0u};
    if ((rc = check_dt(&a, "tdx_pack_halo"))) return rc;
    if (!mean_in) return fail(TDX_ERR_INVALID, "tdx_pack_halo: null buffer");
    // what the neighbours need from us: the lower neighbour's hi-halo is our first row/points, the upper
    // neighbour's lo-halo is our last row/points.
    int n_first = 0, n_last = 0;
    const Geo& g = ctx->geo;
    switch (g.kind) {
        case IVP_FHN2D: n_first = n_last = g.cols; break;
        case IVP_BRUSS: n_first = n_last = 1; break;
        case IVP_LORENZ96: n_first = 1; n_last = 2; break;
        default: break;
    }
    if (!edge_first) n_first = 0;
    if (!edge_last) n_last = 0;
    if (n_first + n_last == 0) return TDX_OK;
    StepHost h = make_step(ctx, &a);
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    return cuda_fail(ctx->ops->pack(g, h, mean_in, edge_first, edge_last, n_first, n_last, (cudaStream_t)stream),
                     "tdx_pack_halo");
}

int tdx_predict_state(tdx_ctx* ctx, double dt, const void* mean_in, void* y_out, void* stream) {
    int rc = check_ready(ctx, "tdx_predict_state");
    if (rc) return rc;
    tdx_step_args a{0.0, dt, 0.0, 0.0, 0u, 0u};
    if ((rc = check_dt(&a, "tdx_predict_state"))) return rc;
    if (!mean_in || !y_out) return fail(TDX_ERR_INVALID, "tdx_predict_state: null buffer");
    const Geo& g = ctx->geo;
    if (g.npts > 2147483647LL / 2) return fail(TDX_ERR_INVALID, "tdx_predict_state: shard too large");
    StepHost h = make_step(ctx, &a);
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    // the "first n points of every field" of the halo packer, with n = all of them, is the predicted state of the shard
    return cuda_fail(ctx->ops->pack(g, h, mean_in, y_out, nullptr, (int)g.npts, 0, (cudaStream_t)stream), "tdx_predict_state");
}

int tdx_dek1_attempt_step(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in, const void* sc_in,
                          void* mean_out, void* sc_out, void* err_out, void* ref_out, const void* halo_lo,
                          const void* halo_hi, double* sq_norm_sum, void* stream) {
    int rc = check_ready(ctx, "tdx_dek1_attempt_step");
    if (rc) return rc;
    if ((rc = check_dt(args, "tdx_dek1_attempt_step"))) return rc;
    if ((rc = check_external(ctx, "tdx_dek1_attempt_step"))) return rc;
    if ((rc = comm_pack_check(ctx, args->flags, "tdx_dek1_attempt_step"))) return rc;
    resolve_halos(ctx, args->flags, halo_lo, halo_hi);
    if (!mean_in || !sc_in || !mean_out || !sc_out) return fail(TDX_ERR_INVALID, "tdx_dek1_attempt_step: null state buffer");
    if (mean_in == mean_out || sc_in == sc_out)
        return fail(TDX_ERR_INVALID, "tdx_dek1_attempt_step: in and out buffers must be distinct (a rejected attempt keeps the input state)");
    if ((rc = check_halos(ctx, halo_lo, halo_hi, "tdx_dek1_attempt_step", args->flags))) return rc;
    StepHost h = make_step(ctx, args);
    if (h.want_norm && !sq_norm_sum) return fail(TDX_ERR_INVALID, "tdx_dek1_attempt_step: sq_norm_sum is required with tolerances");
    if ((rc = comm_reduce_check(ctx, args->flags, h.want_norm != 0, "tdx_dek1_attempt_step"))) return rc;
    // every argument is valid: only now do the exchange's sequence numbers move (all ranks stay in step)
    comm_pack_advance(ctx, args->flags);
    comm_reduce_advance(ctx, args->flags, h.want_norm != 0);
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    return cuda_fail(ctx->ops->dek1(ctx->geo, h, mean_in, sc_in, mean_out, sc_out, err_out, ref_out, halo_lo, halo_hi,
                                    make_ws(ctx, args->flags), sq_norm_sum, (cudaStream_t)stream),
                     "tdx_dek1_attempt_step");
}

int tdx_ek0_sweep_a(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in, const void* halo_lo,
                    const void* halo_hi, double* z_sq_sum, void* stream) {
    int rc = check_ready(ctx, "tdx_ek0_sweep_a");
    if (rc) return rc;
    if ((rc = check_dt(args, "tdx_ek0_sweep_a"))) return rc;
    if ((rc = comm_pack_check(ctx, args->flags, "tdx_ek0_sweep_a"))) return rc;
    resolve_halos(ctx, args->flags, halo_lo, halo_hi);
    if (!mean_in || !z_sq_sum) return fail(TDX_ERR_INVALID, "tdx_ek0_sweep_a: null buffer");
    if ((rc = check_halos(ctx, halo_lo, halo_hi, "tdx_ek0_sweep_a", args->flags))) return rc;
    StepHost h = make_step(ctx, args);
    if ((rc = comm_reduce_check(ctx, args->flags, true, "tdx_ek0_sweep_a"))) return rc;
    comm_pack_advance(ctx, args->flags);
    comm_reduce_advance(ctx, args->flags, true);
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    return cuda_fail(ctx->ops->ek0_a(ctx->geo, h, mean_in, halo_lo, halo_hi, make_ws(ctx, args->flags), z_sq_sum, (cudaStream_t)stream),
                     "tdx_ek0_sweep_a");
}

int tdx_ek0_mid(tdx_ctx* ctx, const tdx_step_args* args, const void* Cl_in, const double* z_sq_sum, int64_t d_global,
                void* Cl_out, void* err_out, void* stream) {
    int rc = check_ready(ctx, "tdx_ek0_mid");
    if (rc) return rc;
    if ((rc = check_dt(args, "tdx_ek0_mid"))) return rc;
    if (!Cl_in || !z_sq_sum || !Cl_out || !err_out) return fail(TDX_ERR_INVALID, "tdx_ek0_mid: null buffer");
    if (d_global < 1) return fail(TDX_ERR_INVALID, "tdx_ek0_mid: d_global must be positive");
    StepHost h = make_step(ctx, args);
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    return cuda_fail(ctx->ops->ek0_mid(h, Cl_in, z_sq_sum, (double)d_global, Cl_out, err_out, make_ws(ctx), (cudaStream_t)stream),
                     "tdx_ek0_mid");
}

int tdx_ek0_sweep_b(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in, void* mean_out, void* ref_out,
                    const void* halo_lo, const void* halo_hi, double* sq_norm_sum, void* stream) {
    int rc = check_ready(ctx, "tdx_ek0_sweep_b");
    if (rc) return rc;
    if ((rc = check_dt(args, "tdx_ek0_sweep_b"))) return rc;
    resolve_halos(ctx, args->flags, halo_lo, halo_hi);
    if (!mean_in || !mean_out) return fail(TDX_ERR_INVALID, "tdx_ek0_sweep_b: null buffer");
    if (mean_in == mean_out) return fail(TDX_ERR_INVALID, "tdx_ek0_sweep_b: in and out buffers must be distinct");
    if ((rc = check_halos(ctx, halo_lo, halo_hi, "tdx_ek0_sweep_b"))) return rc;
    StepHost h = make_step(ctx, args);
    if (h.want_norm && !sq_norm_sum) return fail(TDX_ERR_INVALID, "tdx_ek0_sweep_b: sq_norm_sum is required with tolerances");
    if ((rc = comm_reduce_check(ctx, args->flags, h.want_norm != 0, "tdx_ek0_sweep_b"))) return rc;
    comm_reduce_advance(ctx, args->flags, h.want_norm != 0);
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    return cuda_fail(ctx->ops->ek0_b(ctx->geo, h, mean_in, mean_out, ref_out, halo_lo, halo_hi, make_ws(ctx, args->flags & ~6u), sq_norm_sum,
                                     (cudaStream_t)stream),
                     "tdx_ek0_sweep_b");
}

// ---- peer-memory exchange ---------------------------------------------------------------------------------

int tdx_comm_init(tdx_ctx* ctx, int rank, int world, int lower_rank, int upper_rank, void* handle_out) {
    int rc = check_ready(ctx, "tdx_comm_init");
    if (rc) return rc;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || !handle_out)
        return fail(TDX_ERR_INVALID, "tdx_comm_init: need 2 <= world <= 16, 0 <= rank < world and a handle buffer");
    if (lower_rank >= world || upper_rank >= world) return fail(TDX_ERR_INVALID, "tdx_comm_init: neighbour rank out of range");
    if (ctx->comm_local) return fail(TDX_ERR_INVALID, "tdx_comm_init: already initialised");
    DeviceGuard guard(ctx->device);
    ctx->rank = rank; ctx->world = world; ctx->lower = lower_rank; ctx->upper = upper_rank;
    // same capacity on every rank so that the block layout is identical everywhere
    ctx->halo_cap = (ctx->geo.kind == IVP_FHN2D) ? 2LL * ctx->geo.cols : 4;
    cudaError_t e = cudaMalloc(&ctx->comm_local, comm_bytes(ctx));
    if (e == cudaSuccess) e = cudaMemset(ctx->comm_local, 0, comm_bytes(ctx));
    if (e == cudaSuccess) {
        CommHeader init{};
        init.wait_timeout_ns = ctx->timeout_ns;
        e = cudaMemcpy(ctx->comm_local, &init, sizeof(init), cudaMemcpyHostToDevice);
    }
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ctx->comm_local);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return cuda_fail((int)e, "tdx_comm_init");
    static_assert(sizeof(cudaIpcMemHandle_t) == TDX_COMM_HANDLE_BYTES, "IPC handle size");
    std::memcpy(handle_out, &h, sizeof(h));
    ctx->comm_peer[rank] = ctx->comm_local;
    ctx->halo_seq = 0; ctx->red_seq = 0;
    return TDX_OK;
}

int tdx_comm_connect(tdx_ctx* ctx, int peer, const void* handle) {
    if (!ctx || !ctx->comm_local) return fail(TDX_ERR_INVALID, "tdx_comm_connect: call tdx_comm_init first");
    if (peer < 0 || peer >= ctx->world || !handle) return fail(TDX_ERR_INVALID, "tdx_comm_connect: bad peer or handle");
    if (peer == ctx->rank) return TDX_OK;
    if (ctx->comm_opened[peer]) return TDX_OK;
    DeviceGuard guard(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail((int)e, "tdx_comm_connect (cudaIpcOpenMemHandle)");
    ctx->comm_peer[peer] = (unsigned char*)ptr;
    ctx->comm_opened[peer] = true;
    return TDX_OK;
}

int tdx_comm_pack_halo(tdx_ctx* ctx, double dt, const void* mean_in, void* stream) {
    int rc = check_ready(ctx, "tdx_comm_pack_halo");
    if (rc) return rc;
    if (!ctx->comm_local) return fail(TDX_ERR_INVALID, "tdx_comm_pack_halo: call tdx_comm_init first");
    tdx_step_args a{0.0, dt, 0.0, 0.0, 0u, 0u};
    if ((rc = check_dt(&a, "tdx_comm_pack_halo"))) return rc;
    if (!mean_in) return fail(TDX_ERR_INVALID, "tdx_comm_pack_halo: null buffer");
    if ((ctx->lower >= 0 && !ctx->comm_peer[ctx->lower]) || (ctx->upper >= 0 && !ctx->comm_peer[ctx->upper]))
        return fail(TDX_ERR_INVALID, "tdx_comm_pack_halo: neighbours are not connected");
    ctx->halo_seq += 1;
    PackPeer<double> pk = pack_descriptor(ctx);
    StepHost h = make_step(ctx, &a);
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    return cuda_fail(ctx->ops->pack_peer(ctx->geo, h, mean_in, pk, (cudaStream_t)stream), "tdx_comm_pack_halo");
}

int tdx_comm_allreduce(tdx_ctx* ctx, double* vals, int count, void* stream) {
    if (!ctx || !ctx->comm_local) return fail(TDX_ERR_INVALID, "tdx_comm_allreduce: call tdx_comm_init first");
    if (!vals || count < 1 || count > 4) return fail(TDX_ERR_INVALID, "tdx_comm_allreduce: 1 <= count <= 4");
    AllReduceArgs a{};
    a.local = reinterpret_cast<CommHeader*>(ctx->comm_local);
    for (int r = 0; r < ctx->world; ++r) {
        if (!ctx->comm_peer[r]) return fail(TDX_ERR_INVALID, "tdx_comm_allreduce: not all peers are connected");
        a.peer[r] = reinterpret_cast<CommHeader*>(ctx->comm_peer[r]);
    }
    a.vals = vals; a.count = count; a.rank = ctx->rank; a.world = ctx->world;
    ctx->red_seq += 1;
    a.seq = ctx->red_seq;
    a.timeout_ns = ctx->timeout_ns;
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    comm_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    return cuda_fail((int)cudaGetLastError(), "tdx_comm_allreduce");
}

int tdx_comm_error(tdx_ctx* ctx, int* error_out, void* stream) {
    if (!ctx || !ctx->comm_local || !error_out) return fail(TDX_ERR_INVALID, "tdx_comm_error: no comm block");
    DeviceGuard guard(ctx->device);
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, &reinterpret_cast<CommHeader*>(ctx->comm_local)->error, sizeof(v),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail((int)e, "tdx_comm_error");
    *error_out = (int)v;
    return TDX_OK;
}

static bool three_phase_forced() {
    static const bool on = std::getenv("TDX_EK0_THREE_PHASE") != nullptr;   // experiments / A-B comparison only
    return on;
}

int tdx_ek0_attempt_step(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in, const void* Cl_in,
                         void* mean_out, void* Cl_out, void* err_out, void* ref_out, double* z_sq_sum,
                         double* sq_norm_sum, void* stream) {
    int rc = check_ready(ctx, "tdx_ek0_attempt_step");
    if (rc) return rc;
    if ((rc = check_dt(args, "tdx_ek0_attempt_step"))) return rc;
    if ((rc = check_external(ctx, "tdx_ek0_attempt_step"))) return rc;
    const bool sharded = ctx->halo_lo > 0 || ctx->halo_hi > 0;
    const unsigned comm_flags = TDX_STEP_COMM_PACK | TDX_STEP_COMM_REDUCE;
    if (sharded && (args->flags & comm_flags) != comm_flags)
        return fail(TDX_ERR_INVALID,
                    "tdx_ek0_attempt_step: a sharded context needs TDX_STEP_COMM_PACK | TDX_STEP_COMM_REDUCE (peer-memory "
                    "exchange), or the three phases with the caller's own all-reduce in between");
    if (args->flags & 6u) return fail(TDX_ERR_INVALID, "tdx_ek0_attempt_step: partial tile ranges are not supported");
    if (!z_sq_sum) return fail(TDX_ERR_INVALID, "tdx_ek0_attempt_step: z_sq_sum is required");
    const int64_t d_global = (int64_t)ctx->geo.nf * ctx->geo.grows * ctx->geo.gcols;
    if (three_phase_forced()) {
        tdx_step_args a = *args, b = *args;
        if (!sharded) a.flags = b.flags = args->flags & TDX_STEP_ZERO_JACOBIAN;
        else b.flags = (args->flags & ~TDX_STEP_COMM_PACK) | TDX_STEP_COMM_HALO;
        if ((rc = tdx_ek0_sweep_a(ctx, &a, mean_in, nullptr, nullptr, z_sq_sum, stream))) return rc;
        if ((rc = tdx_ek0_mid(ctx, args, Cl_in, z_sq_sum, d_global, Cl_out, err_out, stream))) return rc;
        return tdx_ek0_sweep_b(ctx, &b, mean_in, mean_out, ref_out, nullptr, nullptr, sq_norm_sum, stream);
    }
    // the whole attempt is one cooperative launch (fhn_2d: row-streaming strips; 1-d fields: tile ranges)
    const unsigned flags = sharded ? args->flags : 0u;
    if ((rc = comm_pack_check(ctx, flags, "tdx_ek0_attempt_step"))) return rc;
    const void* halo_lo = nullptr;
    const void* halo_hi = nullptr;
    resolve_halos(ctx, flags, halo_lo, halo_hi);
    if (!mean_in || !mean_out || !Cl_in || !Cl_out || !err_out) return fail(TDX_ERR_INVALID, "tdx_ek0_attempt_step: null buffer");
    if (mean_in == mean_out) return fail(TDX_ERR_INVALID, "tdx_ek0_attempt_step: in and out buffers must be distinct");
    StepHost h = make_step(ctx, args);
    if (h.want_norm && !sq_norm_sum) return fail(TDX_ERR_INVALID, "tdx_ek0_attempt_step: sq_norm_sum is required with tolerances");
    if ((rc = comm_reduce_check(ctx, flags, true, "tdx_ek0_attempt_step"))) return rc;
    comm_pack_advance(ctx, flags);
    comm_reduce_advance(ctx, flags, true);                   // sum of z^2 ...
    Workspace ws = make_ws(ctx, flags);
    comm_reduce_advance(ctx, flags, h.want_norm != 0);       // ... and the norm sum
    CommReduce reduce_b = ws.reduce;
    reduce_b.seq = ctx->red_seq;
    ctx->mid_seq += 1;
    ctx->cl_seq += 1;                                        // the launch releases cl_seq(base) + 1
    DeviceGuard guard(ctx->device);
    ctx->launches += 1;
    unsigned long long* mid_flag = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(ctx->mid) + 128);
    return cuda_fail(ctx->ops->ek0_fused(ctx->geo, h, mean_in, Cl_in, mean_out, Cl_out, err_out, ref_out, halo_lo, halo_hi, ws,
                                         reduce_b, mid_flag, ctx->mid_seq, (double)d_global, z_sq_sum, sq_norm_sum,
                                         (cudaStream_t)stream),
                     "tdx_ek0_attempt_step");
}

}  // extern "C"


// =====================================================================================================================
// Host-buffer entry point: DiagonalEK1.attempt_step on a state that lives in HOST memory (what a caller without
// device arrays hands over, e.g. the reference's NumPy/JAX-CPU arrays).  The state is cut into chunks of grid rows
// (fhn_2d) / chain points; three streams run  H2D(chunk c+1) | step kernel(chunk c) | D2H(chunk c-1)  concurrently, so
// the call costs about max(H2D, D2H) of the state over PCIe instead of H2D + kernel + D2H.  Every chunk is an
// ordinary sharded context: the stencil's halo values come from tdx_pack_halo on the neighbouring chunks' means.
// =====================================================================================================================
struct HostPipe {
    int nchunks = 0;
    std::vector<tdx_ctx*> ctx;            // one sharded context per chunk
    std::vector<long long> p0, p1;        // owned points [p0, p1) of every chunk (per field)
    std::vector<cudaEvent_t> ev_mean, ev_fac, ev_kernel, ev_d2h;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    unsigned char *mean_in = nullptr, *mean_out = nullptr, *err = nullptr, *ref = nullptr;   // all chunks, compact per chunk
    unsigned char *fac_in[3] = {}, *fac_out[3] = {};                                          // rings over the chunks
    unsigned char *edge_first = nullptr, *edge_last = nullptr;                                // [nchunks][edge_cap]
    double* sq = nullptr;                 // [nchunks] partial norm sums (device)
    double* sq_host = nullptr;            // pinned
    long long edge_cap = 0;
    static constexpr int kSlots = 3;
};

static void destroy_pipe(tdx_ctx* ctx) {
    HostPipe* hp = ctx->pipe;
    if (!hp) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    for (tdx_ctx* c : hp->ctx) tdx_destroy(c);
    for (auto& v : {&hp->ev_mean, &hp->ev_fac, &hp->ev_kernel, &hp->ev_d2h})
        for (cudaEvent_t e : *v) cudaEventDestroy(e);
    if (hp->s_h2d) cudaStreamDestroy(hp->s_h2d);
    if (hp->s_comp) cudaStreamDestroy(hp->s_comp);
    if (hp->s_d2h) cudaStreamDestroy(hp->s_d2h);
    cudaFree(hp->mean_in); cudaFree(hp->mean_out); cudaFree(hp->err); cudaFree(hp->ref);
    for (int i = 0; i < HostPipe::kSlots; ++i) { cudaFree(hp->fac_in[i]); cudaFree(hp->fac_out[i]); }
    cudaFree(hp->edge_first); cudaFree(hp->edge_last); cudaFree(hp->sq);
    if (hp->sq_host) cudaFreeHost(hp->sq_host);
    cudaSetDevice(prev);
    delete hp;
    ctx->pipe = nullptr;
}

static int build_pipe(tdx_ctx* ctx, int nchunks_hint) {
    const Geo& g = ctx->geo;
    const bool rows_axis = g.kind == IVP_FHN2D;
    const long long extent = rows_axis ? g.rows : g.cols;
    long long min_chunk = rows_axis ? 8 : 4096;
    int n = nchunks_hint > 0 ? nchunks_hint : 16;
    if (g.kind == IVP_VDP) n = 1;
    while (n > 1 && extent / n < min_chunk) --n;
    HostPipe* hp = new HostPipe();
    hp->nchunks = n;
    const size_t w = esize(ctx->dtype);
    const long long nn = (long long)ctx->n * ctx->n;
    long long max_pts = 0;
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < n && e == cudaSuccess; ++c) {
        const long long b = extent * c / n, en = extent * (c + 1) / n;
        tdx_ctx* cc = nullptr;
        int rc = tdx_create(&cc, ctx->device, ctx->dtype);
        tdx_problem_desc d = ctx->desc;
        d.shard_begin = ctx->desc.shard_begin + b;     // sub-shards of the parent's own shard
        d.shard_end = ctx->desc.shard_begin + en;
        if (!rc) rc = tdx_set_problem(cc, &d);
        if (rc) {
            tdx_destroy(cc);
            for (tdx_ctx* made : hp->ctx) tdx_destroy(made);
            delete hp;
            return rc;
        }
        hp->ctx.push_back(cc);
        const long long per_unit = rows_axis ? g.cols : 1;
        hp->p0.push_back(b * per_unit);
        hp->p1.push_back(en * per_unit);
        max_pts = std::max(max_pts, (en - b) * per_unit);
    }
    int nf = 0, nl = 0;
    edge_counts(g, nf, nl);
    hp->edge_cap = (long long)g.nf * std::max(nf, nl) + 8;
    auto alloc = [&](unsigned char** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 16); };
    alloc(&hp->mean_in, (size_t)ctx->n * g.d * w);
    alloc(&hp->mean_out, (size_t)ctx->n * g.d * w);
    alloc(&hp->err, (size_t)g.d * w);
    alloc(&hp->ref, (size_t)g.d * w);
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        alloc(&hp->fac_in[i], (size_t)g.nf * max_pts * nn * w);
        alloc(&hp->fac_out[i], (size_t)g.nf * max_pts * nn * w);
    }
    alloc(&hp->edge_first, (size_t)n * hp->edge_cap * w);
    alloc(&hp->edge_last, (size_t)n * hp->edge_cap * w);
    alloc((unsigned char**)&hp->sq, (size_t)n * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(&hp->sq_host, (size_t)n * sizeof(double));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_comp, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_d2h, cudaStreamNonBlocking);
    for (auto* v : {&hp->ev_mean, &hp->ev_fac, &hp->ev_kernel, &hp->ev_d2h})
        for (int c = 0; c < n && e == cudaSuccess; ++c) {
            cudaEvent_t ev;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e == cudaSuccess) v->push_back(ev);
        }
    ctx->pipe = hp;
    if (e != cudaSuccess) {
        destroy_pipe(ctx);
        return cuda_fail((int)e, "tdx_dek1_attempt_step_host (pipeline setup)");
    }
    return TDX_OK;
}

static int host_pipeline(tdx_ctx* ctx, const tdx_step_args* args, const void* h_mean_in, const void* h_sc_in, void* h_mean_out,
                         void* h_sc_out, void* h_err_out, void* h_ref_out, double* h_sq_norm_sum, int nchunks,
                         const void* ext_halo_lo, const void* ext_halo_hi);

extern "C" int tdx_dek1_attempt_step_host(tdx_ctx* ctx, const tdx_step_args* args, const void* h_mean_in,
                                          const void* h_sc_in, void* h_mean_out, void* h_sc_out, void* h_err_out,
                                          void* h_ref_out, double* h_sq_norm_sum, int nchunks) {
    int rc = check_ready(ctx, "tdx_dek1_attempt_step_host");
    if (rc) return rc;
    if (ctx->halo_lo > 0 || ctx->halo_hi > 0)
        return fail(TDX_ERR_INVALID, "tdx_dek1_attempt_step_host: the context must own the whole problem; a shard needs "
                                     "tdx_dek1_attempt_step_host_sharded (halo buffers from its neighbours)");
    return host_pipeline(ctx, args, h_mean_in, h_sc_in, h_mean_out, h_sc_out, h_err_out, h_ref_out, h_sq_norm_sum, nchunks, nullptr,
                         nullptr);
}

extern "C" int tdx_dek1_attempt_step_host_sharded(tdx_ctx* ctx, const tdx_step_args* args, const void* h_mean_in,
                                                  const void* h_sc_in, void* h_mean_out, void* h_sc_out, void* h_err_out,
                                                  void* h_ref_out, const void* d_halo_lo, const void* d_halo_hi,
                                                  double* h_sq_norm_sum, int nchunks) {
    int rc = check_ready(ctx, "tdx_dek1_attempt_step_host_sharded");
    if (rc) return rc;
    if ((rc = check_halos(ctx, d_halo_lo, d_halo_hi, "tdx_dek1_attempt_step_host_sharded"))) return rc;
    return host_pipeline(ctx, args, h_mean_in, h_sc_in, h_mean_out, h_sc_out, h_err_out, h_ref_out, h_sq_norm_sum, nchunks, d_halo_lo,
                         d_halo_hi);
}

static int host_pipeline(tdx_ctx* ctx, const tdx_step_args* args, const void* h_mean_in, const void* h_sc_in, void* h_mean_out,
                         void* h_sc_out, void* h_err_out, void* h_ref_out, double* h_sq_norm_sum, int nchunks,
                         const void* ext_halo_lo, const void* ext_halo_hi) {
    int rc = check_dt(args, "tdx_dek1_attempt_step_host");
    if (rc) return rc;
    if (!h_mean_in || !h_sc_in || !h_mean_out || !h_sc_out) return fail(TDX_ERR_INVALID, "tdx_dek1_attempt_step_host: null state buffer");
    if (args->flags & ~TDX_STEP_ZERO_JACOBIAN) return fail(TDX_ERR_INVALID, "tdx_dek1_attempt_step_host: only TDX_STEP_ZERO_JACOBIAN is accepted");
    const bool want_norm = args->abstol != 0.0 || args->reltol != 0.0;
    if (want_norm && !h_sq_norm_sum) return fail(TDX_ERR_INVALID, "tdx_dek1_attempt_step_host: sq_norm_sum is required with tolerances");
    DeviceGuard guard(ctx->device);
    if (ctx->pipe && nchunks > 0 && ctx->pipe->nchunks != nchunks) destroy_pipe(ctx);
    if (!ctx->pipe && (rc = build_pipe(ctx, nchunks))) return rc;
    HostPipe& hp = *ctx->pipe;
    const Geo& g = ctx->geo;
    const int n = ctx->n, nc = hp.nchunks;
    const size_t w = esize(ctx->dtype);
    const long long nn = (long long)n * n, npts = g.npts;
    const bool parent_sharded = ctx->halo_lo > 0 || ctx->halo_hi > 0;
    const bool cyclic = g.kind == IVP_LORENZ96 && nc > 1 && !parent_sharded;    // a shard's ends talk to other ranks, not to each other
    for (tdx_ctx* c : hp.ctx) std::memcpy(c->ql_packed, ctx->ql_packed, sizeof(ctx->ql_packed));   // tdx_set_diffusion_sqrt of the parent

    const unsigned char* hm_in = (const unsigned char*)h_mean_in;
    const unsigned char* hs_in = (const unsigned char*)h_sc_in;
    unsigned char* hm_out = (unsigned char*)h_mean_out;
    unsigned char* hs_out = (unsigned char*)h_sc_out;
    auto dmean = [&](unsigned char* base, int c) { return base + (size_t)n * g.nf * hp.p0[c] * w; };   // compact (n, d_c) block
    auto dvec = [&](unsigned char* base, int c) { return base + (size_t)g.nf * hp.p0[c] * w; };
    auto nb = [&](int c) -> int { return cyclic ? (c + nc) % nc : ((c < 0 || c >= nc) ? -1 : c); };
    std::vector<char> mean_up(nc, 0), packed(nc, 0);
    cudaError_t e = cudaSuccess;
#define TDX_PIPE_TRY(expr) do { if (e == cudaSuccess) e = (expr); } while (0)

    auto upload_mean = [&](int c) {
        if (c < 0 || mean_up[c]) return;
        const long long np = hp.p1[c] - hp.p0[c];
        for (int f = 0; f < g.nf; ++f)   // rows = derivatives: (n) x (np) block of the (n, d) host array
            TDX_PIPE_TRY(cudaMemcpy2DAsync(dmean(hp.mean_in, c) + (size_t)f * np * w, (size_t)g.nf * np * w,
                                           hm_in + ((size_t)f * npts + hp.p0[c]) * w, (size_t)g.d * w, (size_t)np * w, n,
                                           cudaMemcpyHostToDevice, hp.s_h2d));
        TDX_PIPE_TRY(cudaEventRecord(hp.ev_mean[c], hp.s_h2d));
        mean_up[c] = 1;
    };
    auto pack = [&](int c) {
        if (c < 0 || packed[c] || nc == 1) return;
        TDX_PIPE_TRY(cudaStreamWaitEvent(hp.s_comp, hp.ev_mean[c], 0));
        if (e == cudaSuccess && tdx_pack_halo(hp.ctx[c], args->dt, dmean(hp.mean_in, c), hp.edge_first + (size_t)c * hp.edge_cap * w,
                                              hp.edge_last + (size_t)c * hp.edge_cap * w, hp.s_comp))
            e = cudaErrorUnknown;
        packed[c] = 1;
    };

    for (int c = 0; c < nc && e == cudaSuccess; ++c) {
        const long long np = hp.p1[c] - hp.p0[c];
        const int slot = c % HostPipe::kSlots;
        const int lo = nb(c - 1), hi = nb(c + 1);
        // ---- host -> device: the means this chunk's stencil touches, then its factor blocks
        upload_mean(lo);
        upload_mean(c);
        upload_mean(hi);
        if (c >= HostPipe::kSlots) TDX_PIPE_TRY(cudaStreamWaitEvent(hp.s_h2d, hp.ev_kernel[c - HostPipe::kSlots], 0));
        for (int f = 0; f < g.nf; ++f)
            TDX_PIPE_TRY(cudaMemcpyAsync(hp.fac_in[slot] + (size_t)f * np * nn * w, hs_in + ((size_t)f * npts + hp.p0[c]) * nn * w,
                                         (size_t)np * nn * w, cudaMemcpyHostToDevice, hp.s_h2d));
        TDX_PIPE_TRY(cudaEventRecord(hp.ev_fac[c], hp.s_h2d));
        // ---- compute
        pack(lo);
        pack(c);
        pack(hi);
        TDX_PIPE_TRY(cudaStreamWaitEvent(hp.s_comp, hp.ev_fac[c], 0));
        if (c >= HostPipe::kSlots) TDX_PIPE_TRY(cudaStreamWaitEvent(hp.s_comp, hp.ev_d2h[c - HostPipe::kSlots], 0));
        // stencil neighbours beyond the chunk: the adjacent chunk's packed edge, or (first / last chunk of a shard) the halo the
        // caller received from the neighbouring rank
        const void* halo_lo = hp.ctx[c]->halo_lo > 0 ? (lo >= 0 ? hp.edge_last + (size_t)lo * hp.edge_cap * w : ext_halo_lo) : nullptr;
        const void* halo_hi = hp.ctx[c]->halo_hi > 0 ? (hi >= 0 ? hp.edge_first + (size_t)hi * hp.edge_cap * w : ext_halo_hi) : nullptr;
        if (e == cudaSuccess) {
            rc = tdx_dek1_attempt_step(hp.ctx[c], args, dmean(hp.mean_in, c), hp.fac_in[slot], dmean(hp.mean_out, c), hp.fac_out[slot],
                                       h_err_out ? dvec(hp.err, c) : nullptr, h_ref_out ? dvec(hp.ref, c) : nullptr, halo_lo, halo_hi,
                                       hp.sq + c, hp.s_comp);
            if (rc) {      // drain what is in flight before the caller gets its buffers back
                cudaStreamSynchronize(hp.s_h2d);
                cudaStreamSynchronize(hp.s_comp);
                cudaStreamSynchronize(hp.s_d2h);
                return rc;
            }
        }
        TDX_PIPE_TRY(cudaEventRecord(hp.ev_kernel[c], hp.s_comp));
        // ---- device -> host
        TDX_PIPE_TRY(cudaStreamWaitEvent(hp.s_d2h, hp.ev_kernel[c], 0));
        for (int f = 0; f < g.nf; ++f) {
            TDX_PIPE_TRY(cudaMemcpy2DAsync(hm_out + ((size_t)f * npts + hp.p0[c]) * w, (size_t)g.d * w,
                                           dmean(hp.mean_out, c) + (size_t)f * np * w, (size_t)g.nf * np * w, (size_t)np * w, n,
                                           cudaMemcpyDeviceToHost, hp.s_d2h));
            TDX_PIPE_TRY(cudaMemcpyAsync(hs_out + ((size_t)f * npts + hp.p0[c]) * nn * w, hp.fac_out[slot] + (size_t)f * np * nn * w,
                                         (size_t)np * nn * w, cudaMemcpyDeviceToHost, hp.s_d2h));
            if (h_err_out)
                TDX_PIPE_TRY(cudaMemcpyAsync((unsigned char*)h_err_out + ((size_t)f * npts + hp.p0[c]) * w, dvec(hp.err, c) + (size_t)f * np * w,
                                             (size_t)np * w, cudaMemcpyDeviceToHost, hp.s_d2h));
            if (h_ref_out)
                TDX_PIPE_TRY(cudaMemcpyAsync((unsigned char*)h_ref_out + ((size_t)f * npts + hp.p0[c]) * w, dvec(hp.ref, c) + (size_t)f * np * w,
                                             (size_t)np * w, cudaMemcpyDeviceToHost, hp.s_d2h));
        }
        TDX_PIPE_TRY(cudaEventRecord(hp.ev_d2h[c], hp.s_d2h));
        ctx->launches += 1;
    }
    if (want_norm) TDX_PIPE_TRY(cudaMemcpyAsync(hp.sq_host, hp.sq, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, hp.s_d2h));
    TDX_PIPE_TRY(cudaStreamSynchronize(hp.s_d2h));
    TDX_PIPE_TRY(cudaStreamSynchronize(hp.s_comp));
    TDX_PIPE_TRY(cudaStreamSynchronize(hp.s_h2d));
#undef TDX_PIPE_TRY
    if (e != cudaSuccess) return cuda_fail((int)e, "tdx_dek1_attempt_step_host");
    if (want_norm) {
        double acc = 0.0;
        for (int c = 0; c < nc; ++c) acc += hp.sq_host[c];   // chunk order: deterministic
        *h_sq_norm_sum = acc;
    }
    return TDX_OK;
}


// =====================================================================================================================
// Native accept / reject loop (odefilter.py:171-223, step.py:28-108)
// =====================================================================================================================
static int ensure_host_norm(tdx_ctx* ctx) {
    if (ctx->host_norm) return TDX_OK;
    cudaError_t e = cudaHostAlloc((void**)&ctx->host_norm, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&ctx->host_norm_dev, ctx->host_norm, 0);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->zsq_scratch, 64);
    return e == cudaSuccess ? TDX_OK : cuda_fail((int)e, "full-step loop (pinned scalar)");
}

// wait until the kernel epilogue has replaced the sentinel; polls the stream now and then so that a failed launch
// cannot spin forever
static int wait_norm(tdx_ctx* ctx, cudaStream_t st, double* value) {
    volatile double* p = ctx->host_norm;
    unsigned long long spins = 0;
    for (;;) {
        const double v = *p;
        if (!(v == -1.0)) {
            *value = v;
            return TDX_OK;
        }
        if ((++spins & 0xfffffull) == 0ull) {
            cudaError_t q = cudaStreamQuery(st);
            if (q == cudaSuccess) {           // the stream drained: the value must be there now
                *value = *p;
                if (*value == -1.0) return fail(TDX_ERR_CUDA, "full-step loop: the step kernel finished without writing its norm");
                return TDX_OK;
            }
            if (q != cudaErrorNotReady) return cuda_fail((int)q, "full-step loop");
        }
    }
}

template <class Attempt>
static int full_step_loop(tdx_ctx* ctx, const tdx_steprule* rule, double t, double dt, double tmax, int rate,
                          tdx_full_step_result* result, cudaStream_t st, const char* where, Attempt attempt) {
    if (!rule || !result) return fail(TDX_ERR_INVALID, std::string(where) + ": null step rule or result");
    if (!(dt > 0.0) || !std::isfinite(dt)) return fail(TDX_ERR_INVALID, std::string(where) + ": Invalid step size: dt must be positive and finite");
    int rc = ensure_host_norm(ctx);
    if (rc) return rc;
    const double d_global = (double)ctx->geo.nf * (double)ctx->geo.grows * (double)ctx->geo.gcols;
    int64_t attempts = 0;
    for (;;) {
        ++attempts;
        tdx_step_args a{t, dt, rule->adaptive ? rule->abstol : 0.0, rule->adaptive ? rule->reltol : 0.0, 0u, 0u};
        *ctx->host_norm = -1.0;                      // sentinel: the sum of squares is >= 0 or NaN
        if ((rc = attempt(a))) return rc;
        if (!rule->adaptive) {                       // ConstantSteps: always accepted (step.py:41-42), nothing to read
            result->t_new = t + dt;
            result->dt_accepted = dt;
            result->dt_next = std::min(rule->constant_dt, tmax - (t + dt));
            result->error_norm = std::nan("");
            result->attempts = attempts;
            return TDX_OK;
        }
        double sum = 0.0;
        if ((rc = wait_norm(ctx, st, &sum))) return rc;
        const double norm = std::sqrt(sum / d_global);                       // step.py:101-104
        if (std::isnan(norm)) return fail(TDX_ERR_NUMERIC, std::string(where) + ": error estimate is NaN");
        const bool accepted = norm < 1.0;                                     // step.py:90-91
        double factor = rule->safety_scale * std::pow(norm != 0.0 ? 1.0 / norm : INFINITY, 1.0 / (double)rate);   // step.py:80-88
        factor = std::max(rule->min_change, std::min(factor, rule->max_change));
        const double suggested = factor * dt;
        const double remaining = tmax - (accepted ? t + dt : t);              // odefilter.py:214-219
        const double dt_next = std::min(suggested, remaining);
        if (accepted) {
            result->t_new = t + dt;
            result->dt_accepted = dt;
            result->dt_next = dt_next;
            result->error_norm = norm;
            result->attempts = attempts;
            return TDX_OK;
        }
        if (!(dt_next > 0.0)) return fail(TDX_ERR_INVALID, std::string(where) + ": Invalid step size after a rejected attempt");
        dt = dt_next;
    }
}

static int check_full_step_flags(const tdx_ctx* ctx, uint32_t flags, const char* where) {
    const bool sharded = ctx->halo_lo > 0 || ctx->halo_hi > 0;
    const unsigned comm = TDX_STEP_COMM_PACK | TDX_STEP_COMM_REDUCE;
    if (sharded && (flags & comm) != comm)
        return fail(TDX_ERR_INVALID, std::string(where) + ": a sharded context needs TDX_STEP_COMM_PACK | TDX_STEP_COMM_REDUCE so "
                                                          "that every rank sees the global norm");
    if (flags & ~(comm | TDX_STEP_ZERO_JACOBIAN)) return fail(TDX_ERR_INVALID, std::string(where) + ": unsupported flags");
    return TDX_OK;
}

extern "C" int tdx_dek1_full_step(tdx_ctx* ctx, const tdx_steprule* rule, double t, double dt, double tmax, uint32_t flags,
                                  const void* mean_in, const void* sc_in, void* mean_out, void* sc_out, void* err_out,
                                  void* ref_out, tdx_full_step_result* result, void* stream) {
    int rc = check_ready(ctx, "tdx_dek1_full_step");
    if (rc) return rc;
    if ((rc = check_full_step_flags(ctx, flags, "tdx_dek1_full_step"))) return rc;
    DeviceGuard guard(ctx->device);
    const bool sharded = ctx->halo_lo > 0 || ctx->halo_hi > 0;
    return full_step_loop(ctx, rule, t, dt, tmax, ctx->n, result, (cudaStream_t)stream, "tdx_dek1_full_step",
                          [&](tdx_step_args& a) {
                              a.flags = sharded ? flags : (flags & TDX_STEP_ZERO_JACOBIAN);
                              return tdx_dek1_attempt_step(ctx, &a, mean_in, sc_in, mean_out, sc_out, err_out, ref_out, nullptr,
                                                           nullptr, ctx->host_norm_dev, stream);
                          });
}

extern "C" int tdx_ek0_full_step(tdx_ctx* ctx, const tdx_steprule* rule, double t, double dt, double tmax, uint32_t flags,
                                 const void* mean_in, const void* Cl_in, void* mean_out, void* Cl_out, void* err_out,
                                 void* ref_out, tdx_full_step_result* result, void* stream) {
    int rc = check_ready(ctx, "tdx_ek0_full_step");
    if (rc) return rc;
    if ((rc = check_full_step_flags(ctx, flags, "tdx_ek0_full_step"))) return rc;
    DeviceGuard guard(ctx->device);
    if ((rc = ensure_host_norm(ctx))) return rc;
    const bool sharded = ctx->halo_lo > 0 || ctx->halo_hi > 0;
    return full_step_loop(ctx, rule, t, dt, tmax, ctx->n, result, (cudaStream_t)stream, "tdx_ek0_full_step",
                          [&](tdx_step_args& a) {
                              a.flags = sharded ? flags : 0u;
                              return tdx_ek0_attempt_step(ctx, &a, mean_in, Cl_in, mean_out, Cl_out, err_out, ref_out,
                                                          ctx->zsq_scratch, ctx->host_norm_dev, stream);
                          });
}


// =====================================================================================================================
// Device-resident step controller
// =====================================================================================================================
static_assert(TDX_CTL_HISTORY == kCtlHistory, "history ring size");
static_assert(TDX_CTL_MAX_RING == kMaxRing, "state ring size");

extern "C" int tdx_ctl_begin_ex(tdx_ctx* ctx, const tdx_steprule* rule, double t0, double dt0, double tmax,
                                const tdx_ctl_options* opt, void* stream) {
    int rc = check_ready(ctx, "tdx_ctl_begin");
    if (rc) return rc;
    if (!rule) return fail(TDX_ERR_INVALID, "tdx_ctl_begin: null step rule");
    if (!(dt0 > 0.0) || !std::isfinite(dt0)) return fail(TDX_ERR_INVALID, "tdx_ctl_begin: Invalid step size: dt must be positive and finite");
    const int nbuf = opt ? opt->nbuf : 2;
    if (nbuf < 2 || nbuf > kMaxRing) return fail(TDX_ERR_INVALID, "tdx_ctl_begin: the state ring holds 2 .. TDX_CTL_MAX_RING buffers");
    const long long n_stops = (opt && opt->stop_at) ? opt->n_stops : 0;
    if (n_stops < 0) return fail(TDX_ERR_INVALID, "tdx_ctl_begin: negative number of time stops");
    for (long long i = 1; i < n_stops; ++i)
        if (!(opt->stop_at[i] >= opt->stop_at[i - 1])) return fail(TDX_ERR_INVALID, "tdx_ctl_begin: time stops must be sorted");
    DeviceGuard guard(ctx->device);
    cudaError_t e = cudaSuccess;
    if (!ctx->ctl) {
        e = cudaMalloc((void**)&ctx->ctl, sizeof(StepCtl));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->ctl_host, sizeof(StepCtl));
    }
    if (e == cudaSuccess && !ctx->feed) {
        e = cudaHostAlloc((void**)&ctx->feed, sizeof(HostFeed), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&ctx->feed_dev, ctx->feed, 0);
    }
    if (e == cudaSuccess && n_stops > ctx->ctl_stops_cap) {
        cudaFree(ctx->ctl_stops);
        ctx->ctl_stops = nullptr;
        e = cudaMalloc((void**)&ctx->ctl_stops, (size_t)n_stops * sizeof(double));
        ctx->ctl_stops_cap = (e == cudaSuccess) ? n_stops : 0;
    }
    if (e == cudaSuccess && (rc = ensure_host_norm(ctx))) return rc;
    if (e == cudaSuccess && n_stops > 0)
        e = cudaMemcpyAsync(ctx->ctl_stops, opt->stop_at, (size_t)n_stops * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail((int)e, "tdx_ctl_begin");
    StepCtl& c = *ctx->ctl_host;
    std::memset(&c, 0, sizeof(c));
    c.nu = ctx->n - 1;
    c.t = t0; c.tmax = tmax;
    c.adaptive = rule->adaptive ? 1 : 0;
    c.abstol = rule->abstol; c.reltol = rule->reltol; c.safety_scale = rule->safety_scale;
    c.min_change = rule->min_change; c.max_change = rule->max_change; c.constant_dt = rule->constant_dt;
    c.done = (t0 < tmax) ? 0 : 1;
    c.nbuf = nbuf;
    c.wait_timeout_ns = (long long)ctx->timeout_ns;
    c.stops = n_stops > 0 ? ctx->ctl_stops : nullptr;
    c.n_stops = (int)n_stops;
    c.next_stop = 0;
    if (n_stops > 0) {       // the first step is clipped like every later one (odefilter.py:134-135, 355-373)
        if (t0 >= opt->stop_at[0]) c.next_stop = 1;
        const double loc = c.next_stop < c.n_stops ? opt->stop_at[c.next_stop] : INFINITY;
        if (t0 + dt0 > loc) dt0 = loc - t0;
        if (!(dt0 > 0.0)) return fail(TDX_ERR_INVALID, "tdx_ctl_begin: the first time stop leaves no room for a step");
    }
    c.dt = dt0;
    nordsieck_host(c.nu, dt0, c.p, c.pinv);
    std::memset((void*)ctx->feed, 0, sizeof(HostFeed));
    c.feed = (opt && opt->follow) ? ctx->feed_dev : nullptr;
    ctx->ctl_rule = *rule;
    ctx->ctl_nbuf = nbuf;
    e = cudaMemcpyAsync(ctx->ctl, ctx->ctl_host, sizeof(StepCtl), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);     // the staging copy is reused by tdx_ctl_read
    return e == cudaSuccess ? TDX_OK : cuda_fail((int)e, "tdx_ctl_begin");
}

extern "C" int tdx_ctl_begin(tdx_ctx* ctx, const tdx_steprule* rule, double t0, double dt0, double tmax, void* stream) {
    return tdx_ctl_begin_ex(ctx, rule, t0, dt0, tmax, nullptr, stream);
}

static StepHost ctl_step_host(const tdx_ctx* ctx, uint32_t flags) {
    StepHost h{};                                   // p, p_inv and dt come from the controller block
    std::memcpy(h.ql, ctx->ql_packed, sizeof(h.ql));
    const tdx_steprule& r = ctx->ctl_rule;
    h.abstol = r.adaptive ? r.abstol : 0.0;
    h.reltol = r.adaptive ? r.reltol : 0.0;
    h.want_norm = (h.abstol != 0.0 || h.reltol != 0.0) ? 1 : 0;
    h.flags = flags;
    return h;
}

static int check_ring(const tdx_ctx* ctx, void* const* a, void* const* b, int nbuf, const char* where) {
    if (!a || !b || nbuf != ctx->ctl_nbuf)
        return fail(TDX_ERR_INVALID, std::string(where) + ": the state ring must have the size given to tdx_ctl_begin");
    for (int i = 0; i < nbuf; ++i) {
        if (!a[i] || !b[i]) return fail(TDX_ERR_INVALID, std::string(where) + ": null buffer in the state ring");
        for (int j = 0; j < i; ++j)
            if (a[i] == a[j] || b[i] == b[j]) return fail(TDX_ERR_INVALID, std::string(where) + ": the ring's buffers must be distinct");
    }
    return TDX_OK;
}

// one cooperative launch that runs up to ``k`` attempts
static int dek1_launch_attempts(tdx_ctx* ctx, int k, int64_t max_accepted, uint32_t flags, void* const* mean_ring, void* const* sc_ring,
                                int nbuf, void* const* err_ring, void* const* ref_ring, void* stream, const char* where) {
    const bool sharded = ctx->halo_lo > 0 || ctx->halo_hi > 0;
    const uint32_t f = sharded ? flags : (flags & TDX_STEP_ZERO_JACOBIAN);
    const double d_global = (double)ctx->geo.nf * (double)ctx->geo.grows * (double)ctx->geo.gcols;
    int rc = comm_pack_check(ctx, f, where);
    if (rc) return rc;
    if ((rc = comm_reduce_check(ctx, f, true, where))) return rc;
    const void* halo_lo = nullptr;
    const void* halo_hi = nullptr;
    resolve_halos(ctx, f, halo_lo, halo_hi);
    StepHost h = ctl_step_host(ctx, f);
    comm_pack_advance(ctx, f);
    comm_reduce_advance(ctx, f, true);             // under the controller every attempt ends with a reduction over the ranks
    Workspace ws = make_ws(ctx, f);
    ws.max_accepted = max_accepted;
    comm_pack_advance(ctx, f, (unsigned long long)(k - 1));          // attempt i of the launch uses base + i
    comm_reduce_advance(ctx, f, true, (unsigned long long)(k - 1));
    ctx->launches += 1;
    // the norm sum goes to device scratch: a store to mapped host memory would sit in front of the controller's fence
    const int code = ctx->ops->dek1_ctl(ctx->geo, h, mean_ring, sc_ring, nbuf, k, err_ring, ref_ring, halo_lo, halo_hi, ws, ctx->ctl,
                                        d_global, ctx->zsq_scratch + 1, (cudaStream_t)stream);
    return cuda_fail(code, where);
}

static int ek0_launch_attempts(tdx_ctx* ctx, int k, int64_t max_accepted, uint32_t flags, void* const* mean_ring, void* const* Cl_ring,
                               int nbuf, void* const* err_ring, void* const* ref_ring, void* stream, const char* where) {
    const bool sharded = ctx->halo_lo > 0 || ctx->halo_hi > 0;
    const uint32_t f = sharded ? flags : 0u;
    const double d_global = (double)ctx->geo.nf * (double)ctx->geo.grows * (double)ctx->geo.gcols;
    int rc = comm_pack_check(ctx, f, where);
    if (rc) return rc;
    if ((rc = comm_reduce_check(ctx, f, true, where))) return rc;
    const void* halo_lo = nullptr;
    const void* halo_hi = nullptr;
    resolve_halos(ctx, f, halo_lo, halo_hi);
    StepHost h = ctl_step_host(ctx, f);
    comm_pack_advance(ctx, f);
    comm_reduce_advance(ctx, f, true);
    Workspace ws = make_ws(ctx, f);                 // reduce_a: base sequence number, cl_seq: base
    ws.max_accepted = max_accepted;
    comm_reduce_advance(ctx, f, true);
    CommReduce reduce_b = ws.reduce;
    reduce_b.seq = ctx->red_seq;
    comm_pack_advance(ctx, f, (unsigned long long)(k - 1));
    comm_reduce_advance(ctx, f, true, 2ull * (unsigned long long)(k - 1));   // two reductions per attempt
    ctx->mid_seq += 1;
    const unsigned long long mid_base = ctx->mid_seq;
    ctx->mid_seq += (unsigned long long)(k - 1);
    ctx->cl_seq += (unsigned long long)k;
    ctx->launches += 1;
    unsigned long long* mid_flag = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(ctx->mid) + 128);
    const int code = ctx->ops->ek0_ctl(ctx->geo, h, mean_ring, Cl_ring, nbuf, k, err_ring, ref_ring, halo_lo, halo_hi, ws, reduce_b,
                                       mid_flag, mid_base, ctx->ctl, d_global, ctx->zsq_scratch, ctx->zsq_scratch + 1,
                                       (cudaStream_t)stream);
    return cuda_fail(code, where);
}

extern "C" int tdx_dek1_run_attempts(tdx_ctx* ctx, int k, int64_t max_accepted, uint32_t flags, void* const* mean_ring,
                                     void* const* sc_ring, int nbuf, void* const* err_ring, void* const* ref_ring, void* stream) {
    int rc = check_ready(ctx, "tdx_dek1_run_attempts");
    if (rc) return rc;
    if (!ctx->ctl) return fail(TDX_ERR_INVALID, "tdx_dek1_run_attempts: call tdx_ctl_begin first");
    if ((rc = check_full_step_flags(ctx, flags, "tdx_dek1_run_attempts"))) return rc;
    if ((rc = check_ring(ctx, mean_ring, sc_ring, nbuf, "tdx_dek1_run_attempts"))) return rc;
    if (k < 1) return fail(TDX_ERR_INVALID, "tdx_dek1_run_attempts: k must be positive");
    if (max_accepted < 0) return fail(TDX_ERR_INVALID, "tdx_dek1_run_attempts: max_accepted must not be negative");
    DeviceGuard guard(ctx->device);
    return dek1_launch_attempts(ctx, k, max_accepted, flags, mean_ring, sc_ring, nbuf, err_ring, ref_ring, stream, "tdx_dek1_run_attempts");
}

extern "C" int tdx_ek0_run_attempts(tdx_ctx* ctx, int k, int64_t max_accepted, uint32_t flags, void* const* mean_ring,
                                    void* const* Cl_ring, int nbuf, void* const* err_ring, void* const* ref_ring, void* stream) {
    int rc = check_ready(ctx, "tdx_ek0_run_attempts");
    if (rc) return rc;
    if (!ctx->ctl) return fail(TDX_ERR_INVALID, "tdx_ek0_run_attempts: call tdx_ctl_begin first");
    if ((rc = check_full_step_flags(ctx, flags, "tdx_ek0_run_attempts"))) return rc;
    if ((rc = check_ring(ctx, mean_ring, Cl_ring, nbuf, "tdx_ek0_run_attempts"))) return rc;
    if (!err_ring) return fail(TDX_ERR_INVALID, "tdx_ek0_run_attempts: err_ring is required");
    for (int i = 0; i < nbuf; ++i)
        if (!err_ring[i]) return fail(TDX_ERR_INVALID, "tdx_ek0_run_attempts: null entry in err_ring");
    if (k < 1) return fail(TDX_ERR_INVALID, "tdx_ek0_run_attempts: k must be positive");
    if (max_accepted < 0) return fail(TDX_ERR_INVALID, "tdx_ek0_run_attempts: max_accepted must not be negative");
    DeviceGuard guard(ctx->device);
    return ek0_launch_attempts(ctx, k, max_accepted, flags, mean_ring, Cl_ring, nbuf, err_ring, ref_ring, stream, "tdx_ek0_run_attempts");
}

// the same attempts as k separate launches of one attempt each (the first design of the controller; kept as the A/B partner
// of the multi-attempt launch)
extern "C" int tdx_dek1_enqueue_attempts(tdx_ctx* ctx, int k, uint32_t flags, void* mean_a, void* sc_a, void* mean_b, void* sc_b,
                                         void* err_out, void* ref_out, void* stream) {
    int rc = check_ready(ctx, "tdx_dek1_enqueue_attempts");
    if (rc) return rc;
    if (!ctx->ctl) return fail(TDX_ERR_INVALID, "tdx_dek1_enqueue_attempts: call tdx_ctl_begin first");
    if ((rc = check_full_step_flags(ctx, flags, "tdx_dek1_enqueue_attempts"))) return rc;
    if (!mean_a || !sc_a || !mean_b || !sc_b || mean_a == mean_b || sc_a == sc_b || ctx->ctl_nbuf != 2)
        return fail(TDX_ERR_INVALID, "tdx_dek1_enqueue_attempts: two distinct state buffer pairs are required (ring of 2)");
    if (k < 1) return fail(TDX_ERR_INVALID, "tdx_dek1_enqueue_attempts: k must be positive");
    DeviceGuard guard(ctx->device);
    void* mean_ab[2] = {mean_a, mean_b};
    void* sc_ab[2] = {sc_a, sc_b};
    void* err_ab[2] = {err_out, err_out};
    void* ref_ab[2] = {ref_out, ref_out};
    for (int i = 0; i < k; ++i)
        if ((rc = dek1_launch_attempts(ctx, 1, 0, flags, mean_ab, sc_ab, 2, err_ab, ref_ab, stream, "tdx_dek1_enqueue_attempts"))) return rc;
    return TDX_OK;
}

extern "C" int tdx_ek0_enqueue_attempts(tdx_ctx* ctx, int k, uint32_t flags, void* mean_a, void* Cl_a, void* mean_b, void* Cl_b,
                                        void* err_out, void* ref_out, void* stream) {
    int rc = check_ready(ctx, "tdx_ek0_enqueue_attempts");
    if (rc) return rc;
    if (!ctx->ctl) return fail(TDX_ERR_INVALID, "tdx_ek0_enqueue_attempts: call tdx_ctl_begin first");
    if ((rc = check_full_step_flags(ctx, flags, "tdx_ek0_enqueue_attempts"))) return rc;
    if (!mean_a || !Cl_a || !mean_b || !Cl_b || !err_out || mean_a == mean_b || Cl_a == Cl_b || ctx->ctl_nbuf != 2)
        return fail(TDX_ERR_INVALID, "tdx_ek0_enqueue_attempts: two distinct state buffer pairs (ring of 2) and err_out are required");
    if (k < 1) return fail(TDX_ERR_INVALID, "tdx_ek0_enqueue_attempts: k must be positive");
    DeviceGuard guard(ctx->device);
    void* mean_ab[2] = {mean_a, mean_b};
    void* Cl_ab[2] = {Cl_a, Cl_b};
    void* err_ab[2] = {err_out, err_out};
    void* ref_ab[2] = {ref_out, ref_out};
    for (int i = 0; i < k; ++i)
        if ((rc = ek0_launch_attempts(ctx, 1, 0, flags, mean_ab, Cl_ab, 2, err_ab, ref_ab, stream, "tdx_ek0_enqueue_attempts"))) return rc;
    return TDX_OK;
}

extern "C" int tdx_ctl_read(tdx_ctx* ctx, tdx_ctl_state* out, double* hist_t, double* hist_dt, void* stream) {
    if (!ctx || !ctx->ctl || !out) return fail(TDX_ERR_INVALID, "tdx_ctl_read: call tdx_ctl_begin first");
    DeviceGuard guard(ctx->device);
    cudaError_t e = cudaMemcpyAsync(ctx->ctl_host, ctx->ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    unsigned long long comm_err = 0ull;
    if (e == cudaSuccess && ctx->comm_local)
        e = cudaMemcpyAsync(&comm_err, &reinterpret_cast<CommHeader*>(ctx->comm_local)->error, sizeof(comm_err),
                            cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail((int)e, "tdx_ctl_read");
    const StepCtl& c = *ctx->ctl_host;
    out->t = c.t; out->dt = c.dt; out->last_norm = c.last_norm;
    out->attempts = c.attempts; out->accepted = c.accepted;
    out->cur = c.cur; out->done = c.done; out->failed = c.failed; out->reserved = 0;
    if (comm_err != 0ull && out->failed == 0) {      // a peer wait timed out somewhere in the batch
        out->failed = TDX_FAIL_PEER;
        out->done = 1;
    }
    if (hist_t) std::memcpy(hist_t, c.hist_t, sizeof(c.hist_t));
    if (hist_dt) std::memcpy(hist_dt, c.hist_dt, sizeof(c.hist_dt));
    return TDX_OK;
}

// ---- following a running solve from the host (dense output) -------------------------------------------------------------
extern "C" int tdx_ctl_poll(tdx_ctx* ctx, tdx_ctl_progress* out) {
    if (!ctx || !ctx->feed || !out) return fail(TDX_ERR_INVALID, "tdx_ctl_poll: call tdx_ctl_begin_ex with follow = 1 first");
    out->accepted = (int64_t)ctx->feed->accepted;
    out->attempts = (int64_t)ctx->feed->attempts;
    return TDX_OK;
}

extern "C" int tdx_ctl_history(tdx_ctx* ctx, int64_t state_index, double* t, double* dt, int64_t* attempts) {
    if (!ctx || !ctx->feed || state_index < 1) return fail(TDX_ERR_INVALID, "tdx_ctl_history: no feed, or state index < 1");
    const long long slot = state_index % kCtlHistory;
    if (t) *t = ctx->feed->hist_t[slot];
    if (dt) *dt = ctx->feed->hist_dt[slot];
    if (attempts) *attempts = (int64_t)ctx->feed->hist_att[slot];
    return TDX_OK;
}

extern "C" int tdx_ctl_abort(tdx_ctx* ctx) {
    if (!ctx || !ctx->feed) return fail(TDX_ERR_INVALID, "tdx_ctl_abort: call tdx_ctl_begin_ex with follow = 1 first");
    ctx->feed->abort = 1ull;
    __sync_synchronize();
    return TDX_OK;
}
