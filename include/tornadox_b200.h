/*
 * tornadox_b200 -- C ABI of the B200-native tornadox filter step.
 *
 * The reference (pnkraemer/tornadox) is pure Python on JAX and has no FFI of its own.  The seam this
 * library replaces is the abstract-method pair of ``ODEFilter`` (tornadox/odefilter.py:346-352)
 *
 *     initialize(self, f, t0, tmax, y0, df, df_diagonal)          -> ODEFilterState
 *     attempt_step(self, state, dt, f, t0, tmax, y0, df, df_diagonal) -> (ODEFilterState, info)
 *
 * as implemented by ``ek1.DiagonalEK1`` (tornadox/ek1.py:228-369) and ``ek0.KroneckerEK0``
 * (tornadox/ek0.py:152-193), plus the d-wide reduction of ``AdaptiveSteps.scale_error_estimate``
 * (tornadox/step.py:93-104) and the vector fields of ``ivp.py`` (vanderpol, brusselator, lorenz96,
 * fhn_2d, burgers_1d, wave_1d, threebody, pleiades: tornadox/ivp.py:28-505), which are evaluated inside the step kernels
 * (any other vector field: TDX_IVP_EXTERNAL).
 *
 * Conventions
 *  - plain pointers and sizes only; every array pointer is a DEVICE pointer owned by the caller;
 *  - every call enqueues work on ``stream`` (a cudaStream_t passed as void*) and returns without
 *    synchronising; scalars that the host loop needs are written to device scalars the caller reads;
 *  - return value 0 = OK, otherwise a TDX_ERR_* code; ``tdx_last_error()`` has the message;
 *  - in/out state buffers are distinct, so a rejected attempt leaves the input state untouched
 *    (odefilter.py:186-219 re-attempts from the same state);
 *  - layouts are the reference's: mean (n, d) C-order, DiagonalEK1 factors (d, n, n) C-order,
 *    KroneckerEK0 factor (n, n) C-order; n = nu + 1; dtype is f64 (default) or f32.
 *  - multi-GPU: one context per process/GPU owning a shard (rows of the fhn grid, or a range of the
 *    1-d chain) of BOTH fields.  Halos and scalar sums travel through NVLink peer memory written and read by the step
 *    kernels themselves (tdx_comm_*, flags TDX_STEP_COMM_PACK | TDX_STEP_COMM_REDUCE); a caller without peer access can
 *    still exchange the halo buffers and all-reduce the scalars itself (tdx_pack_halo + the three-phase entry points).
 *  - sums over the state dimension (error norm, sum of z^2) are formed from one partial per 256-coordinate tile (or per
 *    block of 8 grid rows x 256 columns in the fused fhn_2d KroneckerEK0 kernel) in a fixed order over 16 global chunks, so their bits do not depend on how many CTAs ran
 *    and -- when the shard boundaries fall on chunk boundaries, e.g. 2 / 4 / 8 equal shards of a 2048 x 2048 grid -- not
 *    on the number of GPUs either: the accepted-step sequence is identical at 1 / 2 / 4 / 8 GPUs.
 */
#ifndef TORNADOX_B200_H_
#define TORNADOX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDX_ABI_VERSION 1

/* error codes */
#define TDX_OK 0
#define TDX_ERR_INVALID 1   /* bad argument (ValueError on the Python side)              */
#define TDX_ERR_UNSUPPORTED 2 /* no kernel instantiated for this (kind, nu, dtype)       */
#define TDX_ERR_CUDA 3      /* a CUDA runtime call failed                                */

/* dtypes */
#define TDX_F64 0
#define TDX_F32 1

/* vector fields evaluated in-kernel (tornadox/ivp.py) */
#define TDX_IVP_VANDERPOL 0   /* params: [mu] or [mu, 1] (vanderpol_julia, ivp.py:52-73)   ivp.py:28-49 */
#define TDX_IVP_LORENZ96 1    /* params: [forcing]                   ivp.py:268-294 */
#define TDX_IVP_BRUSSELATOR 2 /* params: [] (c derived from N)       ivp.py:219-265 */
#define TDX_IVP_FHN2D 3       /* params: [dx, a, b, k, tau]          ivp.py:409-505 */
#define TDX_IVP_BURGERS1D 4   /* params: [dx, diffusion_param]       ivp.py:76-107  (one GPU only) */
#define TDX_IVP_WAVE1D 5      /* params: [dx, diffusion_param]       ivp.py:110-140 (one GPU only) */
#define TDX_IVP_THREEBODY 6   /* params: []   grid_cols = 4          ivp.py:188-216 (one GPU only) */
#define TDX_IVP_PLEIADES 7    /* params: []   grid_cols = 28         ivp.py:344-406 (one GPU only) */
/* any other vector field: the caller evaluates f (and the Jacobian diagonal) itself, see tdx_predict_state / tdx_set_external.
 * grid_rows = 1, grid_cols = d, one GPU. */
#define TDX_IVP_EXTERNAL 8

typedef struct tdx_ctx tdx_ctx;

/* Problem + shard geometry.  State layout is y = [field0(points); field1(points)] (ivp.py:229,438).
 * fhn_2d: points = grid_rows x grid_cols row-major, shard = rows [shard_begin, shard_end).
 * 1-d problems (lorenz96, brusselator): grid_rows = 1, grid_cols = N, shard = points
 * [shard_begin, shard_end).  vanderpol: grid_rows = grid_cols = 1, one point with two fields. */
typedef struct {
    int32_t kind;          /* TDX_IVP_*                                                 */
    int32_t nu;            /* num_derivatives (odefilter.py:34); n = nu + 1             */
    int64_t grid_rows;     /* global                                                     */
    int64_t grid_cols;     /* global                                                     */
    int64_t shard_begin;   /* first owned row (fhn) / point (1-d)                        */
    int64_t shard_end;     /* one past the last owned row / point                        */
    int32_t n_params;
    double params[8];
} tdx_problem_desc;

/* Per-attempt scalars of AdaptiveSteps (step.py:55-104); for ConstantSteps pass abstol=reltol=0
 * and ignore the norm output. */
typedef struct {
    double t;       /* state.t (the kernels evaluate f at t + dt; all in-scope fields are autonomous) */
    double dt;
    double abstol;
    double reltol;
    uint32_t flags; /* TDX_STEP_* */
    uint32_t reserved;
} tdx_step_args;

/* DiagonalEK0 (ek0.py:196-280): run the DiagonalEK1 kernel with the Jacobian diagonal forced to zero */
#define TDX_STEP_ZERO_JACOBIAN 1u
/* Multi-GPU overlap: process only the tiles that do NOT read the halo buffers (INTERIOR; halo pointers may be
 * NULL) or only those that do (RIM).  The two launches together cover the shard exactly once; each writes its own
 * partial norm sum (add them).  Neither flag (or both) = all tiles in one launch. */
#define TDX_STEP_TILES_INTERIOR 2u
#define TDX_STEP_TILES_RIM 4u
/* tdx_ek0_mid: z_sq_sum points at TWO doubles (interior + rim launches of sweep A) that are added */
#define TDX_STEP_TWO_PARTS 8u

const char* tdx_last_error(void);
int tdx_abi_version(void);

/* context: one per device/shard.  ``device`` is the CUDA ordinal. */
int tdx_create(tdx_ctx** out, int device, int dtype);
int tdx_destroy(tdx_ctx* ctx);
int tdx_set_problem(tdx_ctx* ctx, const tdx_problem_desc* desc);
/* local state dimension d of this shard (n_fields * owned points) */
int64_t tdx_local_dim(const tdx_ctx* ctx);
/* number of halo scalars per side: lo = values needed from the lower neighbour, hi = from the upper;
 * the same counts (swapped) are what tdx_pack_halo produces for the neighbours. */
int tdx_halo_sizes(const tdx_ctx* ctx, int64_t* lo, int64_t* hi);

/* Prior constants A (n,n) and Ql (n,n) the kernels use (iwp.py:19-37), written to HOST buffers. */
int tdx_prior(int nu, double* A_out, double* Ql_out);
/* Replace the context's Ql (row-major (n,n), lower triangular, HOST buffer) by the caller's own factor, e.g. the
 * LAPACK dpotrf result the reference computes in iwp.py:37.  The built-in factor is computed in extended
 * precision; dpotrf in fp64 differs from it by cond(Q) * eps (1e-11 relative at nu = 5), which matters when
 * bit-level agreement with a particular host's reference run is wanted.  Call after tdx_set_problem. */
int tdx_set_diffusion_sqrt(tdx_ctx* ctx, const double* Ql);

/* sqrt.propagate_cholesky_factor / sqrt.sqrtm_to_cholesky and their vmapped versions (sqrt.py:8-30) as a standalone
 * batched op on DEVICE buffers: out[b] (n, n) = lower factor L with L L^T = S1[b] S1[b]^T + S2[b] S2[b]^T, i.e. the
 * transposed R-factor of qr([S1^T ; S2^T]) with LAPACK's sign convention (what jnp.linalg.qr returns on CPU).
 * S2 == NULL: sqrtm_to_cholesky of the square right square root St = S1^T.  ``*_broadcast`` != 0: the same (n, n)
 * matrix for every batch element.  ``rows_given`` != 0: S1 / S2 hold the rows of the stacked matrix St itself.
 * No context needed; 1 <= n - 1 <= 5 (the instantiated orders). */
int tdx_propagate_cholesky_factor(int dtype, int n, int64_t batch, const void* S1, const void* S2, void* out,
                                  int s1_broadcast, int s2_broadcast, int rows_given, void* stream);

/* Nordsieck preconditioner vectors (iwp.py:65-73), HOST buffers of n doubles each. */
int tdx_nordsieck(int nu, double dt, double* p_out, double* p_inv_out);

/* f(t, y) and optionally diag(df)(t, y) of the context's vector field on the local shard
 * (replaces the Python callables ivp.f / ivp.df_diagonal; used for first_dt step.py:111-115,
 * initialisation and tests).  halo_lo/halo_hi: neighbour values of y itself, or NULL. */
int tdx_eval_f(tdx_ctx* ctx, double t, const void* y, void* fy, void* jdiag_or_null,
               const void* halo_lo, const void* halo_hi, void* stream);

/* Predicted state E0 P A P^-1 m (the point where f is evaluated, ek1.py:304-306 / ek0.py:146-147)
 * at the shard's first / last rows or points, for the neighbours' halos.
 * edge_first has halo_hi-of-lower-neighbour layout, edge_last has halo_lo-of-upper-neighbour layout. */
int tdx_pack_halo(tdx_ctx* ctx, double dt, const void* mean_in, void* edge_first, void* edge_last,
                  void* stream);

/* ---- vector fields the library does not know (the reference takes any Python callable f / df_diagonal, odefilter.py:346-352)
 * The filter step needs f and diag(df) at ONE point per attempt, the predicted state m_at = E0 P A P^-1 m (ek1.py:304-306,
 * ek0.py:146-147).  For a TDX_IVP_EXTERNAL context every attempt is
 *     tdx_predict_state(ctx, dt, mean, m_at)          m_at (d) on the device
 *     fx = f(t + dt, m_at), jd = df_diagonal(t + dt, m_at)   the caller's own code on device arrays
 *     tdx_set_external(ctx, fx, jd)                   jd may be NULL (DiagonalEK0 / KroneckerEK0)
 *     tdx_dek1_attempt_step / tdx_ek0_attempt_step    the same fused kernels, reading fx / jd instead of evaluating a stencil
 * The arrays must stay valid until the step has run.  tdx_predict_state works for every context (all fields, local shard). */
int tdx_predict_state(tdx_ctx* ctx, double dt, const void* mean_in, void* y_out, void* stream);
int tdx_set_external(tdx_ctx* ctx, const void* fx, const void* jdiag_or_null);

/* DiagonalEK1.attempt_step (ek1.py:228-260 + 274-369), fused with
 * sum_i (dt * err_i / (abstol + reltol * ref_i))^2 over the local shard (step.py:101-104),
 * which is written to the device scalar ``sq_norm_sum`` (double).  err_out / ref_out may be NULL. */
int tdx_dek1_attempt_step(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in,
                          const void* sc_in, void* mean_out, void* sc_out, void* err_out,
                          void* ref_out, const void* halo_lo, const void* halo_hi,
                          double* sq_norm_sum, void* stream);

/* The same attempt for a state that lives in HOST memory (all array pointers are host pointers with the reference's
 * layouts; pinned memory is needed for full PCIe speed, pageable memory works).  The state is streamed through the
 * device in ``nchunks`` pieces (0 = default) with the copies in both directions overlapping the kernels; the call
 * returns when the outputs are complete in host memory.  ``sq_norm_sum`` is a HOST double.  err_out / ref_out may be
 * NULL.  The context must own the whole problem (no sharding). */
int tdx_dek1_attempt_step_host(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in,
                               const void* sc_in, void* mean_out, void* sc_out, void* err_out,
                               void* ref_out, double* sq_norm_sum, int nchunks);

/* The same for ONE SHARD of a multi-GPU run: the host arrays hold the shard's state ((n, d_local), (d_local, n, n)), and
 * halo_lo / halo_hi are DEVICE buffers with the predicted edge state received from the neighbouring ranks (tdx_pack_halo of
 * the shard's edge rows + the caller's exchange); sq_norm_sum receives the shard's partial sum (all-reduce it). */
int tdx_dek1_attempt_step_host_sharded(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in,
                                       const void* sc_in, void* mean_out, void* sc_out, void* err_out,
                                       void* ref_out, const void* halo_lo, const void* halo_hi,
                                       double* sq_norm_sum, int nchunks);

/* KroneckerEK0.attempt_step (ek0.py:152-193) in three phases, so that a multi-GPU caller can
 * all-reduce the device scalar between them:
 *  sweep_a : predict + measure, writes sum_i z_i^2 over the local shard to ``z_sq_sum`` (double);
 *  mid     : calibration, the single 2n x n QR, gain (ek0.py:123-140, 175); needs the GLOBAL
 *            z_sq_sum and the global dimension; writes Cl_out (n,n), err_out (scalar, dtype) and the
 *            gain into the context;
 *  sweep_b : m_new = P (mp - K z), reference_state = |m_new[0]| (ref_out may be NULL) and
 *            sum_i (dt * err / (abstol + reltol * ref_i))^2 over the local shard -> ``sq_norm_sum`` (double).
 * The scaled error norm is then sqrt(sq_norm_sum / d_global) (step.py:101-104), as for DiagonalEK1. */
int tdx_ek0_sweep_a(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in,
                    const void* halo_lo, const void* halo_hi, double* z_sq_sum, void* stream);
int tdx_ek0_mid(tdx_ctx* ctx, const tdx_step_args* args, const void* Cl_in, const double* z_sq_sum,
                int64_t d_global, void* Cl_out, void* err_out, void* stream);
int tdx_ek0_sweep_b(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in, void* mean_out,
                    void* ref_out, const void* halo_lo, const void* halo_hi,
                    double* sq_norm_sum, void* stream);
/* single-GPU convenience: the three phases back to back on ``stream``. */
int tdx_ek0_attempt_step(tdx_ctx* ctx, const tdx_step_args* args, const void* mean_in,
                         const void* Cl_in, void* mean_out, void* Cl_out, void* err_out,
                         void* ref_out, double* z_sq_sum, double* sq_norm_sum, void* stream);

/* ---- native accept / reject loop --------------------------------------------------------------------------------
 * ODEFilter.perform_full_step (odefilter.py:171-223) with the step rules of step.py:28-108: attempt steps from the input
 * state until one is accepted, entirely inside the library.  The kernel epilogue writes the reduced error norm to
 * mapped pinned host memory, the host loop polls it (a few microseconds instead of a stream synchronisation + a Python
 * round trip per attempt), decides accept / reject, proposes the next dt exactly as AdaptiveSteps.suggest does
 * (step.py:80-91) and relaunches.  Rejected attempts overwrite the same output buffers; the input state is untouched.
 * Sharded contexts need the peer-memory exchange (flags TDX_STEP_COMM_PACK | TDX_STEP_COMM_REDUCE): every rank then sees
 * the same global norm and takes the same decisions.  Blocks until a step is accepted (ConstantSteps: returns after the
 * launch, nothing to wait for). */
typedef struct {
    int32_t adaptive;      /* 1: AdaptiveSteps (step.py:55-108); 0: ConstantSteps (step.py:28-52)          */
    int32_t reserved;
    double abstol, reltol; /* AdaptiveSteps                                                               */
    double min_change, max_change; /* AdaptiveSteps.max_changes                                           */
    double safety_scale;
    double constant_dt;    /* ConstantSteps.dt                                                            */
} tdx_steprule;

typedef struct {
    double t_new;          /* time of the accepted state                                                  */
    double dt_accepted;    /* the step that was accepted                                                  */
    double dt_next;        /* next step: suggestion clipped to tmax - t_new (odefilter.py:214-219)         */
    double error_norm;     /* scaled error norm of the accepted attempt (NaN for ConstantSteps)           */
    int64_t attempts;      /* num_attempted_steps of this call                                            */
} tdx_full_step_result;

#define TDX_ERR_NUMERIC 4  /* the error estimate became NaN (the reference would loop forever, odefilter.py:179-219) */

int tdx_dek1_full_step(tdx_ctx* ctx, const tdx_steprule* rule, double t, double dt, double tmax, uint32_t flags,
                       const void* mean_in, const void* sc_in, void* mean_out, void* sc_out, void* err_out,
                       void* ref_out, tdx_full_step_result* result, void* stream);
int tdx_ek0_full_step(tdx_ctx* ctx, const tdx_steprule* rule, double t, double dt, double tmax, uint32_t flags,
                      const void* mean_in, const void* Cl_in, void* mean_out, void* Cl_out, void* err_out,
                      void* ref_out, tdx_full_step_result* result, void* stream);

/* ---- device-resident step controller ----------------------------------------------------------------------------
 * The same loop with NO host involvement between attempts: the step to attempt, its Nordsieck vectors, the time and
 * the index of the current buffer pair live in a controller block in device memory.  Every launch of a batch reads them,
 * its last CTA decides accept / reject from the (globally reduced) norm, proposes the next dt, flips the buffers on
 * accept and stops the solve at tmax; launches enqueued after that return immediately.  The host enqueues k attempts
 * blindly and looks at the controller once per batch.  State buffers: pair A holds the initial state, pair B is scratch;
 * after tdx_ctl_read the current state is pair ``cur`` (0 = A, 1 = B).
 *   tdx_ctl_begin             reset the controller: rule, t0, first dt, tmax (buffers: cur = A)
 *   tdx_dek1_enqueue_attempts k DiagonalEK1 attempts (all vector fields)
 *   tdx_ek0_enqueue_attempts  k KroneckerEK0 attempts (fhn_2d: the fused kernel; other fields: TDX_ERR_UNSUPPORTED)
 *   tdx_ctl_read              synchronise ``stream`` and copy the controller's scalars (and, optionally, the ring of the
 *                             last TDX_CTL_HISTORY accepted (t, dt) pairs, indexed by accepted-count % TDX_CTL_HISTORY) */
#define TDX_CTL_HISTORY 256
typedef struct {
    double t, dt, last_norm;
    int64_t attempts, accepted;
    int32_t cur, done, failed, reserved;
} tdx_ctl_state;
int tdx_ctl_begin(tdx_ctx* ctx, const tdx_steprule* rule, double t0, double dt0, double tmax, void* stream);
int tdx_dek1_enqueue_attempts(tdx_ctx* ctx, int k, uint32_t flags, void* mean_a, void* sc_a, void* mean_b, void* sc_b,
                              void* err_out, void* ref_out, void* stream);
int tdx_ek0_enqueue_attempts(tdx_ctx* ctx, int k, uint32_t flags, void* mean_a, void* Cl_a, void* mean_b, void* Cl_b,
                             void* err_out, void* ref_out, void* stream);
int tdx_ctl_read(tdx_ctx* ctx, tdx_ctl_state* out, double* hist_t, double* hist_dt, void* stream);

/* ---- many attempts in ONE launch ----------------------------------------------------------------------------------
 * The reference compiles the whole accept / reject loop into one ``lax.while_loop`` (odefilter.py:225-344).  Here
 * tdx_*_run_attempts is ONE persistent cooperative kernel that runs up to ``k`` attempts: after every attempt the last CTA
 * reduces the norm (over the ranks too), updates the controller and releases an in-kernel grid barrier; halo delivery and both
 * all-reduces of a sharded run happen inside the same launch.  The launch ends early when the solve reaches tmax, or after
 * ``max_accepted`` newly accepted steps (0: no limit).
 * State buffers form a RING of ``nbuf`` (mean, factor) pairs (2 <= nbuf <= TDX_CTL_MAX_RING; the same ring must be passed to
 * every launch of a solve): attempt output goes to slot (cur + 1) % nbuf, an accepted attempt advances ``cur``, so accepted
 * state number j (0 = the initial state, in slot 0) lives in slot j % nbuf until nbuf - 1 further steps have been accepted.
 * Dense output without a host round trip per step: launch with max_accepted = nbuf - 1 on one stream and copy the states out
 * of the ring on another stream while the kernel goes on with the next attempts; within one launch no slot is written twice.
 *   tdx_ctl_begin_ex   tdx_ctl_begin with options: ring size, time stops (odefilter.py:355-373: the solve lands exactly on
 *                      every stop), and ``follow``: the kernel publishes its progress to mapped pinned memory
 *   tdx_ctl_poll       (follow) accepted / attempted steps so far, without synchronising anything
 *   tdx_ctl_history    (follow) time, step and attempt count of accepted state ``state_index`` >= 1 (kept for the last
 *                      TDX_CTL_HISTORY states)
 *   tdx_ctl_abort      (follow) stop the running launch at the next attempt boundary
 * A wait inside the kernels (peer halos, reductions, grid barriers) that exceeds TDX_WAIT_TIMEOUT_MS (environment, default
 * 4000) poisons the attempt's norm with NaN and stops the solve on every rank with ``failed`` = 2 (poisoned attempt), 3 (the
 * grid barrier between two attempts), 4 (reserved) or 5 (a peer); 1 = NaN error estimate. */
#define TDX_CTL_MAX_RING 8
typedef struct {
    int32_t nbuf;          /* ring size                                                                            */
    int32_t follow;        /* 1: dense output, the host follows the solve                                          */
    const double* stop_at; /* HOST array of time stops, sorted ascending, or NULL                                   */
    int64_t n_stops;
} tdx_ctl_options;
typedef struct {
    int64_t accepted, attempts;
} tdx_ctl_progress;
int tdx_ctl_begin_ex(tdx_ctx* ctx, const tdx_steprule* rule, double t0, double dt0, double tmax,
                     const tdx_ctl_options* options, void* stream);
/* err_ring / ref_ring: per-slot error_estimate (DiagonalEK1: (d) vectors or NULL; KroneckerEK0: scalars, required) and
 * reference_state vectors (NULL, or NULL entries: not written), so that they belong to the state in the same slot */
int tdx_dek1_run_attempts(tdx_ctx* ctx, int k, int64_t max_accepted, uint32_t flags, void* const* mean_ring,
                          void* const* sc_ring, int nbuf, void* const* err_ring, void* const* ref_ring, void* stream);
int tdx_ek0_run_attempts(tdx_ctx* ctx, int k, int64_t max_accepted, uint32_t flags, void* const* mean_ring,
                         void* const* Cl_ring, int nbuf, void* const* err_ring, void* const* ref_ring, void* stream);
int tdx_ctl_poll(tdx_ctx* ctx, tdx_ctl_progress* out);
int tdx_ctl_history(tdx_ctx* ctx, int64_t state_index, double* t, double* dt, int64_t* attempts);
int tdx_ctl_abort(tdx_ctx* ctx);

/* ---- peer-memory exchange over NVLink (multi-GPU, one process per GPU on one node) -------------------------
 * No reference counterpart (the reference is single-device).  Instead of NCCL calls between the kernels, every
 * context owns a small device block (halo buffers, reduction slots, sequence-numbered flags) that its neighbours
 * map through CUDA IPC; the pack kernel stores the halos directly into the neighbours' blocks, the step kernels
 * spin (with a timeout) on the flag before they read them, and a one-CTA kernel all-reduces the scalars in rank
 * order (deterministic).  Nothing here needs a free SM while a persistent step kernel is running.
 *   tdx_comm_init     allocate the block; ``handle_out`` receives TDX_COMM_HANDLE_BYTES to all-gather
 *   tdx_comm_connect  map rank ``peer``'s block from its handle (call for every peer, then barrier)
 *   tdx_comm_pack_halo   predicted edge state -> neighbours' halo buffers (+ flags); replaces tdx_pack_halo + send/recv
 *   step calls with TDX_STEP_COMM_HALO read those buffers (halo pointers are ignored); in a single launch the tiles
 *                     that need them are processed last and only then is the arrival flag awaited
 *   step calls with TDX_STEP_COMM_REDUCE return the GLOBAL sum (the last CTA exchanges the partial sums)
 *   tdx_comm_allreduce   vals[0..count) summed over ranks in place (count <= 4), device scalars
 *   tdx_comm_error    non-zero after a flag wait timed out (a peer died or diverged) */
#define TDX_COMM_HANDLE_BYTES 64
#define TDX_STEP_COMM_HALO 16u
/* the launch itself first delivers this shard's edge state to the neighbours (a fused tdx_comm_pack_halo: a few of
 * its CTAs do it before they start on their tiles), then behaves like TDX_STEP_COMM_HALO */
#define TDX_STEP_COMM_PACK 64u
/* all-reduce the call's scalar result over the ranks inside the kernel epilogue (needs tdx_comm_connect to all ranks) */
#define TDX_STEP_COMM_REDUCE 32u
int tdx_comm_init(tdx_ctx* ctx, int rank, int world, int lower_rank /* -1: none */, int upper_rank /* -1: none */,
                  void* handle_out);
int tdx_comm_connect(tdx_ctx* ctx, int peer, const void* handle);
int tdx_comm_pack_halo(tdx_ctx* ctx, double dt, const void* mean_in, void* stream);
int tdx_comm_allreduce(tdx_ctx* ctx, double* vals, int count, void* stream);
int tdx_comm_error(tdx_ctx* ctx, int* error_out, void* stream);

/* kernels launched by this context since creation (bench.py's gpu_launches counter) */
int64_t tdx_launch_count(const tdx_ctx* ctx);
/* development aid: copy ``count`` doubles of the context's reduction workspace to the host (synchronous) */
int tdx_debug_read_workspace(tdx_ctx* ctx, double* host_out, int64_t offset, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* TORNADOX_B200_H_ */
