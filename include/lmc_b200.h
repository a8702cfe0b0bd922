/*
 * lmc_b200.h -- C ABI of the B200-native HMC / NUTS hot path (drop-in for eigenfoo/littlemcmc's sampler core).
 *
 * The reference (pure Python, /root/reference/littlemcmc) has no FFI; each entry point below replaces the
 * Python call named in its comment (file:line in the reference tree).  A binding is a ctypes stub: see
 * INTEGRATION.md.  Conventions:
 *   - every pointer is a DEVICE pointer unless the comment says "host"; the library allocates nothing
 *     persistent: all buffers (including scratch) are owned by the caller (PyTorch tensors in our host layer);
 *   - per-chain vectors are rows of row-major [n_chains, ld] float64 arrays, `ld` even and >= ndim, row base
 *     16-byte aligned (so rows can be moved with 128-bit accesses); elements [ndim, ld) are padding;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and
 *     returns without synchronising;
 *   - return value: LMC_OK or a negative LMC_ERR_*; no C++ exception crosses the boundary.  Numerical
 *     failures are per-chain data, not errors: `diverging` statistics and the `status` bit mask
 *     (the analogue of DivergenceInfo / ValueError("Bad initial energy"), base_hmc.py:145-148,164-179);
 *   - re-entrant, no global state: RNG state is explicit (a per-chain 64-bit Philox key + counters derived
 *     from the chain's transition index), unlike the reference's process-global numpy.random.seed stream.
 */
#ifndef LMC_B200_H
#define LMC_B200_H

#ifdef __CUDACC_RTC__ /* run-time compilation of a user target (lmc_user_kernel_build): no host headers */
typedef signed int int32_t;
typedef unsigned int uint32_t;
typedef long int64_t;
typedef unsigned long uint64_t;
typedef unsigned long uintptr_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LMC_ABI_VERSION 3

#define LMC_OK 0
#define LMC_ERR_BADARG (-1)      /* null pointer, bad size/stride/alignment */
#define LMC_ERR_UNSUPPORTED (-2) /* ndim / depth outside what the kernels are instantiated for */
#define LMC_ERR_LAUNCH (-3)      /* cudaGetLastError() != cudaSuccess after the launch */
#define LMC_ERR_WORKSPACE (-4)   /* workspace_bytes too small */

/* ---- built-in target densities (the "user callback" evaluated inside the kernel) ---------------------- */
#define LMC_TARGET_DIAG_GAUSSIAN 0 /* logp = -1/2 sum tau_i q_i^2 ; g = -(tau*q) ; logp = 0.5 * q.g          */
#define LMC_TARGET_FUNNEL 1        /* Neal's funnel: q0=v~N(0,s^2), q_i|v~N(0,e^v)     */

typedef struct lmc_target {
  int32_t kind;      /* LMC_TARGET_*                                                                      */
  int32_t reserved;
  const double* tau; /* DIAG_GAUSSIAN: [ld] precisions 1/sigma^2 shared by all chains (padding = 0)       */
  double v_scale;    /* FUNNEL: s (3.0 in cfg4)                                                           */
} lmc_target;

/* ---- randomness -------------------------------------------------------------------------------------- */
#define LMC_RNG_TAPE 0   /* read pre-drawn numbers (parity tests: same numbers are fed to the CPU oracle)  */
#define LMC_RNG_PHILOX 1 /* counter-based Philox4x32-10 inside the kernel, keyed per chain                 */

typedef struct lmc_rng {
  int32_t mode; /* LMC_RNG_*                                                                              */
  int32_t reserved;
  /* TAPE: standard normals for the momentum draw (quadpotential.py:221-224), one row per transition of
   * this call: normals[(chain * n_trans + t) * ndim + i]                                                 */
  const double* normals;
  /* TAPE: uniforms in [0,1) consumed by sequential counter, reset every transition, in the reference's
   * order (math.py:25 via nuts.py:213,404,321; hmc.py:141,166): uniforms[(chain * n_trans + t) * u_stride + k] */
  const double* uniforms;
  int64_t u_stride;
  const uint64_t* seeds; /* PHILOX: [n_chains] keys (the per-chain seeds of sampling.py:131-134)           */
} lmc_rng;

/* ---- per-chain adaptation scalars: adapt[chain * LMC_ADAPT_STRIDE + k] ------------------------------- */
#define LMC_ADAPT_LOG_STEP 0  /* DualAverageAdaptation._log_step   (step_sizes.py:51)                      */
#define LMC_ADAPT_LOG_BAR 1   /* ._log_bar                                                                 */
#define LMC_ADAPT_HBAR 2      /* ._hbar                                                                    */
#define LMC_ADAPT_COUNT 3     /* ._count                                                                   */
#define LMC_ADAPT_MU 4        /* ._mu = log(10 * initial_step)                                             */
#define LMC_ADAPT_W_FG 5      /* QuadPotentialDiagAdapt._foreground_var.w_sum (quadpotential.py:305)       */
#define LMC_ADAPT_W_BG 6      /* ._background_var.w_sum                                                    */
#define LMC_ADAPT_NSAMPLES 7  /* ._n_samples                                                               */
#define LMC_ADAPT_WINDOW 8    /* .adaptation_window                                                        */
#define LMC_ADAPT_STRIDE 10

/* ---- per-transition statistics: stats[(chain * n_trans + t) * LMC_NSTATS + k], all float64 ------------ */
#define LMC_NSTATS 13
/* NUTS (nuts.py:87-101, 427-435)                 HMC (hmc.py:36-50, 173-181)                              */
#define LMC_STAT_DEPTH 0           /* depth            | n_steps                                           */
#define LMC_STAT_TREE_SIZE 1       /* tree_size        | path_length                                       */
#define LMC_STAT_ACCEPT 2          /* mean_tree_accept | accept                                            */
#define LMC_STAT_ENERGY 3          /* energy                                                               */
#define LMC_STAT_ENERGY_ERROR 4    /* energy_error                                                         */
#define LMC_STAT_MAX_ENERGY_ERROR 5/* max_energy_error | accepted                                          */
#define LMC_STAT_MODEL_LOGP 6      /* model_logp                                                           */
#define LMC_STAT_DIVERGING 7       /* diverging                                                            */
#define LMC_STAT_TUNE 8            /* tune                                                                 */
#define LMC_STAT_STEP_SIZE 9       /* step_size  (post-update: the NEXT draw's, base_hmc.py:161,188)       */
#define LMC_STAT_STEP_SIZE_BAR 10  /* step_size_bar                                                        */
#define LMC_STAT_N_UNIFORMS 11     /* uniforms consumed by this transition (bookkeeping, not in reference) */
#define LMC_STAT_REACHED_MAX_TREEDEPTH 12 /* NUTS: 1 when the doubling loop ran out without a divergence or a U-turn
                                      (the `else` of nuts.py:212-220, what NUTS._reached_max_treedepth counts); HMC: 0 */

/* ---- status bits: status[chain] ------------------------------------------------------------------------ */
#define LMC_STATUS_BAD_INITIAL_ENERGY 1 /* non-finite start energy: the reference raises ValueError
                                           (base_hmc.py:145-148); the chain stops and its remaining
                                           trace/stats rows are NaN                                         */
#define LMC_STATUS_TAPE_EXHAUSTED 2     /* TAPE mode ran past u_stride                                      */

/*
 * Arguments of one sampling call: `n_trans` consecutive transitions of every chain, i.e. the loop body of
 * sampling._iter_sample (sampling.py:507-513) around BaseHMC._astep (base_hmc.py:140-190), for all chains.
 */
typedef struct lmc_sampler_args {
  int32_t abi_version; /* LMC_ABI_VERSION                                                                  */
  int32_t n_chains;
  int32_t ndim;        /* model_ndim                                                                        */
  int32_t reserved0;
  int64_t ld;          /* row stride (elements) of every [n_chains, ld] array below                         */
  lmc_target target;

  /* chain state, updated in place */
  double* q;           /* [n_chains, ld] current position (`q0` of _astep, replaced by `hmc_step.end.q`)    */
  double* var;         /* [n_chains, ld] QuadPotentialDiag(Adapt)._var / .v: diagonal of the inverse mass   */

  /* adaptation (potential.update: quadpotential.py:231-245; step_adapt.update: step_sizes.py:71-92) */
  int32_t adapt_mass;  /* 1: QuadPotentialDiagAdapt, 0: static QuadPotentialDiag                            */
  int32_t adapt_step_size; /* BaseHMC.adapt_step_size                                                       */
  double* mean_fg;     /* [n_chains, ld] _foreground_var.mean       (NULL allowed if !adapt_mass)           */
  double* rawvar_fg;   /* [n_chains, ld] _foreground_var.raw_var                                            */
  double* mean_bg;     /* [n_chains, ld] _background_var.mean                                               */
  double* rawvar_bg;   /* [n_chains, ld] _background_var.raw_var                                            */
  double* adapt;       /* [n_chains, LMC_ADAPT_STRIDE]                                                      */
  double window_multiplier; /* adaptation_window_multiplier                                                 */
  double target_accept, gamma, k, t0; /* DualAverageAdaptation parameters                                   */

  /* schedule */
  int64_t iter0;       /* iter_count of the first transition of this call (0 at the start of a run)         */
  int64_t n_tune;      /* transitions with iter_count < n_tune are tuning (sampling.py:503,510-511)          */
  int32_t n_trans;     /* transitions to run in this call                                                   */
  int32_t reserved1;

  /* step-method parameters (nuts.py:103-121, hmc.py:52-69) */
  double Emax;
  int32_t max_treedepth;       /* NUTS */
  int32_t early_max_treedepth; /* NUTS: cap while tune && iter_count < 200 (nuts.py:205-208)                */
  double path_length;          /* HMC  */
  int32_t max_steps;           /* HMC  */
  int32_t reserved2;

  lmc_rng rng;

  /* outputs */
  double* trace;       /* trace[chain * trace_chain_stride + t * trace_draw_stride + i], i < ndim           */
  int64_t trace_chain_stride;
  int64_t trace_draw_stride;
  double* stats;       /* [n_chains, n_trans, LMC_NSTATS]                                                   */
  int32_t* status;     /* [n_chains] OR-ed LMC_STATUS_* bits                                                */

  /* BaseHMC.step_rand (base_hmc.py:154-155): the reference passes the current step size through a user hook before
   * every transition.  NULL: no hook.  Otherwise [n_chains] step sizes that REPLACE step_adapt.current() for every
   * transition of this call (the host layer evaluates the hook and launches one transition at a time); the
   * dual-averaging update still runs on its own state, as in the reference. */
  const double* step_size_override;

  /* scratch and launch control */
  void* workspace;     /* >= lmc_workspace_bytes(...) bytes, 16-byte aligned                                */
  int64_t workspace_bytes;
  void* stream;        /* cudaStream_t                                                                      */
  int32_t tune_group;  /* 0 = library picks the kernel and threads-per-chain; 1: force the chunked warp-per-chain
                          NUTS kernel (ndim <= 256); 2: force the chunked CTA-per-chain NUTS kernel (128 threads per
                          chain, ndim <= 1024; 203 / 204: with 3 / 4 resident CTAs per SM); >= 32 otherwise: force
                          32/64/128/256/512/1024 threads per chain with the register-resident kernel; < 0: force
                          -tune_group (64/128/256) with the lean NUTS kernel (experiments and tests)          */
  int32_t tune_smem_vecs; /* -1 = library picks how many scratch vectors live in shared memory; else force  */
  int32_t tune_max_slots; /* 0 = library picks the number of resident chain slots; else cap it              */
  int32_t tune_chunk;  /* chunked kernels: leaves per chunk (warp: 2, 4, 8, 16; CTA: 2, 4); 0 = library picks    */

  /* One launch for a whole run with a host-resident trace (ABI v3; fused and user-target kernels only, the callback and
   * dense state machines require 0 / NULL).  The reference discards tuning draws after the fact (sampling.py:473-476)
   * and fills its trace draw by draw (:513); here
   *   trace_skip      the first `trace_skip` transitions of the call are not kept: transition t writes trace row
   *                   max(t - trace_skip, 0) (a discarded draw lands on row 0 and is overwritten by the first kept one),
   *                   so `trace` holds n_trans - trace_skip rows per chain; statistics are written for every transition;
   *   progress        device counters, one per block of `progress_block` KEPT draws (zeroed by the caller): a chain
   *                   adds 1 to counter b after its kept draw (b + 1) * progress_block - 1 (or its last one) is
   *                   globally visible.  counter b == n_chains  <=>  rows [b * progress_block, (b + 1) * progress_block)
   *                   of every chain are final: the caller's copy engine can ship them while the launch keeps sampling. */
  int32_t trace_skip;
  int32_t progress_block;
  int32_t* progress;
} lmc_sampler_args;

/* Library / ABI version (LMC_ABI_VERSION of the build). */
int lmc_abi_version(void);

/* Bytes of scratch lmc_nuts_sample / lmc_hmc_sample need for this problem (host call, no GPU work).
 * `kind`: 0 = NUTS, 1 = HMC; `tune_group` as in lmc_sampler_args.  `max_treedepth` here and in the other *_bytes
 * functions is the SCRATCH depth: max(args.max_treedepth, args.early_max_treedepth) -- trees grow to the early cap
 * during the first 200 tuning transitions even when it exceeds max_treedepth (nuts.py:205-208).  Returns < 0 on
 * unsupported sizes. */
int64_t lmc_workspace_bytes(int32_t kind, int32_t n_chains, int32_t ndim, int32_t max_treedepth,
                            int32_t tune_group);

/* n_trans NUTS transitions for all chains: replaces NUTS._astep -> NUTS._hamiltonian_step -> _Tree.extend /
 * _build_subtree / _single_step (nuts.py:204-224, 284-417) with CpuLeapfrogIntegrator.step
 * (integration.py:100-121), QuadPotentialDiag(Adapt).velocity/energy/random/update
 * (quadpotential.py:206-245, 367-387) and DualAverageAdaptation.current/update (step_sizes.py:58-92). */
int lmc_nuts_sample(const lmc_sampler_args* args);

/* Same for HamiltonianMC._astep -> _hamiltonian_step (hmc.py:140-182). */
int lmc_hmc_sample(const lmc_sampler_args* args);

/*
 * User-written target densities inside the fused sampler kernels.  The reference calls an arbitrary Python
 * `logp_dlogp_func` once per leapfrog (integration.py:62,115; base_hmc.py:34); here a density written as CUDA C++ -- a
 * type with the `pre / grad / finish` protocol of the built-in targets (csrc/lmc_device.cuh, "built-in target
 * densities") -- is compiled INTO the sampler kernel at run time (NVRTC, sm_100a) and runs at the speed of the built-in
 * ones.  `source` defines `struct <type_name>`; the struct is passed to the kernel by value as `target_bytes`
 * (conventionally its only member is `const double* params`, a device pointer to the density's parameters).
 *   kind / ndim / chunk / tape: the launch the kernel is specialised for (KIND 0 = NUTS, 1 = HMC; chunk as
 *   lmc_sampler_args.tune_chunk; tape != 0: randomness from tapes, else in-kernel Philox -- only the chunked kernel
 *   (NUTS, ndim <= 256) is specialised on it);  include_dirs: where lmc_sampler*.cuh and lmc_b200.h live;
 *   cache_path (nullable): the compiled cubin is stored there (+ ".name") and reused by later builds.
 * lmc_user_kernel_log(): compiler output of this thread's last build.  lmc_user_sample == lmc_nuts_sample /
 * lmc_hmc_sample with args->target ignored.
 */
typedef struct lmc_user_kernel lmc_user_kernel;
int lmc_user_kernel_build(const char* source, const char* type_name, int32_t kind, int32_t ndim, int32_t chunk,
                          int32_t tape, const char* const* include_dirs, int32_t n_include_dirs, const char* cache_path,
                          lmc_user_kernel** out);
const char* lmc_user_kernel_log(void);
int lmc_user_sample(lmc_user_kernel* kernel, const lmc_sampler_args* args, const void* target_bytes);
int lmc_user_kernel_destroy(lmc_user_kernel* kernel);

/* CpuLeapfrogIntegrator.compute_state (integration.py:52-66) for all chains with a built-in target:
 * g = dlogp(q), v = var*p, energy = 0.5 p.v - logp.  var_stride = ld for per-chain var, 0 to broadcast one row. */
int lmc_compute_state(const lmc_target* target, int32_t n_chains, int32_t ndim, int64_t ld, const double* q,
                      const double* p, const double* var, int64_t var_stride, double* v, double* g,
                      double* energy, double* logp, void* stream);

/* CpuLeapfrogIntegrator.step (integration.py:100-121) for all chains with a built-in target; eps[chain] may be
 * negative.  In-place (q_out == q etc.) is allowed. */
int lmc_leapfrog_step(const lmc_target* target, int32_t n_chains, int32_t ndim, int64_t ld, const double* eps,
                      const double* q, const double* p, const double* g, const double* var, int64_t var_stride,
                      double* q_out, double* p_out, double* v_out, double* g_out, double* energy, double* logp,
                      void* stream);

/* The two halves of integration.py:100-121 around an external (torch) gradient evaluation:
 *   half1: p <- p + eps/2 * g ; q <- q + eps * (var*p)                          (lines 105-112)
 *   half2: p <- p + eps/2 * g_new ; v = var*p ; energy = 0.5 p.v - logp[chain]  (lines 116-119)
 * `active` (nullable) masks chains: rows with active[chain]==0 are left untouched. */
int lmc_leapfrog_half1(int32_t n_chains, int32_t ndim, int64_t ld, const double* eps, const int32_t* active,
                       double* q, double* p, const double* g, const double* var, int64_t var_stride, void* stream);
int lmc_leapfrog_half2(int32_t n_chains, int32_t ndim, int64_t ld, const double* eps, const int32_t* active,
                       double* p, double* v, const double* g_new, const double* logp, const double* var,
                       int64_t var_stride, double* energy, void* stream);

/*
 * Callback mode: the same transitions around an EXTERNAL gradient -- the user's logp_dlogp_func (base_hmc.py:34,
 * called at integration.py:62 and :115) evaluated by the caller as one batched op over all chains between two
 * launches.  Every chain is a resumable state machine kept in `machine`:
 *     lmc_callback_begin(kind, &c);                         // q_eval <- base.q, every chain asks for its first gradient
 *     while (*n_running > 0) {                              // host reads the device counter now and then
 *         logp_eval[c], g_eval[c, :] = f(q_eval[c, :]);     // caller's op, all chains, same stream
 *         lmc_callback_advance(kind, &c);                   // each chain consumes its gradient and runs to the next
 *     }                                                     // point where it needs one (or finishes)
 * Chains are not in lock step (one may be deep in a tree while another starts its next draw); finished chains idle.
 * `base.target` and `base.workspace` are ignored; everything else in `base` means what it means for lmc_nuts_sample
 * (state arrays updated in place, trace / stats / status written per transition, TAPE or PHILOX randomness).
 */
typedef struct lmc_callback_args {
  lmc_sampler_args base;
  double* q_eval;          /* [n_chains, ld] out: where the callback must be evaluated next (padding = 0)        */
  const double* g_eval;    /* [n_chains, ld] in : dlogp at q_eval (padding ignored)                              */
  const double* logp_eval; /* [n_chains]     in : logp at q_eval                                                 */
  void* machine;           /* >= lmc_callback_state_bytes(...) bytes, 16-byte aligned, owned by the caller       */
  int64_t machine_bytes;
  int32_t* n_running;      /* device counter: chains that still need gradient evaluations                        */
} lmc_callback_args;

/* Bytes of per-chain machine state for callback mode (host call).  `kind`: 0 = NUTS, 1 = HMC. */
int64_t lmc_callback_state_bytes(int32_t kind, int32_t n_chains, int32_t ndim, int32_t max_treedepth);
/* Start `base.n_trans` transitions of every chain: replaces the entry of BaseHMC._astep (base_hmc.py:140-143). */
int lmc_callback_begin(int32_t kind, const lmc_callback_args* args);
/* Consume (logp_eval, g_eval) and advance every chain to its next evaluation point: replaces the code between two
 * logp_dlogp_func calls of the reference (integration.py:116-119 -> nuts.py:352-417 / 315-340 -> base_hmc.py:161-190 ->
 * integration.py:105-112). */
int lmc_callback_advance(int32_t kind, const lmc_callback_args* args);

/*
 * Device-driven callback loop.  The host loop above costs a launch group and, now and then, a host synchronisation per
 * gradient evaluation; here the whole `while (*n_running > 0)` runs on the GPU as ONE graph launch: a CUDA-graph WHILE
 * conditional node whose body is the caller's captured iteration(s) -- `body_graph`, a cudaGraph_t holding one or more
 * repetitions of [the caller's logp/grad op on q_eval -> lmc_callback_advance], captured by the caller on its own stream
 * (torch.cuda.graph) -- followed by a one-thread kernel that reads *n_running and sets the loop condition.  No host
 * involvement between the first and the last gradient evaluation.
 *   lmc_callback_begin(kind, &c);                                    // as above (outside the graph)
 *   lmc_callback_loop_create(body_graph, c.n_running, iters, max_iters, &loop);
 *   lmc_callback_loop_launch(loop, stream);                           // asynchronous; the run is over when it completes
 *   lmc_callback_loop_destroy(loop);                                  // after the stream has drained
 * `iters` (device int32, zeroed by create's caller) counts executed bodies; the loop also stops after `max_iters`
 * bodies (a safety net: *n_running is then still > 0 and the caller reports it).  The body graph is cloned: the caller
 * may destroy its own copy after create.
 */
typedef struct lmc_callback_loop lmc_callback_loop;
int lmc_callback_loop_create(void* body_graph, const int32_t* n_running, int32_t* iters, int64_t max_iters,
                             lmc_callback_loop** loop_out);
int lmc_callback_loop_launch(lmc_callback_loop* loop, void* stream);
int lmc_callback_loop_destroy(lmc_callback_loop* loop);

/*
 * Dense-mass mode: transitions with a DENSE mass matrix -- QuadPotentialFull / QuadPotentialFullInv /
 * QuadPotentialFullAdapt (quadpotential.py:390-615).  With a dense matrix three things are batched operations over
 * chains that cannot live inside one chain's thread group: the gradient (as in callback mode), the velocity
 * v = M^-1 p (velocity(): a matrix-vector product per chain, :407-412, :449-451), and the momentum draw
 * p0 = potential.random() (a triangular solve / product with the Cholesky factor, :414-417, :453-456); an adaptive
 * matrix is also updated outside (update(): rank-1 covariance update + Cholesky, :528-554).  The state machine is the
 * callback-mode one with more evaluation points; after every lmc_dense_advance each chain publishes in need[chain]
 * which results it waits for, the caller computes them for those chains and calls advance again:
 *
 *   LMC_NEED_UPDATE  potential.update(q[chain]) -- the transition that just ended was a tuning one (do this first)
 *   LMC_NEED_MOM     p0_eval[chain]   = potential.random() with the standard normals n_eval[chain]
 *   LMC_NEED_GRAD    logp_eval[chain], g_eval[chain] = logp_dlogp_func(q_eval[chain])
 *   LMC_NEED_VEL     v_eval[chain, r] = velocity(x_eval[chain, r]) for r = 0 (a momentum) and r = 1 (a gradient)
 *
 * One leapfrog costs one gradient and one two-vector velocity evaluation: the velocity at the half-kicked momentum
 * p + dt g (integration.py:108-111) is formed as velocity(p) + dt velocity(g) -- velocity() is linear, so this is the
 * same vector up to rounding -- which is why every state carries w = velocity(g) next to v = velocity(p).  Stack
 * entries and trajectory edges store their velocities (a dense velocity cannot be recomputed in place).  Everything
 * else in `base` means what it means for lmc_nuts_sample; base.var, the Welford arrays, base.target and
 * base.workspace are ignored; dual averaging of the step size stays inside the kernel.
 */
#define LMC_NEED_GRAD 1
#define LMC_NEED_VEL 2
#define LMC_NEED_MOM 4
#define LMC_NEED_UPDATE 8
/* IN (set by the caller in need[chain] before lmc_dense_advance): "not served yet" -- the chain is left exactly where it
 * is and keeps its request.  A caller whose potential.update is expensive per CALL rather than per chain (one batched
 * Cholesky for any number of matrices) holds the chains that ask for it until enough of them wait. */
#define LMC_NEED_HOLD 16

typedef struct lmc_dense_args {
  lmc_sampler_args base;
  double* q_eval;          /* [n_chains, ld]    out: position the gradient is needed at (padding = 0)              */
  const double* g_eval;    /* [n_chains, ld]    in                                                                 */
  const double* logp_eval; /* [n_chains]        in                                                                 */
  double* x_eval;          /* [n_chains, 2, ld] out: momentum (row 0) and gradient (row 1) whose velocity is needed */
  const double* v_eval;    /* [n_chains, 2, ld] in : velocity(x_eval rows)                                         */
  double* n_eval;          /* [n_chains, ld]    out: standard normals of the next momentum draw                    */
  const double* p0_eval;   /* [n_chains, ld]    in : potential.random() for those normals                          */
  int32_t* need;           /* [n_chains]        out: OR of LMC_NEED_* the chain waits for (0: finished); in: HOLD   */
  void* machine;           /* >= lmc_dense_state_bytes(...) bytes, 16-byte aligned, owned by the caller            */
  int64_t machine_bytes;
  int32_t* n_running;      /* device counter: chains that have not finished                                        */
} lmc_dense_args;

int64_t lmc_dense_state_bytes(int32_t kind, int32_t n_chains, int32_t ndim, int32_t max_treedepth);
/* Start base.n_trans transitions of every chain (need = GRAD | MOM everywhere). */
int lmc_dense_begin(int32_t kind, const lmc_dense_args* args);
/* Consume what each chain asked for and run it to its next evaluation point. */
int lmc_dense_advance(int32_t kind, const lmc_dense_args* args);

/* y[c, r, :] = A_c x[c, r, :] for the `n_idx` chains c = idx[i] (idx == NULL: chains 0..n_idx-1) and r < nrhs (1 or
 * 2): QuadPotentialFull(Adapt).velocity (quadpotential.py:449-451), the HBM-bound operation of dense-mass sampling
 * (8 D^2 bytes of matrix per chain and call; both right-hand sides share one pass over it).  A_c = A + c *
 * chain_stride (elements; 0 = one matrix shared by all chains), row-major [ndim, lda], lda even, 16-byte aligned.
 * x and y are [n_chains, nrhs, ld]. */
int lmc_dense_matvec(const int32_t* idx, int32_t n_idx, const double* A, int64_t chain_stride, int64_t lda,
                     int32_t ndim, int64_t ld, const double* x, double* y, int32_t nrhs, void* stream);

/* _WeightedCovariance.add_sample (quadpotential.py:607-613) on the foreground and background estimators of the
 * chains idx[0..n_idx): n_samples += 1; mean += (x - mean) / n_samples; raw_cov += new_diff old_diff^T; then
 * cov = raw_cov_fg / (n_samples_fg - 1) (current_covariance, :615-621 via _update_from_weightvar, :520-526).
 * x: [n_chains, ld] (the chains' positions); mean_*: [n_chains, ld]; raw_*, cov: [n_chains, ndim, lda];
 * nsamp: [n_chains, 2] float64 (fg, bg), updated in place.  cov == NULL skips the refresh (update_window > 1). */
int lmc_dense_cov_update(const int32_t* idx, int32_t n_idx, int32_t ndim, int64_t ld, int64_t lda, const double* x,
                         double* mean_fg, double* raw_fg, double* mean_bg, double* raw_bg, double* nsamp, double* cov,
                         void* stream);

/* Dump the numbers the PHILOX mode would consume into tapes (tests: Philox path == tape path == oracle).
 * normals: [n_chains, n_trans, ndim]; uniforms: [n_chains, n_trans, u_stride]. */
int lmc_rng_fill(const uint64_t* seeds, int32_t n_chains, int32_t ndim, int64_t iter0, int32_t n_trans,
                 int64_t u_stride, double* normals, double* uniforms, void* stream);

/* Copy `height` rows of `width_bytes` from device memory (row pitch spitch) to pinned host memory (row pitch
 * dpitch) on `stream`: how sample() streams blocks of the [chains, draws, ndim] trace out while sampling continues
 * (the reference's `trace[:, i] = q`, sampling.py:513, followed by its host-side reshape :208). */
int lmc_memcpy2d_d2h(void* dst_host, int64_t dpitch, const void* src_device, int64_t spitch, int64_t width_bytes,
                     int64_t height, void* stream);

/* Per-chain moments of the draws, the reduction behind cross-chain diagnostics (the reference has
 * only the per-run warnings of base_hmc.py:202-230 and leaves convergence checks to the caller).  One pass over
 * trace[chain * chain_stride + t * draw_stride + i]: the n_draws draws of every chain are cut into n_seg contiguous
 * segments (the last one takes the remainder) and for each (chain, segment, i) the kernel writes the segment's mean
 * and centred sum of squares M2 = sum (x - mean)^2.  mean, m2: [n_chains, n_seg, ndim].  n_seg = 2 gives the half
 * chains of split R-hat.  HBM-bound: 8 bytes per draw element, read once. */
int lmc_chain_moments(const double* trace, int32_t n_chains, int32_t n_draws, int32_t ndim, int64_t chain_stride,
                      int64_t draw_stride, int32_t n_seg, double* mean, double* m2, void* stream);

/* Last CUDA error string seen by the library on this thread (host pointer, static storage). */
const char* lmc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* LMC_B200_H */
