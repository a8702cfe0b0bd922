// Callback mode: NUTS / HMC transitions around an EXTERNAL gradient (the user's logp_dlogp_func evaluated as a batched
// torch op on [n_chains, ndim] between two launches), reference integration.py:115 `logp, dlogp = logp_dlogp_func(q)`.
//
// Every chain is a resumable state machine.  One lmc_callback_advance launch gives each chain the gradient it asked
// for and runs it to its NEXT evaluation point:
//
//     INIT  (gradient at the transition's start position)  -> momentum draw, compute_state, first doubling, half-kick
//     LEAF  (gradient at the new leapfrog position)        -> finish the kick, energy, leaf + merges (+ end of the
//            doubling, U-turn checks, end of the transition with both adaptations, trace and statistics), half-kick
//     DONE  all n_trans transitions finished: the chain idles
//
// so the host loop is `while running: (logp, grad) = f(q_eval); advance()`.  Chains do NOT move in lock step: one chain
// may be at leaf 37 of draw 3 while its neighbour starts draw 5; a launch costs one gradient evaluation for everybody
// and the number of launches is max over chains of (leapfrogs + transitions), not the sum over draws of the largest
// tree.  Finished / diverged chains are masked, not waited for.  All tree arithmetic is the code of the fused kernel
// (lmc_tree.cuh), so both modes make bit-identical decisions given the same gradients.
#include <cuda_runtime.h>
#include <stdint.h>

#include "lmc_common.h"
#include "lmc_device.cuh"
#include "lmc_tree.cuh"

namespace lmc {

enum { KIND_NUTS = 0, KIND_HMC = 1 };
enum { PH_INIT = 0, PH_LEAF = 1, PH_DONE = 2 };

struct CbScalars {
  int phase, t, d, dir;
  int max_depth, last_dir, n_steps, pad0;
  unsigned i, uc, free_slots, pad1;
  long long n_leaves;
  double E0, logp0, eps, path_length;
  TrajScalars tr;
};
struct CbMachine {
  CbScalars s;
  StackScalars ss;
};
constexpr size_t kMachineBytes = (sizeof(CbMachine) + 255) & ~(size_t)255;

__host__ __device__ constexpr int cb_vecs(int kind, int max_depth) { return kind == KIND_NUTS ? ws_vecs_nuts(max_depth) + 1 : 2; }

template <int G>
__host__ __device__ constexpr int cb_block() { return G >= 64 ? G : 128; }

static bool cb_pick_shape(int ndim, int* G, int* NP) {
  const int pairs = (ndim + 1) / 2;
  // callback mode is launch-latency bound: spread a chain over as many threads as its row has pairs
  static const int table[][2] = {{32, 1}, {64, 1}, {128, 1}, {256, 1}, {256, 2}, {512, 2}, {512, 4}, {1024, 4}};
  for (auto& s : table)
    if (s[0] * s[1] >= pairs) { *G = s[0]; *NP = s[1]; return true; }
  return false;
}

__global__ void cb_begin_kernel(const lmc_callback_args c, size_t vec_off) {
  const lmc_sampler_args& a = c.base;
  const int chain = blockIdx.x;
  CbMachine* M = reinterpret_cast<CbMachine*>(reinterpret_cast<char*>(c.machine) + (size_t)chain * kMachineBytes);
  if (threadIdx.x == 0) {
    M->s.phase = a.n_trans > 0 ? PH_INIT : PH_DONE;
    M->s.t = 0;
    if (chain == 0) *c.n_running = a.n_trans > 0 ? a.n_chains : 0;
  }
  for (int e = threadIdx.x; e < (int)a.ld; e += blockDim.x)
    c.q_eval[(size_t)chain * a.ld + e] = e < a.ndim ? a.q[(size_t)chain * a.ld + e] : 0.0;
  (void)vec_off;
}

// Registers are capped at 128 per thread (512 resident threads per SM-quarter's worth of registers): the kernel is a
// chain of dependent L2 round trips per CTA, so what counts is how many chains are resident at once -- 1024 chains of
// 64 threads fit in ONE wave at 8 CTAs per SM (145 registers gave 6: a second wave of 136 CTAs doubled the launch)
template <int G>
__host__ __device__ constexpr int cb_min_blocks() { return cb_block<G>() >= 512 ? 1 : 512 / cb_block<G>(); }

template <int G, int NP, int KIND>
__global__ void __launch_bounds__(cb_block<G>(), cb_min_blocks<G>()) cb_advance_kernel(const lmc_callback_args c, size_t vec_off, int n_vecs) {
  constexpr int CPB = cb_block<G>() / G;
  constexpr int VS = G * NP;
  const lmc_sampler_args& a = c.base;
  __shared__ double red_s[CPB * 2 * Group<G>::kWarps * kRedSlots];
  const int gib = threadIdx.x / G;
  const int lane = threadIdx.x - gib * G;
  const int chain = blockIdx.x * CPB + gib;
  if (chain >= a.n_chains) return;  // G == 32: per-warp exit; G >= 64: CPB == 1, whole block exits together
  CbMachine* const M = reinterpret_cast<CbMachine*>(reinterpret_cast<char*>(c.machine) + (size_t)chain * kMachineBytes);
  // the rows every phase needs are requested BEFORE the machine state is looked at: one L2 round trip instead of two on
  // the critical path of a launch that is nothing but dependent round trips (a finished chain wastes three loads)
  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  const size_t off = (size_t)chain * a.ld;
  double2 q[NP], p[NP], g[NP], var[NP];
  load_row<G, NP>(c.q_eval + off, lane, ldh, q);
  load_row<G, NP>(c.g_eval + off, lane, ldh, g);
  load_row<G, NP>(a.var + off, lane, ldh, var);
  const double logp = c.logp_eval[chain];
  CbScalars s = M->s;
  if (s.phase == PH_DONE) return;
  Group<G> grp(lane, red_s + gib * (2 * Group<G>::kWarps * kRedSlots));
  Scratch<G, NP> sc;
  sc.sm = nullptr;
  sc.ws = reinterpret_cast<double2*>(reinterpret_cast<char*>(c.machine) + vec_off) + (size_t)chain * n_vecs * VS;
  sc.n_smem = 0;
  sc.lane = lane;
  StackScalars* const ss = &M->ss;
  const int V_P = KIND == KIND_NUTS ? ws_vecs_nuts(scratch_depth(a)) : 0;  // half-kicked momentum between launches
  const int V_Q0 = 1;                                                     // HMC: the transition's start position
  const int tail = vid_tail(scratch_depth(a));

  mask_tail<G, NP>(lane, D, q);
  mask_tail<G, NP>(lane, D, g);
  mask_tail<G, NP>(lane, D, var);

  double* const ad = a.adapt + (size_t)chain * LMC_ADAPT_STRIDE;
  const long long it = a.iter0 + s.t;  // BaseHMC.iter_count
  const bool tune = it < a.n_tune;
  const bool adapt_step = tune && a.adapt_step_size;  // base_hmc.py:151
  const uint64_t seed = (a.rng.mode == LMC_RNG_PHILOX) ? a.rng.seeds[chain] : 0ull;
  const size_t row = (size_t)chain * a.n_trans + s.t;
  int status = 0;
  auto next_uniform = [&]() -> double {
    double u;
    if (a.rng.mode == LMC_RNG_TAPE) {
      if ((long long)s.uc < a.rng.u_stride) {
        u = a.rng.uniforms[row * a.rng.u_stride + s.uc];
      } else {
        u = 0.5;
        status |= LMC_STATUS_TAPE_EXHAUSTED;
      }
    } else {
      u = philox_uniform(seed, it, s.uc);
    }
    ++s.uc;
    return u;
  };

  bool new_doubling = false, trans_end = false, dead = false, diverging = false, reached_max = false;
  double accept_stat = 0.0, stat_a = 0.0, stat_b = 0.0, stat_energy = 0.0, stat_energy_error = 0.0, stat_c = 0.0,
         stat_logp = 0.0;

  if (s.phase == PH_INIT) {
    // ---- p0 = potential.random(); start = integrator.compute_state(q0, p0)  (base_hmc.py:142-148) ----------------
    s.uc = 0;
    draw_momentum<G, NP>(lane, D, a.rng.mode == LMC_RNG_TAPE ? a.rng.normals + row * D : nullptr, seed, it, var, p);
    double k1[1] = {0.0};
#pragma unroll
    for (int k = 0; k < NP; ++k) k1[0] = dot2(k1[0], p[k], mul2(var[k], p[k]));
    grp.allreduce(k1);
    s.E0 = 0.5 * k1[0] - logp;  // integration.py:63-65
    s.logp0 = logp;
    if (!isfinite(s.E0)) {
      status |= LMC_STATUS_BAD_INITIAL_ENERGY;
      dead = true;
    } else {
      s.eps = exp(adapt_step ? ad[LMC_ADAPT_LOG_STEP] : ad[LMC_ADAPT_LOG_BAR]);  // step_sizes.py:58-69
      if (a.step_size_override) s.eps = a.step_size_override[chain];             // step_rand hook, base_hmc.py:154-155
      if constexpr (KIND == KIND_NUTS) {
        s.max_depth = (tune && it < 200) ? a.early_max_treedepth : a.max_treedepth;  // nuts.py:205-208
        s.tr = TrajScalars{xf_zero(), xf_zero(), 0.0, s.E0, logp, 0, 0};
        s.d = 0;
        s.last_dir = 0;
        tree_init<G, NP>(sc, tail, q, p, g);
        new_doubling = s.max_depth > 0;
        trans_end = !new_doubling;
        reached_max = trans_end;  // for/else of nuts.py:212-220 with an empty range
      } else {
        s.path_length = next_uniform() * a.path_length;             // hmc.py:141
        s.n_steps = hmc_n_steps(s.path_length, s.eps, a.max_steps);  // :142-143
        s.i = 0;
        s.dir = 1;
#pragma unroll
        for (int k = 0; k < NP; ++k) sc.st(V_Q0, k, q[k]);
      }
    }
  } else {
    // ---- second half of the leapfrog (integration.py:116-119) with the gradient that just arrived ------------------
    const double dt = 0.5 * (s.dir > 0 ? s.eps : -s.eps);
    double k1[1] = {0.0};
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      p[k] = axpy2(sc.ld(V_P, k), dt, g[k]);
      k1[0] = dot2(k1[0], p[k], mul2(var[k], p[k]));
    }
    grp.allreduce(k1);
    const double E = 0.5 * k1[0] - logp;
    if constexpr (KIND == KIND_NUTS) {
      // ---- _single_step / _build_subtree for leaf s.i of doubling s.d (nuts.py:344-417) ----------------------------
      double2 cur_lp[NP], cur_ps[NP];
      CurTree cur{xf_zero(), xf_zero(), 0.0, 0.0, kLeafProp};
      int fail = 0, lvl = 0;
      ++s.n_leaves;
      if (!leaf_init<NP>(E, logp, s.E0, a.Emax, p, s.tr.max_dE, cur, cur_lp, cur_ps)) {
        fail = 1;
      } else {
        unsigned jbits = s.i;
        while (jbits & 1u) {
          if (merge_level<G, NP>(sc, grp, ss, lvl, var, p, cur_lp, cur_ps, cur, s.free_slots, next_uniform())) {
            fail = 2;
            break;
          }
          jbits >>= 1;
          ++lvl;
        }
      }
      if (!fail && s.i + 1 < (1u << s.d)) {
        push_cur<G, NP>(sc, ss, lvl, q, p, cur_lp, cur_ps, cur, s.free_slots);
        ++s.i;
      } else {  // the doubling is over (nuts.py:315-340)
        ++s.tr.depth;
        s.tr.n_prop += s.n_leaves;
        if (fail) {
          diverging = (fail == 1);
          trans_end = true;
        } else if (extend_top<G, NP>(sc, grp, tail, s.dir, var, q, p, cur_lp, cur_ps, cur, s.tr, next_uniform())) {
          trans_end = true;
        } else if (s.d + 1 < s.max_depth) {
          const int base = (s.dir > 0 ? T_RQ : T_LQ);  // self.right / self.left = tree.right (:304 / :313)
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            sc.st(tvid(tail, base + 0), k, q[k]);
            sc.st(tvid(tail, base + 1), k, p[k]);
            sc.st(tvid(tail, base + 2), k, g[k]);
          }
          s.last_dir = s.dir;
          ++s.d;
          new_doubling = true;
        } else {
          trans_end = true;  // max_treedepth reached (nuts.py:218-220)
          reached_max = true;
        }
      }
    } else {
      ++s.i;
      if ((int)s.i >= s.n_steps) {  // hmc.py:151-181
        trans_end = true;
        double dE;
        diverging = hmc_energy_check(s.E0, E, a.Emax, dE, accept_stat);
        bool accepted = false;
        if (!diverging) accepted = !(next_uniform() >= accept_stat);
        if (!accepted) {
#pragma unroll
          for (int k = 0; k < NP; ++k) q[k] = sc.ld(V_Q0, k);
        }
        stat_a = (double)s.n_steps;
        stat_b = s.path_length;
        stat_energy = E;
        stat_energy_error = dE;
        stat_c = accepted ? 1.0 : 0.0;
        stat_logp = logp;
      }
    }
  }

  if constexpr (KIND == KIND_NUTS) {
    if (trans_end) {  // _Tree.stats (nuts.py:419-435)
      accept_stat = mean_tree_accept(s.tr);
      stat_a = (double)s.tr.depth;
      stat_b = (double)s.tr.n_prop;
      stat_energy = s.tr.prop_E;
      stat_energy_error = s.tr.prop_E - s.E0;
      stat_c = s.tr.max_dE;
      stat_logp = s.tr.prop_logp;
#pragma unroll
      for (int k = 0; k < NP; ++k) q[k] = sc.ld(tvid(tail, T_PROPQ), k);  // hmc_step.end.q
    }
  }

  double* const srow = a.stats + row * LMC_NSTATS;
  double* const trow = a.trace + (size_t)chain * a.trace_chain_stride + (size_t)s.t * a.trace_draw_stride;
  if (trans_end) {
    // ---- close BaseHMC._astep (base_hmc.py:161-190): both adaptations, trace, statistics, new position ------------
    DualAvg da{ad[LMC_ADAPT_LOG_STEP], ad[LMC_ADAPT_LOG_BAR], ad[LMC_ADAPT_HBAR], ad[LMC_ADAPT_COUNT], ad[LMC_ADAPT_MU]};
    if (adapt_step) dual_average_update(da, accept_stat, a.target_accept, a.gamma, a.k, a.t0);
    WelfordScalars wel{ad[LMC_ADAPT_W_FG], ad[LMC_ADAPT_W_BG], (long long)ad[LMC_ADAPT_NSAMPLES],
                       (long long)ad[LMC_ADAPT_WINDOW]};
    if (tune && a.adapt_mass) {
      welford_update<G, NP>(lane, D, ldh, a.mean_fg + off, a.rawvar_fg + off, a.mean_bg + off, a.rawvar_bg + off, q, var,
                            wel, a.window_multiplier);
      store_row<G, NP>(a.var + off, lane, ldh, var);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      if (2 * j < D) trow[2 * j] = q[k].x;
      if (2 * j + 1 < D) trow[2 * j + 1] = q[k].y;
    }
    store_row<G, NP>(a.q + off, lane, ldh, q);
    store_row<G, NP>(c.q_eval + off, lane, ldh, q);  // the next transition's INIT evaluates here
    group_barrier<G>();  // every lane has read the adaptation scalars before lane 0 overwrites them
    if (lane == 0) {
      srow[LMC_STAT_DEPTH] = stat_a;
      srow[LMC_STAT_TREE_SIZE] = stat_b;
      srow[LMC_STAT_ACCEPT] = accept_stat;
      srow[LMC_STAT_ENERGY] = stat_energy;
      srow[LMC_STAT_ENERGY_ERROR] = stat_energy_error;
      srow[LMC_STAT_MAX_ENERGY_ERROR] = stat_c;
      srow[LMC_STAT_MODEL_LOGP] = stat_logp;
      srow[LMC_STAT_DIVERGING] = diverging ? 1.0 : 0.0;
      srow[LMC_STAT_TUNE] = tune ? 1.0 : 0.0;
      srow[LMC_STAT_STEP_SIZE] = exp(da.log_step);
      srow[LMC_STAT_STEP_SIZE_BAR] = exp(da.log_bar);
      srow[LMC_STAT_N_UNIFORMS] = (double)s.uc;
      srow[LMC_STAT_REACHED_MAX_TREEDEPTH] = reached_max ? 1.0 : 0.0;
      ad[LMC_ADAPT_LOG_STEP] = da.log_step;
      ad[LMC_ADAPT_LOG_BAR] = da.log_bar;
      ad[LMC_ADAPT_HBAR] = da.hbar;
      ad[LMC_ADAPT_COUNT] = da.count;
      ad[LMC_ADAPT_W_FG] = wel.w_fg;
      ad[LMC_ADAPT_W_BG] = wel.w_bg;
      ad[LMC_ADAPT_NSAMPLES] = (double)wel.n_samples;
      ad[LMC_ADAPT_WINDOW] = (double)wel.window;
    }
    ++s.t;
    s.phase = s.t < a.n_trans ? PH_INIT : PH_DONE;
  } else if (dead) {
    // base_hmc.py:145-148: the reference raises; the chain is flagged and stops, its remaining rows are NaN
    const double nan = CUDART_NAN;
    for (int tt = s.t; tt < a.n_trans; ++tt) {
      double* tr2 = a.trace + (size_t)chain * a.trace_chain_stride + (size_t)tt * a.trace_draw_stride;
      for (int e = lane; e < D; e += G) tr2[e] = nan;
      if (lane == 0) {
        double* s2 = a.stats + ((size_t)chain * a.n_trans + tt) * LMC_NSTATS;
        for (int n = 0; n < LMC_NSTATS; ++n) s2[n] = nan;
      }
    }
    s.phase = PH_DONE;
  } else {
    if (new_doubling) {  // nuts.py:213 and the edge the new subtree grows from (:297 / :306)
      s.dir = (next_uniform() < 0.5) ? 1 : -1;
      if (s.last_dir != 0 && s.last_dir != s.dir) {
        const int base = (s.dir > 0 ? T_RQ : T_LQ);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          q[k] = sc.ld(tvid(tail, base + 0), k);
          p[k] = sc.ld(tvid(tail, base + 1), k);
          g[k] = sc.ld(tvid(tail, base + 2), k);
        }
      }
      s.i = 0;
      s.n_leaves = 0;
      s.free_slots = 0xffffffffu;
    }
    // ---- first half of the next leapfrog (integration.py:105-112): the callback evaluates at q_eval ----------------
    const double e = s.dir > 0 ? s.eps : -s.eps;
    const double dt = 0.5 * e;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      p[k] = axpy2(p[k], dt, g[k]);
      q[k] = axpy2(q[k], e, mul2(var[k], p[k]));
      sc.st(V_P, k, p[k]);
    }
    store_row<G, NP>(c.q_eval + off, lane, ldh, q);
    s.phase = PH_LEAF;
  }
  if (lane == 0) {
    M->s = s;
    if (status) atomicOr(a.status + chain, status);
    if (s.phase == PH_DONE) atomicSub(c.n_running, 1);
  }
}

template <int KIND>
static int cb_launch(const lmc_callback_args& c, size_t vec_off, int n_vecs, int G, int NP) {
  const lmc_sampler_args& a = c.base;
  cudaStream_t st = (cudaStream_t)a.stream;
#define LMC_CASE(gg, np)                                                                                     \
  if (G == gg && NP == np) {                                                                                 \
    constexpr int CPB = cb_block<gg>() / gg;                                                                 \
    cb_advance_kernel<gg, np, KIND><<<(a.n_chains + CPB - 1) / CPB, cb_block<gg>(), 0, st>>>(c, vec_off, n_vecs); \
    LMC_CUDA(cudaGetLastError());                                                                            \
    return LMC_OK;                                                                                           \
  }
  LMC_CASE(32, 1) LMC_CASE(64, 1) LMC_CASE(128, 1) LMC_CASE(256, 1) LMC_CASE(256, 2) LMC_CASE(512, 2) LMC_CASE(512, 4)
  LMC_CASE(1024, 4)
#undef LMC_CASE
  return LMC_ERR_UNSUPPORTED;
}

static int cb_check(int kind, const lmc_callback_args* c, int* G, int* NP, size_t* vec_off, int* n_vecs) {
  if (!c || (kind != KIND_NUTS && kind != KIND_HMC)) return LMC_ERR_BADARG;
  const lmc_sampler_args& a = c->base;
  if (a.abi_version != LMC_ABI_VERSION) return LMC_ERR_BADARG;
  if (a.n_chains < 0 || a.ndim < 1 || a.n_trans < 0 || a.ld < a.ndim || (a.ld & 1)) return LMC_ERR_BADARG;
  if (!a.q || !a.var || !a.adapt || !a.trace || !a.stats || !a.status) return LMC_ERR_BADARG;
  if (!c->q_eval || !c->g_eval || !c->logp_eval || !c->machine || !c->n_running) return LMC_ERR_BADARG;
  if (((uintptr_t)a.q | (uintptr_t)a.var | (uintptr_t)c->q_eval | (uintptr_t)c->g_eval | (uintptr_t)c->machine) & 15)
    return LMC_ERR_BADARG;
  if (a.adapt_mass) {
    if (!a.mean_fg || !a.rawvar_fg || !a.mean_bg || !a.rawvar_bg) return LMC_ERR_BADARG;
    if (((uintptr_t)a.mean_fg | (uintptr_t)a.rawvar_fg | (uintptr_t)a.mean_bg | (uintptr_t)a.rawvar_bg) & 15)
      return LMC_ERR_BADARG;
  }
  if (a.rng.mode == LMC_RNG_TAPE) {
    if (!a.rng.normals || !a.rng.uniforms || a.rng.u_stride < 1) return LMC_ERR_BADARG;
  } else if (a.rng.mode == LMC_RNG_PHILOX) {
    if (!a.rng.seeds) return LMC_ERR_BADARG;
  } else {
    return LMC_ERR_BADARG;
  }
  if (kind == KIND_NUTS) {
    if (a.max_treedepth < 1 || a.max_treedepth > kMaxDepth) return LMC_ERR_UNSUPPORTED;
    if (a.early_max_treedepth < 0 || a.early_max_treedepth > kMaxDepth) return LMC_ERR_UNSUPPORTED;
  } else if (a.max_steps < 1) {
    return LMC_ERR_BADARG;
  }
  if (a.trace_skip != 0 || a.progress) return LMC_ERR_UNSUPPORTED;  // single-launch host traces: fused kernels only
  if (!cb_pick_shape(a.ndim, G, NP)) return LMC_ERR_UNSUPPORTED;
  *n_vecs = cb_vecs(kind, scratch_depth(a));
  *vec_off = (size_t)a.n_chains * kMachineBytes;
  const size_t need = *vec_off + (size_t)a.n_chains * *n_vecs * (size_t)(*G * *NP) * sizeof(double2);
  if ((size_t)c->machine_bytes < need) return LMC_ERR_WORKSPACE;
  return LMC_OK;
}

}  // namespace lmc

extern "C" int64_t lmc_callback_state_bytes(int32_t kind, int32_t n_chains, int32_t ndim, int32_t max_treedepth) {
  int G, NP;
  if (n_chains < 0 || (kind != lmc::KIND_NUTS && kind != lmc::KIND_HMC)) return LMC_ERR_BADARG;
  if (kind == lmc::KIND_NUTS && (max_treedepth < 1 || max_treedepth > lmc::kMaxDepth)) return LMC_ERR_UNSUPPORTED;
  if (!lmc::cb_pick_shape(ndim, &G, &NP)) return LMC_ERR_UNSUPPORTED;
  return (int64_t)((size_t)n_chains * lmc::kMachineBytes +
                   (size_t)n_chains * lmc::cb_vecs(kind, max_treedepth) * (size_t)(G * NP) * sizeof(double2));
}

extern "C" int lmc_callback_begin(int32_t kind, const lmc_callback_args* c) {
  int G, NP, n_vecs;
  size_t vec_off;
  const int rc = lmc::cb_check(kind, c, &G, &NP, &vec_off, &n_vecs);
  if (rc != LMC_OK) return rc;
  if (c->base.n_chains == 0) return LMC_OK;
  lmc::cb_begin_kernel<<<c->base.n_chains, 128, 0, (cudaStream_t)c->base.stream>>>(*c, vec_off);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

extern "C" int lmc_callback_advance(int32_t kind, const lmc_callback_args* c) {
  int G, NP, n_vecs;
  size_t vec_off;
  const int rc = lmc::cb_check(kind, c, &G, &NP, &vec_off, &n_vecs);
  if (rc != LMC_OK) return rc;
  if (c->base.n_chains == 0) return LMC_OK;
  return kind == lmc::KIND_NUTS ? lmc::cb_launch<lmc::KIND_NUTS>(*c, vec_off, n_vecs, G, NP)
                                : lmc::cb_launch<lmc::KIND_HMC>(*c, vec_off, n_vecs, G, NP);
}

// ---- device-driven loop: WHILE conditional graph node around the caller's captured iteration(s) ---------------------------
namespace lmc {
__global__ void cb_loop_cond_kernel(cudaGraphConditionalHandle handle, const int32_t* n_running, int32_t* iters,
                                    long long max_iters) {
  const int it = ++(*iters);
  cudaGraphSetConditional(handle, (*n_running > 0 && (long long)it < max_iters) ? 1u : 0u);
}
}  // namespace lmc

struct lmc_callback_loop {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
};

extern "C" int lmc_callback_loop_create(void* body_graph, const int32_t* n_running, int32_t* iters, int64_t max_iters,
                                        lmc_callback_loop** loop_out) {
  if (!body_graph || !n_running || !iters || !loop_out || max_iters < 1) return LMC_ERR_BADARG;
  cudaGraph_t parent = nullptr;
  LMC_CUDA(cudaGraphCreate(&parent, 0));
  cudaGraphConditionalHandle handle;
  // default value 1 at every launch: the body runs at least once (do-while on the chains still running)
  LMC_CUDA(cudaGraphConditionalHandleCreate(&handle, parent, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams wp = {};
  wp.type = cudaGraphNodeTypeConditional;
  wp.conditional.handle = handle;
  wp.conditional.type = cudaGraphCondTypeWhile;
  wp.conditional.size = 1;
  cudaGraphNode_t while_node;
  LMC_CUDA(cudaGraphAddNode(&while_node, parent, nullptr, 0, &wp));
  cudaGraph_t body = wp.conditional.phGraph_out[0];
  cudaGraphNode_t child;
  LMC_CUDA(cudaGraphAddChildGraphNode(&child, body, nullptr, 0, (cudaGraph_t)body_graph));  // clones body_graph
  long long mi = (long long)max_iters;
  void* kargs[] = {&handle, &n_running, &iters, &mi};
  cudaKernelNodeParams kp = {};
  kp.func = (void*)lmc::cb_loop_cond_kernel;
  kp.gridDim = dim3(1);
  kp.blockDim = dim3(1);
  kp.kernelParams = kargs;
  cudaGraphNode_t cond;
  LMC_CUDA(cudaGraphAddKernelNode(&cond, body, &child, 1, &kp));
  cudaGraphExec_t exec = nullptr;
  LMC_CUDA(cudaGraphInstantiate(&exec, parent, 0));
  lmc_callback_loop* loop = new lmc_callback_loop{parent, exec};
  *loop_out = loop;
  return LMC_OK;
}

extern "C" int lmc_callback_loop_launch(lmc_callback_loop* loop, void* stream) {
  if (!loop) return LMC_ERR_BADARG;
  LMC_CUDA(cudaGraphLaunch(loop->exec, (cudaStream_t)stream));
  return LMC_OK;
}

extern "C" int lmc_callback_loop_destroy(lmc_callback_loop* loop) {
  if (!loop) return LMC_OK;
  cudaGraphExecDestroy(loop->exec);
  cudaGraphDestroy(loop->graph);
  delete loop;
  return LMC_OK;
}
