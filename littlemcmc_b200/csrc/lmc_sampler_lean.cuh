// "Lean" NUTS sampler kernel: the same transitions as sampler_kernel<.., KIND_NUTS> (lmc_sampler.cuh), with only the
// live phase-space point (q, p, grad) in registers.  The mass-matrix diagonal `var` and the running p_sum of the subtree
// being assembled live in shared memory, and the subtree's left-edge momentum is never copied at all: it is referenced
// by the id of the stack vector that already holds it.
//
// Why.  sampler_kernel keeps six vectors in registers (q, p, g, var, cur_lp, cur_ps): at D = 1000 that is 168 registers
// x 128 threads per chain and three chains per SM, and its throughput still grows with the number of resident chains
// (1 / 2 / 3 chains per SM: 4.9 / 8.3 / 9.9 x 10^7 leapfrog/s).  With three vectors in registers the same 128 threads
// x 4 pairs need 128 registers: four chains per SM, +8% (1.07 x 10^8 in the same probe).  Measured and rejected on the
// way: 64 threads x 8 pairs (six chains per SM but twice the vector instructions per warp: 7.8 x 10^7), 256 x 2
// (8.0 x 10^7), five chains per SM at 96 registers (spills: 8.0 x 10^7).
//
// Everything observable is identical to sampler_kernel: same arithmetic per element, same reduction tree per group
// shape, same uniforms in the same order; tests/test_gpu_parity.py runs both against the oracle.
#pragma once
#include "lmc_sampler.cuh"

namespace lmc {

// shared-memory vectors of the lean kernel that are not tree scratch
enum { LS_VAR = 0, LS_CPS = 1, LS_COUNT = 2 };

template <class Target, int G, int NP>
__device__ __forceinline__ void lean_eval_energy_kick(const Target& tgt, Group<G>& grp, int D, int ldh,
                                                      const double2* s_var, const double2 (&q)[NP], double2 (&p)[NP],
                                                      double2 (&g)[NP], double dt, bool kick, double& energy, double& logp) {
  double pre[2] = {0.0, 0.0};
  if constexpr (Target::kPre > 0) {
    tgt.template pre<G, NP>(grp.lane, D, q, pre);
    grp.allreduce(pre);
  }
  double acc[2];
  acc[1] = tgt.template grad<G, NP>(grp.lane, D, ldh, q, g, pre);
  acc[0] = 0.0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    if (kick) p[k] = axpy2(p[k], dt, g[k]);
    acc[0] = dot2(acc[0], p[k], mul2(s_var[k * G], p[k]));
  }
  grp.allreduce(acc);
  logp = tgt.finish(acc[1], pre);
  energy = 0.5 * acc[0] - logp;
}

template <class Target, int G, int NP, int MINCTAS>
__global__ void __launch_bounds__(G, MINCTAS) sampler_lean_kernel(const lmc_sampler_args a, const Target tgt,
                                                                 const KernelCfg cfg) {
  static_assert(G >= 64, "the lean kernel owns one chain per CTA");
  constexpr int VS = G * NP;
  extern __shared__ double2 smem2[];
  const int lane = threadIdx.x;
  const int slot = blockIdx.x;
  double2* const s_var = smem2 + (size_t)LS_VAR * VS + lane;
  double2* const s_cps = smem2 + (size_t)LS_CPS * VS + lane;
  double* const red = reinterpret_cast<double*>(smem2 + (size_t)(LS_COUNT + cfg.n_smem_vecs) * VS);
  StackScalars* const ss = reinterpret_cast<StackScalars*>(red + 2 * Group<G>::kWarps * kRedSlots);
  Scratch<G, NP> sc;
  sc.sm = smem2 + (size_t)LS_COUNT * VS;
  sc.ws = reinterpret_cast<double2*>(reinterpret_cast<char*>(a.workspace) + sched_bytes(a.n_chains)) +
          (size_t)slot * cfg.ws_vecs * VS;
  sc.n_smem = cfg.n_smem_vecs;
  sc.lane = lane;
  __shared__ int s_pop[2];
  const auto var_f = [=](int k) { return s_var[k * G]; };
  const auto cps_f = [=](int k) { return s_cps[k * G]; };
  const auto set_cps = [=](int k, double2 v) { s_cps[k * G] = v; };
  const auto no_lp = [](int, double2) {};  // the merged subtree's left edge stays where it is: referenced by id
  Group<G> grp(lane, red);
  const SchedView sv = sched_view(a.workspace, a.n_chains, a.n_trans);
  const unsigned total_units = (unsigned)a.n_chains * (unsigned)a.n_trans;

  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  const int tail = vid_tail(scratch_depth(a));

  for (;;) {
    // ---- pop the next (chain, transition) unit (lmc_sampler.cuh: scheduler) -------------------------------------------
    int chain = -1, t = 0;
    if (lane == 0) {
      sched_pop(sv, total_units, (unsigned)a.n_chains, chain, t);
      s_pop[0] = chain;
      s_pop[1] = t;
    }
    __syncthreads();
    chain = s_pop[0];
    t = s_pop[1];
    if (chain == -1) break;
    bool dead = ((unsigned)chain & kDeadBit) != 0u;
    chain &= 0x7fffffff;
    const size_t row = (size_t)chain * a.n_trans + t;
    // output rows of this unit: computed where they are written (epilogue / dead-chain fill), not held across the tree
    auto stats_row = [&]() -> double* { return a.stats + row * LMC_NSTATS; };
    auto trace_row = [&]() -> double* {
      return a.trace + (size_t)chain * a.trace_chain_stride +
             (size_t)(t > a.trace_skip ? t - a.trace_skip : 0) * a.trace_draw_stride;
    };
    int status = 0;

    if (!dead) {
      double2 q[NP], p[NP], g[NP];
      load_row_cg<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
      mask_tail<G, NP>(lane, D, q);
      {
        double2 var[NP];
        load_row_cg<G, NP>(a.var + (size_t)chain * a.ld, lane, ldh, var);
        mask_tail<G, NP>(lane, D, var);
#pragma unroll
        for (int k = 0; k < NP; ++k) s_var[k * G] = var[k];
      }
      // (the adaptation scalars are loaded in the epilogue, where they are used: holding 18 registers of them across
      //  the whole tree made the compiler spill hot tree state instead)
      const uint64_t seed = (a.rng.mode == LMC_RNG_PHILOX) ? a.rng.seeds[chain] : 0ull;
      const long long it = a.iter0 + t;
      const bool tune = it < a.n_tune;
      const bool adapt_step = tune && a.adapt_step_size;
      unsigned uc = 0;
      double u_lane = 0.0;
      auto next_uniform = [&]() -> double {
        double u;
        if (a.rng.mode == LMC_RNG_TAPE) {
          if ((long long)uc < a.rng.u_stride) {
            u = a.rng.uniforms[row * a.rng.u_stride + uc];
          } else {
            u = 0.5;
            status |= LMC_STATUS_TAPE_EXHAUSTED;
          }
        } else {
          if ((uc & 31u) == 0u) u_lane = philox_uniform(seed, it, uc + (unsigned)(lane & 31));
          u = __shfl_sync(0xffffffffu, u_lane, (int)(uc & 31u));
        }
        ++uc;
        return u;
      };

      // ---- p0 = potential.random()  (quadpotential.py:221-224 / 374-376) ------------------------------------------
      {
        const double* normals_row = a.rng.mode == LMC_RNG_TAPE ? a.rng.normals + row * D : nullptr;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const int j = lane + k * G;
          double2 n = make_double2(0.0, 0.0);
          if (normals_row) {
            if (2 * j < D) n.x = normals_row[2 * j];
            if (2 * j + 1 < D) n.y = normals_row[2 * j + 1];
          } else if (2 * j < D) {
            n = philox_normal_pair(seed, it, (uint32_t)j);
          }
          const double2 vk = s_var[k * G];
          p[k].x = (2 * j < D) ? mul_rn(inv_sqrt_cold(vk.x), n.x) : 0.0;
          p[k].y = (2 * j + 1 < D) ? mul_rn(inv_sqrt_cold(vk.y), n.y) : 0.0;
        }
      }

      // ---- start = integrator.compute_state(q0, p0)  (integration.py:52-66) ----------------------------------------
      double E0, logp0;
      lean_eval_energy_kick<Target, G, NP>(tgt, grp, D, ldh, s_var, q, p, g, 0.0, false, E0, logp0);
      if (!isfinite(E0)) {
        status |= LMC_STATUS_BAD_INITIAL_ENERGY;
        dead = true;
      } else {
        double eps = exp_cold(__ldcg(a.adapt + (size_t)chain * LMC_ADAPT_STRIDE +
                                     (adapt_step ? LMC_ADAPT_LOG_STEP : LMC_ADAPT_LOG_BAR)));
        if (a.step_size_override) eps = __ldg(a.step_size_override + chain);
        bool diverging = false, reached_max = false;

        const int max_depth = (tune && it < 200) ? a.early_max_treedepth : a.max_treedepth;  // nuts.py:205-208
        TrajScalars tr{xf_zero(), xf_zero(), 0.0, E0, logp0, 0, 0};
        int reg_edge = 0;
        tree_init<G, NP>(sc, tail, q, p, g);
        reached_max = max_depth <= 0;
        for (int d = 0; d < max_depth; ++d) {  // nuts.py:212
          const int dir = (next_uniform() < 0.5) ? 1 : -1;
          if (reg_edge != 0 && reg_edge != dir) {
            const int base = (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              q[k] = sc.ld(tvid(tail, base + 0), k);
              p[k] = sc.ld(tvid(tail, base + 1), k);
              g[k] = sc.ld(tvid(tail, base + 2), k);
            }
          }
          const double eps_d = dir > 0 ? eps : -eps;
          const double dt = 0.5 * eps_d;
          unsigned free_slots = 0xffffffffu;
          int fail = 0;
          long long n_leaves = 0;
          int cur_lp_id = -1;  // scratch vector holding the subtree's left-edge momentum; -1: the subtree is one leaf (p)
          CurTree cur{xf_zero(), xf_zero(), 0.0, 0.0, kLeafProp};

          const unsigned n_leaf_total = 1u << d;
          for (unsigned i = 0; i < n_leaf_total; ++i) {
            // ---- leapfrog (integration.py:100-121), var from shared memory --------------------------------------------
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              p[k] = axpy2(p[k], dt, g[k]);
              q[k] = axpy2(q[k], eps_d, mul2(s_var[k * G], p[k]));
            }
            double E, logp;
            lean_eval_energy_kick<Target, G, NP>(tgt, grp, D, ldh, s_var, q, p, g, dt, true, E, logp);
            ++n_leaves;
            if (!leaf_scalars(E, logp, E0, a.Emax, tr.max_dE, cur)) {
              fail = 1;
              break;
            }
            if ((i & 1u) == 0u) {
              if (i + 1 < n_leaf_total) push_leaf<G, NP>(sc, ss, q, p, cur, free_slots);
              continue;  // (the single leaf of the first doubling stays in registers: cur_lp_id == -1)
            }
            // ---- odd leaf: merge with stack entry 0, then with level 1, 2, .. for every further trailing 1-bit of i
            // (nuts.py:387-417; the arithmetic is lmc_tree.cuh's, the operands are reached where this kernel keeps them:
            // var and the running p_sum in shared memory, the left edge by the id of the stack vector that holds it)
            if (merge_leaves<G, NP>(sc, grp, ss, var_f, PairArray<NP>{p}, PairArray<NP>{p}, set_cps, no_lp, cur, free_slots,
                                    next_uniform)) {
              fail = 2;
              break;
            }
            cur_lp_id = vid_stack(0, 0);
            unsigned jbits = i >> 1;
            int lvl = 1;
            while (jbits & 1u) {  // merge with stack entry lvl >= 1
              const int lp_id = cur_lp_id;
              if (merge_upper<G, NP>(sc, grp, ss, lvl, var_f, PairArray<NP>{p},
                                     [&](int k) { return sc.ld(lp_id, k); }, cps_f, set_cps, no_lp, cur, free_slots,
                                     next_uniform)) {
                fail = 2;
                break;
              }
              cur_lp_id = vid_stack(lvl, 0);
              jbits >>= 1;
              ++lvl;
            }
            if (fail) break;
            if (i + 1 < n_leaf_total) {  // push "cur" at level lvl >= 1
              if (cur.pslot == kLeafProp) {
                cur.pslot = __ffs(free_slots) - 1;
                free_slots &= ~(1u << cur.pslot);
#pragma unroll
                for (int k = 0; k < NP; ++k) sc.st(vid_prop(cur.pslot), k, q[k]);
              }
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                sc.st(vid_stack(lvl, 0), k, sc.ld(cur_lp_id, k));
                sc.st(vid_stack(lvl, 1), k, p[k]);
                sc.st(vid_stack(lvl, 2), k, s_cps[k * G]);
              }
              if (lane == 0) {
                ss->wm[lvl] = cur.w.m;
                ss->we[lvl] = cur.w.e;
                ss->am[lvl] = cur.a.m;
                ss->ae[lvl] = cur.a.e;
                ss->pE[lvl] = cur.pE;
                ss->plogp[lvl] = cur.plogp;
                ss->pslot[lvl] = cur.pslot;
              }
            }
          }
          ++tr.depth;
          tr.n_prop += n_leaves;
          if (fail) {
            diverging = (fail == 1);
            break;
          }
          // ---- top of _Tree.extend (nuts.py:321-340): T.left.p = vec(cur_lp_id) or p, T.p_sum = s_cps or p --------------
          {
            const bool single = cur_lp_id < 0;
            const int lp_id = cur_lp_id;
            if (extend_top_f<G, NP>(sc, grp, tail, dir, var_f, PairArray<NP>{q}, PairArray<NP>{p},
                                    [&](int k) { return single ? p[k] : sc.ld(lp_id, k); },
                                    [&](int k) { return single ? p[k] : s_cps[k * G]; }, cur, tr, next_uniform()))
              break;
          }
          if (d + 1 < max_depth) {
            const int base = (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              sc.st(tvid(tail, base + 0), k, q[k]);
              sc.st(tvid(tail, base + 1), k, p[k]);
              sc.st(tvid(tail, base + 2), k, g[k]);
            }
            reg_edge = dir;
          } else {
            reached_max = true;
          }
        }
        const double accept_stat = mean_tree_accept(tr);
#pragma unroll
        for (int k = 0; k < NP; ++k) q[k] = sc.ld(tvid(tail, T_PROPQ), k);  // hmc_step.end.q

        double* const ad = a.adapt + (size_t)chain * LMC_ADAPT_STRIDE;
        DualAvg da{__ldcg(ad + LMC_ADAPT_LOG_STEP), __ldcg(ad + LMC_ADAPT_LOG_BAR), __ldcg(ad + LMC_ADAPT_HBAR),
                   __ldcg(ad + LMC_ADAPT_COUNT), __ldcg(ad + LMC_ADAPT_MU)};
        WelfordScalars wel{__ldcg(ad + LMC_ADAPT_W_FG), __ldcg(ad + LMC_ADAPT_W_BG),
                           (long long)__ldcg(ad + LMC_ADAPT_NSAMPLES), (long long)__ldcg(ad + LMC_ADAPT_WINDOW)};
        if (adapt_step) dual_average_update(da, accept_stat, a.target_accept, a.gamma, a.k, a.t0);
        if (tune && a.adapt_mass) {
          const size_t off = (size_t)chain * a.ld;
          double2 var[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k) var[k] = s_var[k * G];
          welford_update<G, NP>(lane, D, ldh, a.mean_fg + off, a.rawvar_fg + off, a.mean_bg + off, a.rawvar_bg + off, q,
                                var, wel, a.window_multiplier);
          store_row<G, NP>(a.var + off, lane, ldh, var);
        }
        double* const trow = trace_row();
        double* const srow = stats_row();
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const int j = lane + k * G;
          // the trace is write-once streaming output: evict-first, so it does not push the tree scratch out of L2
          if (2 * j < D) __stcs(trow + 2 * j, q[k].x);
          if (2 * j + 1 < D) __stcs(trow + 2 * j + 1, q[k].y);
        }
        if (lane == 0) {
          srow[LMC_STAT_DEPTH] = (double)tr.depth;
          srow[LMC_STAT_TREE_SIZE] = (double)tr.n_prop;
          srow[LMC_STAT_ACCEPT] = accept_stat;
          srow[LMC_STAT_ENERGY] = tr.prop_E;
          srow[LMC_STAT_ENERGY_ERROR] = tr.prop_E - E0;
          srow[LMC_STAT_MAX_ENERGY_ERROR] = tr.max_dE;
          srow[LMC_STAT_MODEL_LOGP] = tr.prop_logp;
          srow[LMC_STAT_DIVERGING] = diverging ? 1.0 : 0.0;
          srow[LMC_STAT_TUNE] = tune ? 1.0 : 0.0;
          srow[LMC_STAT_STEP_SIZE] = exp_cold(da.log_step);
          srow[LMC_STAT_STEP_SIZE_BAR] = exp_cold(da.log_bar);
          srow[LMC_STAT_N_UNIFORMS] = (double)uc;
          srow[LMC_STAT_REACHED_MAX_TREEDEPTH] = reached_max ? 1.0 : 0.0;
        }
        store_row<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
        __syncthreads();  // every thread has read the adaptation scalars (this epilogue) before lane 0 overwrites them
        if (lane == 0) {
          ad[LMC_ADAPT_LOG_STEP] = da.log_step;
          ad[LMC_ADAPT_LOG_BAR] = da.log_bar;
          ad[LMC_ADAPT_HBAR] = da.hbar;
          ad[LMC_ADAPT_COUNT] = da.count;
          ad[LMC_ADAPT_W_FG] = wel.w_fg;
          ad[LMC_ADAPT_W_BG] = wel.w_bg;
          ad[LMC_ADAPT_NSAMPLES] = (double)wel.n_samples;
          ad[LMC_ADAPT_WINDOW] = (double)wel.window;
        }
      }
      if (lane == 0 && status) atomicOr(a.status + chain, status);
    }
    if (dead) {
      const double nan = CUDART_NAN;
      double* const trow = trace_row();
      double* const srow = stats_row();
      for (int e = lane; e < D; e += G) trow[e] = nan;
      if (lane == 0)
        for (int s = 0; s < LMC_NSTATS; ++s) srow[s] = nan;
    }
    __threadfence();
    __syncthreads();
    if (lane == 0 && completes_block(a, t)) report_block(a, t);
    if (lane == 0 && t + 1 < a.n_trans) sched_push(sv, (unsigned)a.n_chains, chain, t + 1, dead);
  }
}

template <class Target, int G, int NP, int MINCTAS>
int launch_lean(const lmc_sampler_args& a, const Target& tgt) {
  constexpr int VS = G * NP;
  auto kern = sampler_lean_kernel<Target, G, NP, MINCTAS>;
  int dev = 0, n_sm = 0, smem_optin = 0;
  LMC_CUDA(cudaGetDevice(&dev));
  LMC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  LMC_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  KernelCfg cfg;
  cfg.ws_vecs = ws_vecs_nuts(scratch_depth(a));
  const size_t red_bytes = 2 * Group<G>::kWarps * kRedSlots * sizeof(double) + sizeof(StackScalars);
  const size_t vec_bytes = (size_t)VS * sizeof(double2);
  const size_t fixed = red_bytes + LS_COUNT * vec_bytes;
  if (fixed > (size_t)smem_optin) return LMC_ERR_UNSUPPORTED;
  int occ0 = 0;
  LMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fixed));
  LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, kern, G, fixed));
  if (occ0 < 1) return LMC_ERR_UNSUPPORTED;
  const size_t per_cta = (size_t)(227 * 1024) / occ0 - 1024;
  const size_t cap = per_cta < (size_t)smem_optin ? per_cta : (size_t)smem_optin;
  int n_smem = cap > fixed ? (int)((cap - fixed) / vec_bytes) : 0;
  const int hot = vid_tail(scratch_depth(a));
  if (n_smem > hot) n_smem = hot;
  if (a.tune_smem_vecs >= 0) n_smem = a.tune_smem_vecs < hot ? a.tune_smem_vecs : hot;
  cfg.n_smem_vecs = n_smem;
  const size_t smem = fixed + (size_t)n_smem * vec_bytes;
  if (smem > (size_t)smem_optin) return LMC_ERR_UNSUPPORTED;
  LMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G, smem));
  if (occ < 1) return LMC_ERR_UNSUPPORTED;
  long long grid = (long long)n_sm * occ;
  if (a.tune_max_slots > 0 && grid > a.tune_max_slots) grid = a.tune_max_slots;
  if (grid > a.n_chains) grid = a.n_chains;
  if (grid < 1) grid = 1;
  const long long need = (long long)sched_bytes(a.n_chains) + grid * (long long)cfg.ws_vecs * (long long)vec_bytes;
  if (need > a.workspace_bytes) return LMC_ERR_WORKSPACE;
  if ((long long)a.n_chains * a.n_trans >= (1ll << 31)) return LMC_ERR_UNSUPPORTED;
  sched_init_kernel<<<(a.n_chains + 255) / 256, 256, 0, (cudaStream_t)a.stream>>>(a.workspace, a.n_chains, a.n_trans);
  kern<<<(unsigned)grid, G, smem, (cudaStream_t)a.stream>>>(a, tgt, cfg);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

}  // namespace lmc
