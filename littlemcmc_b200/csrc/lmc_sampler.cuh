// lmc_nuts_sample / lmc_hmc_sample: whole MCMC transitions (momentum draw, initial state, trajectory, both
// adaptations, trace + statistics) for thousands of independent chains in ONE launch.
//
// Execution model.  The unit of work is one transition of one chain, run by a thread group (lmc_device.cuh) that keeps
// the chain's live phase-space point (q, p, grad), its mass-matrix diagonal and every tree scalar in registers for
// the whole transition.  Chains never synchronise with each other, so there is no lock-step loss: a chain that needs
// 1023 leapfrogs for a draw does not hold up one that needs 3.  The grid is persistent (one resident group per
// "slot", fed by a FIFO of chains, see the scheduler below), so the tree scratch is per resident slot, not per chain.
// For 513..1024 dimensions the default is the lean variant of this kernel (lmc_sampler_lean.cuh).
//
// NUTS tree (reference nuts.py:251-435) is built with the iterative binary-counter stack of SURVEY.md A.1:
// leaf i is merged with stack level 0,1,.. for every trailing 1-bit of i, which visits merges in exactly the
// post-order of the reference's recursion and therefore consumes uniforms in the same order.  A stack entry
// keeps only left.p, right.p, p_sum (velocities are recomputed as var*p, bit-identical to the stored ones)
// and an index into a small pool of proposal-position vectors, so choosing a proposal moves an index, never
// a vector.  The hottest scratch vectors (level-0/1 entries and the first proposal slots) live in shared
// memory, the rest in an L2-resident global workspace.
#pragma once
#ifndef __CUDACC_RTC__
#include "lmc_common.h"
#endif
#include "lmc_device.cuh"
#include "lmc_tree.cuh"

namespace lmc {

enum { KIND_NUTS = 0, KIND_HMC = 1 };

struct KernelCfg {
  int n_smem_vecs;  // scratch vectors per group kept in shared memory
  int ws_vecs;      // scratch vectors per slot in the global workspace (ids are absolute: smem ids unused there)
};

template <int G>
__host__ __device__ constexpr int block_threads() { return G >= 64 ? G : 128; }

// ---- work scheduler ----------------------------------------------------------------------------------------------
// The unit of work is ONE transition of ONE chain.  Chains wait in a FIFO ring in the workspace header; a resident
// thread group pops a chain, runs its next transition, writes the chain's state back to HBM and pushes the chain
// to the tail.  Trees differ in size by orders of magnitude between chains and draws (1 .. 2^max_treedepth
// leapfrogs), and the number of chains is rarely a multiple of the resident groups: with a static chain -> group
// assignment the launch lasts as long as its unluckiest group, with the FIFO every group stays busy until the
// queue drains.  Results do not depend on the schedule: all randomness is a function of (chain seed, iteration).
//   header: unsigned head, tail, pad[2];  unsigned long long ring[n_chains] = (ticket + 1) << 32 | dead << 31 | unit,
//           unit = chain * n_trans + index of the chain's next transition within this call (< 2^31, checked at launch)
// A chain is in the ring at most once, so a ring of n_chains entries never overwrites an unread entry.  The entry says
// everything about the unit, so a pop is two dependent L2 round trips (ticket, entry) and a push two (ticket, publish).
struct SchedView {
  unsigned* ctr;             // [0] = head (pop tickets), [1] = tail (push tickets)
  unsigned long long* ring;  // [n_chains]
  unsigned n_trans;
};
__host__ __device__ inline size_t sched_bytes(int n_chains) {
  return (((size_t)16 + (size_t)n_chains * 12) + 255) & ~(size_t)255;
}
__host__ __device__ inline SchedView sched_view(void* workspace, int n_chains, int n_trans) {
  unsigned* c = reinterpret_cast<unsigned*>(workspace);
  return SchedView{c, reinterpret_cast<unsigned long long*>(c + 4), (unsigned)n_trans};
}
constexpr unsigned kDeadBit = 0x80000000u;

__device__ __forceinline__ void sched_decode(const SchedView& sv, unsigned long long v, int& chain, int& t) {
  const unsigned unit = (unsigned)v & 0x7fffffffu;
  const unsigned c = unit / sv.n_trans;
  t = (int)(unit - c * sv.n_trans);
  chain = (int)(c | ((unsigned)v & kDeadBit));
}
// A pop in three pieces, so that a kernel can issue the memory operations early and consume them late (the chunked CTA
// kernel takes the ticket of its NEXT unit when a transition starts and reads the entry while the epilogue runs):
//   ticket  position in the pop order;  peek  one look at the ticket's ring entry (0 when the launch has no such unit);
//   take    wait until the entry is there, consume it.  Taking a ticket early cannot deadlock: a group only ever WAITS in
//   take, after it has pushed its own chain, and every ticket below total_units is filled by some push.
__device__ __forceinline__ unsigned sched_ticket(const SchedView& sv) { return atomicAdd(&sv.ctr[0], 1u); }
__device__ __forceinline__ unsigned long long sched_peek(const SchedView& sv, unsigned ticket, unsigned total_units,
                                                         unsigned n_chains) {
  if (ticket >= total_units) return 0ull;
  return *(volatile unsigned long long*)(sv.ring + (ticket % n_chains));
}
// The ring slot is ZEROED once read: a pusher only ever writes into a consumed slot, so a pusher that stalls between
// taking its ticket and storing the entry cannot be lapped by the ticket one ring later (the late store used
// to overwrite the newer entry and the popper of that ticket spun forever).
__device__ __forceinline__ void sched_take(const SchedView& sv, unsigned ticket, unsigned long long v, unsigned total_units,
                                           unsigned n_chains, int& chain, int& t) {
  chain = -1;
  t = 0;
  if (ticket >= total_units) return;
  volatile unsigned long long* e = sv.ring + (ticket % n_chains);
  while ((unsigned)(v >> 32) != ticket + 1u) {  // only when the queue ran dry (fewer chains than groups, or the tail)
    __nanosleep(100);
    v = *e;
  }
  __threadfence();  // acquire: the previous owner's state writes are ordered before its push
  *e = 0ull;        // consumed
  sched_decode(sv, v, chain, t);
}
// One thread of a group takes the next unit: chain (bit 31 = dead flag) or -1 when the launch is over, and the index of
// the chain's next transition.
__device__ __forceinline__ void sched_pop(const SchedView& sv, unsigned total_units, unsigned n_chains, int& chain, int& t) {
  const unsigned h = sched_ticket(sv);
  sched_take(sv, h, sched_peek(sv, h, total_units, n_chains), total_units, n_chains, chain, t);
}
// Hand the chain back for its transition t_next (the caller fenced its state writes).
__device__ __forceinline__ void sched_push(const SchedView& sv, unsigned n_chains, int chain, int t_next, bool dead) {
  const unsigned tk = atomicAdd(&sv.ctr[1], 1u);
  const unsigned long long entry = ((unsigned long long)(tk + 1u) << 32) | (dead ? kDeadBit : 0u) |
                                   ((unsigned)chain * sv.n_trans + (unsigned)t_next);
  unsigned long long* slot = sv.ring + (tk % n_chains);
  while (atomicCAS(slot, 0ull, entry) != 0ull) __nanosleep(100);  // wait for the slot's previous entry to be consumed
}

// lmc_sampler_args.progress: after transition t of a chain is globally visible (the caller fenced), one thread of the
// group reports the kept-draw block it completes, if any
__device__ __forceinline__ bool completes_block(const lmc_sampler_args& a, int t) {
  if (!a.progress || t < a.trace_skip) return false;
  const int kept = t - a.trace_skip + 1;  // kept draws of this chain so far
  return kept % a.progress_block == 0 || t + 1 == a.n_trans;
}
__device__ __forceinline__ void report_block(const lmc_sampler_args& a, int t) {
  atomicAdd(a.progress + (t - a.trace_skip) / a.progress_block, 1);
}

static __global__ void sched_init_kernel(void* workspace, int n_chains, int n_trans) {
  const SchedView sv = sched_view(workspace, n_chains, n_trans);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    sv.ctr[0] = 0u;
    sv.ctr[1] = (unsigned)n_chains;
  }
  if (i < n_chains) sv.ring[i] = ((unsigned long long)(i + 1) << 32) | ((unsigned)i * (unsigned)n_trans);
}


// Instantiated (threads per chain, pairs per thread, min resident CTAs per SM).  The third column caps registers:
// 65536 / (block_threads * min_ctas) per thread.
#ifndef LMC_MC_128_4
#define LMC_MC_128_4 3
#endif
#ifndef LMC_MC_256_2
#define LMC_MC_256_2 2
#endif
#ifndef LMC_MC_32_1
#define LMC_MC_32_1 4
#endif
#ifndef LMC_MC_32_2
#define LMC_MC_32_2 4
#endif
#define LMC_SHAPES(X) \
  X(32, 1, LMC_MC_32_1) X(32, 2, LMC_MC_32_2) X(32, 4, 3) X(64, 4, 4) X(64, 8, 4) X(128, 2, 4) X(128, 4, LMC_MC_128_4) \
  X(256, 2, LMC_MC_256_2) X(256, 4, 1) X(512, 2, 1) X(512, 4, 1) X(1024, 4, 1)

template <int G, int NP>
__host__ __device__ constexpr int min_ctas() {
#define LMC_X(g, np, mc) if (G == g && NP == np) return mc;
  LMC_SHAPES(LMC_X)
#undef LMC_X
  return 1;
}

template <class Target, int G, int NP, int KIND>
__global__ void __launch_bounds__(block_threads<G>(), min_ctas<G, NP>()) sampler_kernel(const lmc_sampler_args a, const Target tgt,
                                                                      const KernelCfg cfg) {
  constexpr int BLOCK = block_threads<G>();
  constexpr int CPB = BLOCK / G;  // chains (groups) per block
  constexpr int VS = G * NP;      // pairs per scratch vector (full coverage: scratch needs no bounds checks)

  extern __shared__ double2 smem2[];
  const int gib = threadIdx.x / G;
  const int lane = threadIdx.x - gib * G;
  const int slot = blockIdx.x * CPB + gib;
  double* const red = reinterpret_cast<double*>(smem2 + (size_t)CPB * cfg.n_smem_vecs * VS) +
                      gib * (2 * Group<G>::kWarps * kRedSlots);
  StackScalars* const ss = reinterpret_cast<StackScalars*>(
      reinterpret_cast<double*>(smem2 + (size_t)CPB * cfg.n_smem_vecs * VS) + CPB * (2 * Group<G>::kWarps * kRedSlots)) + gib;
  Scratch<G, NP> sc;
  sc.sm = smem2 + (size_t)gib * cfg.n_smem_vecs * VS;
  sc.ws = reinterpret_cast<double2*>(reinterpret_cast<char*>(a.workspace) + sched_bytes(a.n_chains)) +
          (size_t)slot * cfg.ws_vecs * VS;
  sc.n_smem = cfg.n_smem_vecs;
  sc.lane = lane;
  __shared__ int s_pop[2];
  Group<G> grp(lane, red);
  const SchedView sv = sched_view(a.workspace, a.n_chains, a.n_trans);
  const unsigned total_units = (unsigned)a.n_chains * (unsigned)a.n_trans;

  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  const int tail = vid_tail(scratch_depth(a));

  for (;;) {
    // ---- pop the next (chain, transition) unit ---------------------------------------------------------------------
    int chain = -1, t = 0;
    if (lane == 0) {
      sched_pop(sv, total_units, (unsigned)a.n_chains, chain, t);
      if constexpr (G > 32) {
        s_pop[0] = chain;
        s_pop[1] = t;
      }
    }
    if constexpr (G == 32) {
      chain = __shfl_sync(0xffffffffu, chain, 0);
      t = __shfl_sync(0xffffffffu, t, 0);
    } else {
      __syncthreads();
      chain = s_pop[0];
      t = s_pop[1];
    }
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
      // chain state lives in HBM between transitions and may have been written by another SM: bypass L1 (ld.cg)
      double2 q[NP], p[NP], g[NP], var[NP];
      load_row_cg<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
      load_row_cg<G, NP>(a.var + (size_t)chain * a.ld, lane, ldh, var);
      mask_tail<G, NP>(lane, D, q);
      mask_tail<G, NP>(lane, D, var);

      // (the adaptation scalars are loaded in the epilogue, where they are used: holding 18 registers of them across
      //  the whole tree made the compiler spill hot tree state instead)
      const uint64_t seed = (a.rng.mode == LMC_RNG_PHILOX) ? a.rng.seeds[chain] : 0ull;

      const long long it = a.iter0 + t;  // BaseHMC.iter_count
      const bool tune = it < a.n_tune;
      const bool adapt_step = tune && a.adapt_step_size;  // base_hmc.py:151
      unsigned uc = 0;                                    // uniforms consumed by this transition
      double u_lane = 0.0;  // PHILOX: uniform number (uc & ~31) + (lane & 31) of this transition, refilled every 32
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
          // one Philox block per lane yields the next 32 uniforms of the stream; they are handed out by shuffle
          if ((uc & 31u) == 0u) u_lane = philox_uniform(seed, it, uc + (unsigned)(lane & 31));
          u = __shfl_sync(0xffffffffu, u_lane, (int)(uc & 31u));
        }
        ++uc;
        return u;
      };

      // ---- p0 = potential.random()  (quadpotential.py:221-224 / 374-376) ---------------------------------------
      draw_momentum<G, NP>(lane, D, a.rng.mode == LMC_RNG_TAPE ? a.rng.normals + row * D : nullptr, seed, it, var, p);

      // ---- start = integrator.compute_state(q0, p0)  (integration.py:52-66) -------------------------------------
      double E0, logp0;
      eval_energy<false>(tgt, grp, D, ldh, q, p, g, var, 0.0, E0, logp0);
      if (!isfinite(E0)) {  // base_hmc.py:145-148: the reference raises; we flag the chain and stop it (rows -> NaN)
        status |= LMC_STATUS_BAD_INITIAL_ENERGY;
        dead = true;
      } else {
        double eps = exp_cold(__ldcg(a.adapt + (size_t)chain * LMC_ADAPT_STRIDE +
                                     (adapt_step ? LMC_ADAPT_LOG_STEP : LMC_ADAPT_LOG_BAR)));  // step_sizes.py:58-69
        if (a.step_size_override) eps = __ldg(a.step_size_override + chain);  // step_rand hook, base_hmc.py:154-155

        double accept_stat, stat_a, stat_b, stat_energy, stat_energy_error, stat_c, stat_logp;
        bool diverging = false, reached_max = false;

        if constexpr (KIND == KIND_NUTS) {
          // ---- NUTS._hamiltonian_step + _Tree  (nuts.py:204-224, 251-435) ---------------------------------------
          const int max_depth = (tune && it < 200) ? a.early_max_treedepth : a.max_treedepth;  // nuts.py:205-208
          TrajScalars tr{xf_zero(), xf_zero(), 0.0, E0, logp0, 0, 0};
          int reg_edge = 0;  // which trajectory edge (q,p,g) currently sits in registers: 0 both (start), +1 R, -1 L
          tree_init<G, NP>(sc, tail, q, p, g);
          reached_max = max_depth <= 0;          // for/else of nuts.py:212-220 with an empty range
          for (int d = 0; d < max_depth; ++d) {  // nuts.py:212
            // logbern(log 0.5): log(u) < log(0.5) <=> u < 0.5 (log is monotone; the two can only disagree for the
            // single double adjacent to 0.5)                                                          nuts.py:213
            const int dir = (next_uniform() < 0.5) ? 1 : -1;
            if (reg_edge != 0 && reg_edge != dir) {  // fetch the edge we extend from (nuts.py:297 / 306)
              const int base = (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                q[k] = sc.ld(tvid(tail, base + 0), k);
                p[k] = sc.ld(tvid(tail, base + 1), k);
                g[k] = sc.ld(tvid(tail, base + 2), k);
              }
            }
            const double eps_d = dir > 0 ? eps : -eps;
            unsigned free_slots = 0xffffffffu;  // proposal-slot pool: bit s set = slot s free
            int fail = 0;                       // 1 = diverging, 2 = turning
            long long n_leaves = 0;
            double2 cur_lp[NP], cur_ps[NP];
            CurTree cur{xf_zero(), xf_zero(), 0.0, 0.0, kLeafProp};

            const unsigned n_leaf_total = 1u << d;
            for (unsigned i = 0; i < n_leaf_total; ++i) {  // leaves of _build_subtree in integration order
              double E, logp;
              leapfrog(tgt, grp, D, ldh, eps_d, q, p, g, var, E, logp);  // nuts.py:347
              ++n_leaves;
              if (!leaf_scalars(E, logp, E0, a.Emax, tr.max_dE, cur)) {
                fail = 1;
                break;
              }
              if ((i & 1u) == 0u) {
                // even leaf: it becomes stack entry 0 (the next leaf merges with it); the only even LAST leaf is the
                // single leaf of the first doubling, which stays in registers as the whole subtree
                if (i + 1 < n_leaf_total) {
                  push_leaf<G, NP>(sc, ss, q, p, cur, free_slots);
                  if constexpr (G == 32) __syncwarp();
                } else {
#pragma unroll
                  for (int k = 0; k < NP; ++k) cur_lp[k] = cur_ps[k] = p[k];
                }
                continue;
              }
              // odd leaf: merge with stack entry 0, then with level 1, 2, .. for every further trailing 1-bit of i
              // (nuts.py:387-417, post-order of the reference's recursion)
              if (merge_leaf_pair<G, NP>(sc, grp, ss, var, p, cur_lp, cur_ps, cur, free_slots, next_uniform())) {
                fail = 2;
                break;
              }
              unsigned jbits = i >> 1;
              int lvl = 1;
              while (jbits & 1u) {
                __builtin_assume(lvl >= 1);  // the level-0 branches of merge_level are dead here
                if (merge_level<G, NP>(sc, grp, ss, lvl, var, p, cur_lp, cur_ps, cur, free_slots, next_uniform())) {
                  fail = 2;
                  break;
                }
                jbits >>= 1;
                ++lvl;
              }
              if (fail) break;
              __builtin_assume(lvl >= 1);
              if (i + 1 < n_leaf_total) {  // push "cur" at level lvl (the last leaf's result stays in registers)
                // One writer.  Readers see it after at least one group barrier (the next leaf's energy reduction)
                // and finished reading the previous occupant before the barrier that preceded this point.
                push_cur<G, NP>(sc, ss, lvl, q, p, cur_lp, cur_ps, cur, free_slots);
                if constexpr (G == 32) __syncwarp();
              }
            }
            ++tr.depth;            // nuts.py:315
            tr.n_prop += n_leaves;  // :316
            if (fail) {            // :318-319 -> :216-217 (the subtree is discarded, no uniform is drawn)
              diverging = (fail == 1);
              break;
            }
            if (extend_top<G, NP>(sc, grp, tail, dir, var, q, p, cur_lp, cur_ps, cur, tr, next_uniform())) break;  // :340
            if (d + 1 < max_depth) {  // self.right / self.left = tree.right (:304 / :313)
              const int base = (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                sc.st(tvid(tail, base + 0), k, q[k]);
                sc.st(tvid(tail, base + 1), k, p[k]);
                sc.st(tvid(tail, base + 2), k, g[k]);
              }
              reg_edge = dir;
            } else {
              reached_max = true;  // the loop runs out: neither a divergence nor a U-turn (nuts.py:218-220)
            }
          }
          // _Tree.stats (nuts.py:419-435)
          accept_stat = mean_tree_accept(tr);
          stat_a = (double)tr.depth;
          stat_b = (double)tr.n_prop;
          stat_energy = tr.prop_E;
          stat_energy_error = tr.prop_E - E0;
          stat_c = tr.max_dE;
          stat_logp = tr.prop_logp;
#pragma unroll
          for (int k = 0; k < NP; ++k) q[k] = sc.ld(tvid(tail, T_PROPQ), k);  // hmc_step.end.q
        } else {
          // ---- HamiltonianMC._hamiltonian_step (hmc.py:140-182) -------------------------------------------------
          double2 q0[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k) q0[k] = q[k];
          const double path_length = next_uniform() * a.path_length;  // :141
          const int n_steps = hmc_n_steps(path_length, eps, a.max_steps);  // :142-143
          double E = E0, logp = logp0;
          for (int s = 0; s < n_steps; ++s) leapfrog(tgt, grp, D, ldh, eps, q, p, g, var, E, logp);  // :149-150
          double dE;
          diverging = hmc_energy_check(E0, E, a.Emax, dE, accept_stat);  // :154-164
          bool accepted = false;
          if (!diverging) accepted = !(next_uniform() >= accept_stat);  // :166 (no draw when diverging)
          if (!accepted) {
#pragma unroll
            for (int k = 0; k < NP; ++k) q[k] = q0[k];
          }
          stat_a = (double)n_steps;
          stat_b = path_length;
          stat_energy = E;  // end-of-trajectory values even when rejected (:173-181)
          stat_energy_error = dE;
          stat_c = accepted ? 1.0 : 0.0;
          stat_logp = logp;
        }

        // ---- step_adapt.update(accept_stat, adapt_step)  (step_sizes.py:71-92) ---------------------------------
        double* const ad = a.adapt + (size_t)chain * LMC_ADAPT_STRIDE;
        DualAvg da{__ldcg(ad + LMC_ADAPT_LOG_STEP), __ldcg(ad + LMC_ADAPT_LOG_BAR), __ldcg(ad + LMC_ADAPT_HBAR),
                   __ldcg(ad + LMC_ADAPT_COUNT), __ldcg(ad + LMC_ADAPT_MU)};
        WelfordScalars wel{__ldcg(ad + LMC_ADAPT_W_FG), __ldcg(ad + LMC_ADAPT_W_BG),
                           (long long)__ldcg(ad + LMC_ADAPT_NSAMPLES), (long long)__ldcg(ad + LMC_ADAPT_WINDOW)};
        if (adapt_step) dual_average_update(da, accept_stat, a.target_accept, a.gamma, a.k, a.t0);
        // ---- potential.update(end.q, end.q_grad, tune)  (quadpotential.py:231-245, 322-338) --------------------
        if (tune && a.adapt_mass) {
          const size_t off = (size_t)chain * a.ld;
          welford_update<G, NP>(lane, D, ldh, a.mean_fg + off, a.rawvar_fg + off, a.mean_bg + off, a.rawvar_bg + off, q,
                                var, wel, a.window_multiplier);
        }

        // ---- outputs: trace[:, i] = q (sampling.py:513) and the stats dict (base_hmc.py:185-188) ----------------
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
          srow[LMC_STAT_DEPTH] = stat_a;
          srow[LMC_STAT_TREE_SIZE] = stat_b;
          srow[LMC_STAT_ACCEPT] = accept_stat;
          srow[LMC_STAT_ENERGY] = stat_energy;
          srow[LMC_STAT_ENERGY_ERROR] = stat_energy_error;
          srow[LMC_STAT_MAX_ENERGY_ERROR] = stat_c;
          srow[LMC_STAT_MODEL_LOGP] = stat_logp;
          srow[LMC_STAT_DIVERGING] = diverging ? 1.0 : 0.0;
          srow[LMC_STAT_TUNE] = tune ? 1.0 : 0.0;
          srow[LMC_STAT_STEP_SIZE] = exp_cold(da.log_step);
          srow[LMC_STAT_STEP_SIZE_BAR] = exp_cold(da.log_bar);
          srow[LMC_STAT_N_UNIFORMS] = (double)uc;
          srow[LMC_STAT_REACHED_MAX_TREEDEPTH] = reached_max ? 1.0 : 0.0;
        }

        // ---- write the chain's state back (its next transition may run on another SM) ---------------------------
        store_row<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
        store_row<G, NP>(a.var + (size_t)chain * a.ld, lane, ldh, var);
        group_barrier<G>();  // every thread has read the adaptation scalars (this epilogue) before lane 0 overwrites them
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
      }  // finite initial energy
      if (lane == 0 && status) atomicOr(a.status + chain, status);
    }  // !dead
    if (dead) {  // a stopped chain: its remaining rows are NaN
      const double nan = CUDART_NAN;
      double* const trow = trace_row();
      double* const srow = stats_row();
      for (int e = lane; e < D; e += G) trow[e] = nan;
      if (lane == 0)
        for (int s = 0; s < LMC_NSTATS; ++s) srow[s] = nan;
    }

    // ---- push the chain back for its next transition ---------------------------------------------------------------
    __threadfence();  // release: this thread's state writes become visible before the push below
    group_barrier<G>();
    if (lane == 0 && completes_block(a, t)) report_block(a, t);
    if (lane == 0 && t + 1 < a.n_trans) sched_push(sv, (unsigned)a.n_chains, chain, t + 1, dead);
  }
}

#ifndef __CUDACC_RTC__
// ---- host side: pick (G, NP), shared-memory split and grid; launch ---------------------------------------------------
struct Shape { int G, NP; };

inline bool pick_shape(int ndim, int force_group, Shape* out) {
  const int pairs = (ndim + 1) / 2;
  // default: the narrowest group whose 4 pairs/thread cover the row (more work per thread amortises reductions)
  static const Shape table[] = {{32, 1}, {32, 2}, {32, 4}, {64, 4}, {128, 4}, {256, 4}, {512, 4}, {1024, 4}};
  static const Shape all[] = {
#define LMC_X(g, np, mc) {g, np},
      LMC_SHAPES(LMC_X)
#undef LMC_X
  };
  if (force_group) {
    for (const Shape& s : all)
      if (s.G == force_group && s.G * s.NP >= pairs) { *out = s; return true; }
    return false;
  }
  for (const Shape& s : table)
    if (s.G * s.NP >= pairs) { *out = s; return true; }
  return false;
}

// `kern`: the kernel to launch -- a __global__ function of this library, or a cudaKernel_t of a module compiled at run
// time for a user target (lmc_user.cu); `tgt`: host pointer to the kernel's by-value target argument.
template <int G, int NP, int KIND>
int launch_kernel(const void* kern, const lmc_sampler_args& a, const void* tgt) {
  constexpr int BLOCK = block_threads<G>();
  constexpr int CPB = BLOCK / G;
  constexpr int VS = G * NP;
  int dev = 0, n_sm = 0, smem_optin = 0;
  LMC_CUDA(cudaGetDevice(&dev));
  LMC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  LMC_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

  KernelCfg cfg;
  cfg.ws_vecs = KIND == KIND_NUTS ? ws_vecs_nuts(scratch_depth(a)) : 0;
  const size_t red_bytes = (size_t)CPB * (2 * Group<G>::kWarps * kRedSlots * sizeof(double) + sizeof(StackScalars));
  const size_t vec_bytes = (size_t)VS * sizeof(double2);
  // shared-memory policy: give each CTA an equal share of the SM for the CTAs the register file can hold, and
  // fill it with the hottest scratch vectors (at most the stack part: edges / p_sum stay global).
  int n_smem = 0;
  if (KIND == KIND_NUTS) {
    int occ0 = 0;
    LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, kern, BLOCK, red_bytes));
    if (occ0 < 1) occ0 = 1;
    const size_t per_cta = (size_t)(227 * 1024) / occ0 - 1024;  // 1 KB/CTA reserved by the driver
    const size_t cap = per_cta < (size_t)smem_optin ? per_cta : (size_t)smem_optin;
    n_smem = cap > red_bytes ? (int)((cap - red_bytes) / (CPB * vec_bytes)) : 0;
    const int hot = vid_tail(scratch_depth(a));
    if (n_smem > hot) n_smem = hot;
    if (a.tune_smem_vecs >= 0) n_smem = a.tune_smem_vecs < hot ? a.tune_smem_vecs : hot;
  }
  cfg.n_smem_vecs = n_smem;
  const size_t smem = red_bytes + (size_t)CPB * n_smem * vec_bytes;
  if (smem > (size_t)smem_optin) return LMC_ERR_UNSUPPORTED;
  LMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
  if (occ < 1) return LMC_ERR_UNSUPPORTED;

  long long blocks_needed = ((long long)a.n_chains + CPB - 1) / CPB;
  long long grid = (long long)n_sm * occ;  // persistent: every CTA is resident, so a group waiting on the ring never deadlocks
  if (a.tune_max_slots > 0 && grid * CPB > a.tune_max_slots) grid = (a.tune_max_slots + CPB - 1) / CPB;
  if (grid > blocks_needed) grid = blocks_needed;
  if (grid < 1) grid = 1;
  const long long need = (long long)sched_bytes(a.n_chains) + grid * CPB * (long long)cfg.ws_vecs * (long long)vec_bytes;
  if (need > a.workspace_bytes) return LMC_ERR_WORKSPACE;
  if ((long long)a.n_chains * a.n_trans >= (1ll << 31)) return LMC_ERR_UNSUPPORTED;  // 32-bit scheduler tickets
  sched_init_kernel<<<(a.n_chains + 255) / 256, 256, 0, (cudaStream_t)a.stream>>>(a.workspace, a.n_chains, a.n_trans);
  void* kargs[] = {const_cast<lmc_sampler_args*>(&a), const_cast<void*>(tgt), &cfg};
  LMC_CUDA(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(BLOCK), kargs, smem, (cudaStream_t)a.stream));
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

template <class Target, int G, int NP, int KIND>
int launch(const lmc_sampler_args& a, const Target& tgt) {
  return launch_kernel<G, NP, KIND>(reinterpret_cast<const void*>(sampler_kernel<Target, G, NP, KIND>), a, &tgt);
}

template <class Target, int KIND>
int dispatch_shape(const lmc_sampler_args& a, const Target& tgt) {
  Shape s;
  if (!pick_shape(a.ndim, a.tune_group, &s)) return LMC_ERR_UNSUPPORTED;
#define LMC_X(g, np, mc) \
  if (s.G == g && s.NP == np) return launch<Target, g, np, KIND>(a, tgt);
  LMC_SHAPES(LMC_X)
#undef LMC_X
  return LMC_ERR_UNSUPPORTED;
}
#endif  // !__CUDACC_RTC__

}  // namespace lmc
