// Chunked warp-per-chain NUTS kernel: the sampler for small dimensions (ndim <= 256; BASELINE configs 2 and 4).
//
// Why a second kernel.  With one warp per chain the round-1 kernel (lmc_sampler.cuh) executed ~1300 instructions and
// two dependent warp reductions per leapfrog, at ~19 cycles per instruction: every leaf is a chain of
// leapfrog -> butterfly -> exp -> merge -> butterfly -> branchy bookkeeping, and with 1024 chains there are fewer than
// two warps per scheduler to hide any of it (profiles/r02a_cfg2_ncu_full.md: 13% issue slots, 0.06 of roofline).
//
// What is different.  Inside one subtree the leapfrog TRAJECTORY does not depend on any tree decision (decisions only
// pick proposals and stop early), so the tree is built in CHUNKS of B = 2^b consecutive leaves:
//   1. B leapfrogs back to back, state in registers, NO reduction in between (targets whose gradient needs a sum,
//      Target::kPre > 0, keep that one); every leaf's momentum goes to a shared-memory ring, its position to a global
//      ring, and the per-lane partial sums of its kinetic energy / log density to a shared-memory table;
//   2. all dot products of all B-1 merges INSIDE the chunk (levels 0 .. b-1 of reference nuts.py:387-398) are formed
//      from the ring, every lane on its own columns, partials into the same table;
//   3. ONE transposed reduction: lane r sums row r of the table (no shuffles), energies -> exp() on B lanes in
//      parallel, U-turn flags -> one ballot;
//   4. the chunk's leaves and merges are resolved by all lanes at once: the first failing event (diverging leaf, turning
//      merge) in the reference's post-order is a warp minimum over closed-form event numbers, the subtree of a chunk
//      without one is a log2(B)-step tree over lanes (same uniforms in the same order as the recursion; the leaves after
//      a failure are discarded, and were the only wasted work);
//   5. the chunk's subtree (level b) is merged with the stack levels >= b exactly as before (lmc_tree.cuh merge_level,
//      vectors in the L2-resident workspace: touched once per B leaves).
// Same arithmetic per element and the same decisions as the reference; dot products are summed in a fixed but
// different order than the butterfly (like BLAS ddot, unspecified), so results agree with the oracle to the same
// ~1e-13 as the other kernels.  Shared memory per resident chain: B * NP * 512 B of momentum ring + table.
#pragma once
#include "lmc_sampler.cuh"

namespace lmc {

struct WarpCfg {
  int slot_bytes;  // shared memory per warp slot
  int ws_vecs;     // global scratch vectors per slot (tree stack + trajectory + position ring)
  int n_smem_vecs; // tree-scratch vectors kept in shared memory (ids < n_smem_vecs)
};

// vectors of global scratch per slot: tree stack + trajectory (as the other kernels) + the position ring of one chunk
__host__ __device__ constexpr int ws_vecs_warp(int sdepth, int chunk) { return ws_vecs_nuts(sdepth) + chunk; }

template <int NP, int B>
struct WarpLayout {
  static constexpr int VS = 32 * NP;                 // pairs per vector
  static constexpr int kRows = 6 * B - 6;            // table rows: 2B energy partials + (4B - 6) merge dot products
  static constexpr int kLog = (B == 2 ? 1 : B == 4 ? 2 : B == 8 ? 3 : 4);
  static_assert(B == 2 || B == 4 || B == 8 || B == 16, "chunk of 2, 4, 8 or 16 leaves");
  static_assert(4 * B - 6 >= 8 && 4 * B - 6 + 2 >= 8, "TableGroup reuses 8 dot-product rows and 8 totals");
  // byte offsets inside a slot
  static constexpr int oRingP = 0;
  static constexpr int oPs = oRingP + B * VS * 16;
  static constexpr int oVar = oPs + (B / 2) * VS * 16;
  // sticky launches, NP <= 2: 1/sqrt(var) next to var, refreshed only when var changes (after tuning the momentum draw
  // of a chain that stays on its warp needs no IEEE sqrt + divide per element); NP = 4 would lose a resident chain per SM
  static constexpr bool kIstd = NP <= 2;
  static constexpr int oIstd = oVar + VS * 16;
  static constexpr int oPart = oIstd + (kIstd ? VS * 16 : 0);
  static constexpr int kRowLd = 34;                  // doubles per table row (32 + 2: conflict-free 128-bit row reads)
  static constexpr int oVal = oPart + kRows * kRowLd * 8;
  // val: E[B], logp[B], dE[B], wm[B], pre[2B], dot[4B-6 (+pad)], local stack (kLog + 1) x 5 doubles, we[B] ints, lpidx ints
  static constexpr int nValD = 6 * B + (4 * B - 6 + 2) + 5 * (kLog + 1);
  static constexpr int oInts = oVal + nValD * 8;
  static constexpr int nInts = B + 3 * (kLog + 1) + 1;
  static constexpr int oSS = ((oInts + nInts * 4) + 15) & ~15;
  // sticky launches: the chain's adaptation scalars and both step sizes stay here between its transitions (12 doubles)
  static constexpr int oKeep = ((oSS + (int)sizeof(StackScalars)) + 15) & ~15;
  static constexpr int kFixedBytes = ((oKeep + 12 * 8) + 15) & ~15;  // (staged target parameters,) scratch vectors follow
  // a target's per-dimension parameters (StageTraits) are staged after the fixed part where that costs no resident chain:
  // NP = 2 (7 chains per SM with or without the extra KB); NP = 1 would drop from 10 to 9 chains per SM
  static constexpr bool kStageOk = NP == 2;
};
template <class Target, int NP, int B>
__host__ __device__ constexpr int warp_stage_vecs() {
  return WarpLayout<NP, B>::kStageOk ? StageTraits<Target>::kVecs : 0;
}
// the same number as a device variable: how the host learns it for a kernel compiled at run time around a user target
// (lmc_user.cu reads it from the loaded module; the slot size of the launch must match what the kernel was built with)
template <class Target, int NP, int B>
__device__ const int warp_stage_probe = warp_stage_vecs<Target, NP, B>();

// The warp's all-reduce through the shared-memory table instead of a shuffle butterfly: lane n stores value n of every
// lane into row n, lane r sums row r, everybody reads the N totals back.  Same latency as the butterfly (two warp
// barriers and ~50 instructions instead of 5 x N shuffle pairs), but 4x less CODE per call site -- what matters here:
// the kernel's warm footprint was 42 KB against a 32 KB instruction cache, 14 KB of it unrolled butterflies
// (profiles/r02z_cfg2_ncu_full.md: no_instructions 26% of the stall samples).
template <int LD>
struct TableGroup {
  int lane;
  double* rows;  // [8][LD] partials (LD >= 34: conflict-free 128-bit row reads)
  double* tot;   // [8] totals
  template <int N>
  __device__ __forceinline__ void allreduce(double (&v)[N]) {
    static_assert(N <= 8, "TableGroup reduces up to 8 values");
#pragma unroll
    for (int n = 0; n < N; ++n) rows[n * LD + lane] = v[n];
    __syncwarp();
    if (lane < N) {
      const double2* r = reinterpret_cast<const double2*>(rows + lane * LD);
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const double2 x = r[k], y = r[k + 1];
        s0 += x.x;
        s1 += x.y;
        s2 += y.x;
        s3 += y.y;
      }
      tot[lane] = (s0 + s1) + (s2 + s3);
    }
    __syncwarp();
#pragma unroll
    for (int n = 0; n < N; ++n) v[n] = tot[n];
    __syncwarp();  // the totals are read before the next reduction overwrites them
  }
};

// LMC_WARP_LANE_TREE 1 (default): the scalar pass over a chunk is resolved by all lanes at once (phase 4 below);
// 0: the sequential walk in the reference's post-order (kept for comparison: same results)
#ifndef LMC_WARP_LANE_TREE
#define LMC_WARP_LANE_TREE 1
#endif
// LMC_WARP_TIMING: lane 0 of every CTA's first warp accumulates clock64() deltas per phase (index = the phase that just
// ENDED) and block 0 prints its totals -- a development probe (-DLMC_WARP_TIMING=1 variant build, tools/quick_bench.py)
#ifdef LMC_WARP_TIMING
#include <stdio.h>
#define LMC_WTICK(i)                                 \
  if (threadIdx.x == 0) {                            \
    const long long now_ = clock64();                \
    s_tacc[i] += now_ - s_tacc[15];                  \
    s_tacc[15] = now_;                               \
  }
#else
#define LMC_WTICK(i)
#endif

template <class Target, int NP, int B, int WPB, int MINB, bool TAPE>
__global__ void __launch_bounds__(32 * WPB, MINB) sampler_warp_kernel(const lmc_sampler_args a, const Target tgt_in,
                                                                      const WarpCfg cfg) {
  using LY = WarpLayout<NP, B>;
  constexpr int G = 32;
  constexpr int VS = LY::VS;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double2 smem2[];
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * WPB + wib;
  char* const base = reinterpret_cast<char*>(smem2) + (size_t)wib * cfg.slot_bytes;
  double2* const ring_p = reinterpret_cast<double2*>(base + LY::oRingP) + lane;  // [B][NP][32]
  double2* const psbuf = reinterpret_cast<double2*>(base + LY::oPs) + lane;     // [B/2][NP][32]
  double2* const s_var = reinterpret_cast<double2*>(base + LY::oVar) + lane;    // [NP][32]
  double2* const s_istd = reinterpret_cast<double2*>(base + LY::oIstd) + lane;  // [NP][32] (kIstd)
  double* const part = reinterpret_cast<double*>(base + LY::oPart);             // [kRows][34]: row = value, column = lane
  double* const vE = reinterpret_cast<double*>(base + LY::oVal);
  double* const vLogp = vE + B;
  double* const vdE = vLogp + B;
  double* const vWm = vdE + B;
  double* const vPre = vWm + B;        // [B][2]
  double* const vDot = vPre + 2 * B;   // [4B - 6]
  double* const lstk = vDot + (4 * B - 6 + 2);  // local stack: [kLog + 1][5] = wm, am, pE, plogp, (unused)
  int* const vWe = reinterpret_cast<int*>(base + LY::oInts);  // [B]
  int* const lstk_i = vWe + B;                                // [kLog + 1][3] = we, ae, pidx
  StackScalars* const ss = reinterpret_cast<StackScalars*>(base + LY::oSS);
  double* const keep = reinterpret_cast<double*>(base + LY::oKeep);  // [0..8] adaptation scalars, [10] exp(log_step), [11] exp(log_bar)
  constexpr int kStage = warp_stage_vecs<Target, NP, B>();
  Scratch<G, NP> sc;
  sc.sm = reinterpret_cast<double2*>(base + LY::kFixedBytes + kStage * VS * 16);
  sc.ws = reinterpret_cast<double2*>(reinterpret_cast<char*>(a.workspace) + sched_bytes(a.n_chains)) +
          (size_t)slot * cfg.ws_vecs * VS;
  sc.n_smem = cfg.n_smem_vecs;
  sc.lane = lane;
  // every all-reduce of this kernel goes through the table (TableGroup), on rows that are idle at that point:
  //   grp   the two-value reductions inside a leapfrog (a target's pre-sums, the initial energy): the dot-product rows,
  //         which phase 1 does not touch;   tgrp  the six-value reductions of the stack merges and of extend (once per
  //         chunk / doubling, after the chunk's table has been consumed): its first rows.  Totals land in vDot.
  TableGroup<LY::kRowLd> grp{lane, part + 2 * B * LY::kRowLd, vDot};
  TableGroup<LY::kRowLd> tgrp{lane, part, vDot};
  const SchedView sv = sched_view(a.workspace, a.n_chains, a.n_trans);
  const unsigned total_units = (unsigned)a.n_chains * (unsigned)a.n_trans;

  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  // the target as this warp evaluates it: its per-dimension parameters copied to shared memory once per launch where the
  // slot has room (an L1-missing read-only load at the head of every doubling otherwise), else the caller's object
  using EvalTarget = typename Cond<(kStage > 0), typename StageTraits<Target>::Staged, Target>::type;
  const EvalTarget tgt = [&]() -> EvalTarget {
    if constexpr (kStage > 0) {
      const EvalTarget e = StageTraits<Target>::template make<NP>(
          tgt_in, reinterpret_cast<double2*>(base + LY::kFixedBytes) + lane, lane, ldh);
      __syncwarp();
      return e;
    } else {
      return tgt_in;
    }
  }();
  const int sdepth = scratch_depth(a);
  const int tail = vid_tail(sdepth);
  // position ring of the current chunk: B vectors after the tree scratch of this slot (global, written once per leaf,
  // read back only for the one position per chunk that survives as a proposal)
  double2* const ring_q = sc.ws + (size_t)ws_vecs_nuts(sdepth) * VS + lane;
  auto skew = [&](int row) -> double* { return part + row * LY::kRowLd + lane; };
  // sum of table row `row` over the 32 lanes' partials (fixed order; four interleaved accumulators)
  auto row_sum = [&](int row) -> double {
    const double2* r = reinterpret_cast<const double2*>(part + row * LY::kRowLd);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const double2 x = r[k], y = r[k + 1];
      s0 += x.x;
      s1 += x.y;
      s2 += y.x;
      s3 += y.y;
    }
    return (s0 + s1) + (s2 + s3);
  };

#ifdef LMC_WARP_TIMING
  __shared__ long long s_tacc[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 15; ++i) s_tacc[i] = 0;
    s_tacc[15] = clock64();
  }
#endif
  // "sticky" launches: when every chain has its own resident warp there is nothing to schedule -- warp `slot` runs all
  // the transitions of chain `slot` back to back, position and mass matrix stay on chip, and the FIFO's atomics, fences
  // and state reloads (3 dependent L2 round trips + 2 membars per transition) disappear.  Same results either way.
  const bool sticky = (unsigned)a.n_chains <= gridDim.x * (unsigned)WPB;
  int t_next = 0;
  bool sticky_dead = false;
  double2 q[NP];
  uint64_t seed_keep = 0ull;  // sticky launches: the chain's seed is read once, not once per transition

  for (;;) {
    // ---- the next (chain, transition) unit: own chain (sticky) or popped from the FIFO (lmc_sampler.cuh: scheduler) ----
    int chain = -1, t = 0;
    if (sticky) {
      if (slot < a.n_chains && t_next < a.n_trans) {
        chain = slot | (sticky_dead ? (int)kDeadBit : 0);
        t = t_next++;
      }
    } else {
      if (lane == 0) sched_pop(sv, total_units, (unsigned)a.n_chains, chain, t);
      chain = __shfl_sync(FULL, chain, 0);
      t = __shfl_sync(FULL, t, 0);
    }
    if (chain == -1) break;
    LMC_WTICK(0);
    bool dead = ((unsigned)chain & kDeadBit) != 0u;
    chain &= 0x7fffffff;
    const size_t row = (size_t)chain * a.n_trans + t;
    auto stats_row = [&]() -> double* { return a.stats + row * LMC_NSTATS; };
    auto trace_row = [&]() -> double* {
      return a.trace + (size_t)chain * a.trace_chain_stride +
             (size_t)(t > a.trace_skip ? t - a.trace_skip : 0) * a.trace_draw_stride;
    };
    int status = 0;

    if (!dead) {
      double2 p[NP], g[NP];
      if (!sticky || t == 0) {
        load_row_cg<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
        mask_tail<G, NP>(lane, D, q);
        double2 var[NP];
        load_row_cg<G, NP>(a.var + (size_t)chain * a.ld, lane, ldh, var);
        mask_tail<G, NP>(lane, D, var);
#pragma unroll
        for (int k = 0; k < NP; ++k) s_var[k * 32] = var[k];
      }
      if (!TAPE && (!sticky || t == 0)) seed_keep = a.rng.seeds[chain];
      const uint64_t seed = TAPE ? 0ull : seed_keep;
      const long long it = a.iter0 + t;
      const bool tune = it < a.n_tune;
      const bool adapt_step = tune && a.adapt_step_size;
      unsigned uc = 0;
      double u_lane = 0.0;
      auto next_uniform = [&]() -> double {
        double u;
        if constexpr (TAPE) {
          if ((long long)uc < a.rng.u_stride) {
            u = a.rng.uniforms[row * a.rng.u_stride + uc];
          } else {
            u = 0.5;
            status |= LMC_STATUS_TAPE_EXHAUSTED;
          }
        } else {
          if ((uc & 31u) == 0u) u_lane = philox_uniform_cold(seed, it, uc + (unsigned)lane);
          u = __shfl_sync(FULL, u_lane, (int)(uc & 31u));
        }
        ++uc;
        return u;
      };

      // ---- p0 = potential.random()  (quadpotential.py:221-224 / 374-376) ------------------------------------------
      {
        const double* normals_row = TAPE ? a.rng.normals + row * D : nullptr;
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
          double2 is;
          if constexpr (LY::kIstd) {
            // var changed since this warp last formed 1/sqrt(var)?  (first transition of the launch, a chain popped from
            // the FIFO, or the previous transition updated the mass matrix)
            if (!sticky || t == 0 || (it - 1 < a.n_tune && a.adapt_mass)) {
              const double2 vk = s_var[k * 32];
              is = make_double2(inv_sqrt_cold(vk.x), inv_sqrt_cold(vk.y));
              s_istd[k * 32] = is;
            } else {
              is = s_istd[k * 32];
            }
          } else {
            const double2 vk = s_var[k * 32];
            is = make_double2(inv_sqrt_cold(vk.x), inv_sqrt_cold(vk.y));
          }
          p[k].x = (2 * j < D) ? mul_rn(is.x, n.x) : 0.0;
          p[k].y = (2 * j + 1 < D) ? mul_rn(is.y, n.y) : 0.0;
        }
      }

      LMC_WTICK(1);
      // ---- start = integrator.compute_state(q0, p0)  (integration.py:52-66) ----------------------------------------
      double E0, logp0;
      {
        double pre[2] = {0.0, 0.0};
        if constexpr (Target::kPre > 0) {
          tgt.template pre<G, NP>(lane, D, q, pre);
          grp.allreduce(pre);
        }
        double acc[2];
        acc[1] = tgt.template grad<G, NP>(lane, D, ldh, q, g, pre);
        acc[0] = 0.0;
#pragma unroll
        for (int k = 0; k < NP; ++k) acc[0] = dot2(acc[0], p[k], mul2(s_var[k * 32], p[k]));
        grp.allreduce(acc);
        logp0 = tgt.finish(acc[1], pre);
        E0 = 0.5 * acc[0] - logp0;
      }
      if (!isfinite(E0)) {
        status |= LMC_STATUS_BAD_INITIAL_ENERGY;
        dead = true;
      } else {
        // a sticky chain's step sizes were formed by its previous epilogue (for the statistics row): no L2 round trip and
        // no exp() at the head of the transition
        double eps = (sticky && t > 0) ? keep[adapt_step ? 10 : 11]
                                       : exp_cold(__ldcg(a.adapt + (size_t)chain * LMC_ADAPT_STRIDE +
                                                         (adapt_step ? LMC_ADAPT_LOG_STEP : LMC_ADAPT_LOG_BAR)));
        if (a.step_size_override) eps = __ldg(a.step_size_override + chain);
        bool diverging = false, reached_max = false;
        const int max_depth = (tune && it < 200) ? a.early_max_treedepth : a.max_treedepth;  // nuts.py:205-208
        TrajScalars tr{xf_zero(), xf_zero(), 0.0, E0, logp0, 0, 0};
        int reg_edge = 0;
        tree_init<G, NP>(sc, tail, q, p, g);
        reached_max = max_depth <= 0;
        for (int d = 0; d < max_depth; ++d) {  // nuts.py:212
          LMC_WTICK(2);
          const int dir = (next_uniform() < 0.5) ? 1 : -1;
          if (reg_edge != 0 && reg_edge != dir) {
            const int eb = (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              q[k] = sc.ld(tvid(tail, eb + 0), k);
              p[k] = sc.ld(tvid(tail, eb + 1), k);
              g[k] = sc.ld(tvid(tail, eb + 2), k);
            }
          }
          const double eps_d = dir > 0 ? eps : -eps;
          const double dt = 0.5 * eps_d;
          unsigned free_slots = 0xffffffffu;  // pool of proposal slots of the stack levels >= b
          int fail = 0;                       // 1 = diverging, 2 = turning
          long long n_leaves = 0;
          const unsigned n_leaf_total = 1u << d;
          const int Bc = n_leaf_total < (unsigned)B ? (int)n_leaf_total : B;  // leaves per chunk
          const int bc = 31 - __clz(Bc);                                     // its level
          const unsigned n_chunks = n_leaf_total / (unsigned)Bc;
          CurTree cur{xf_zero(), xf_zero(), 0.0, 0.0, kLeafProp};
          int pidx = 0;  // ring index of cur's proposal while cur.pslot == kLeafProp ("still in the position ring")

          for (unsigned c = 0; c < n_chunks && !fail; ++c) {
            LMC_WTICK(3);
            // ---- 1. Bc leapfrogs (integration.py:100-121), no reduction between them ------------------------------------
            for (int s = 0; s < Bc; ++s) {
              double pre[2] = {0.0, 0.0};
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                p[k] = axpy2(p[k], dt, g[k]);
                q[k] = axpy2(q[k], eps_d, mul2(s_var[k * 32], p[k]));
              }
              if constexpr (Target::kPre > 0) {
                tgt.template pre<G, NP>(lane, D, q, pre);
                grp.allreduce(pre);
                if (lane == 0) {
                  vPre[2 * s] = pre[0];
                  vPre[2 * s + 1] = pre[1];
                }
              }
              const double lp_part = tgt.template grad<G, NP>(lane, D, ldh, q, g, pre);
              double k_part = 0.0;
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                p[k] = axpy2(p[k], dt, g[k]);
                k_part = dot2(k_part, p[k], mul2(s_var[k * 32], p[k]));
                ring_p[(s * NP + k) * 32] = p[k];
                ring_q[(size_t)(s * NP + k) * 32] = q[k];
              }
              *skew(s) = k_part;
              *skew(B + s) = lp_part;
            }
            LMC_WTICK(4);
            // ---- 2. dot products of the merges inside the chunk (nuts.py:387-398), every lane on its own columns ------------
            if (Bc >= 2) {
              for (int k2 = 0; k2 < Bc / 2; ++k2) {  // level 0: two leaves
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  const double2 pa = ring_p[((2 * k2) * NP + k) * 32], pb = ring_p[((2 * k2 + 1) * NP + k) * 32];
                  const double2 vk = s_var[k * 32];
                  const double2 ps = add2(pa, pb);          // p_sum = tree1.p_sum + tree2.p_sum (:390)
                  d0 = dot2(d0, ps, mul2(vk, pa));          // p_sum . left.v
                  d1 = dot2(d1, ps, mul2(vk, pb));          // p_sum . right.v
                  psbuf[(k2 * NP + k) * 32] = ps;
                }
                *skew(2 * B + 2 * k2) = d0;
                *skew(2 * B + 2 * k2 + 1) = d1;
              }
              int row0 = 2 * B + Bc;
              for (int l = 1; (2 << l) <= Bc; ++l) {  // level l: two subtrees of `span` leaves each
                const int span = 1 << l;
                const int n_m = Bc / (2 * span);
                for (int k2 = 0; k2 < n_m; ++k2) {
                  const int first = k2 * 2 * span;
                  double d6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                  for (int k = 0; k < NP; ++k) {
                    const double2 t1_lp = ring_p[((first)*NP + k) * 32], t1_rp = ring_p[((first + span - 1) * NP + k) * 32];
                    const double2 t2_lp = ring_p[((first + span) * NP + k) * 32];
                    const double2 t2_rp = ring_p[((first + 2 * span - 1) * NP + k) * 32];
                    const double2 t1_ps = psbuf[((2 * k2) * NP + k) * 32], t2_ps = psbuf[((2 * k2 + 1) * NP + k) * 32];
                    const double2 vk = s_var[k * 32];
                    const double2 ps = add2(t1_ps, t2_ps);   // :390
                    const double2 ps1 = add2(t1_ps, t2_lp);  // tree1.p_sum + tree2.left.p (:394)
                    const double2 ps2 = add2(t1_rp, t2_ps);  // tree1.right.p + tree2.p_sum (:396)
                    const double2 v1l = mul2(vk, t1_lp), v1r = mul2(vk, t1_rp);
                    const double2 v2l = mul2(vk, t2_lp), v2r = mul2(vk, t2_rp);
                    d6[0] = dot2(d6[0], ps, v1l);
                    d6[1] = dot2(d6[1], ps, v2r);
                    d6[2] = dot2(d6[2], ps1, v1l);
                    d6[3] = dot2(d6[3], ps1, v2l);
                    d6[4] = dot2(d6[4], ps2, v1r);
                    d6[5] = dot2(d6[5], ps2, v2r);
                    psbuf[(k2 * NP + k) * 32] = ps;  // in place: entries 2 k2, 2 k2 + 1 >= k2 are not read again
                  }
#pragma unroll
                  for (int e = 0; e < 6; ++e) *skew(row0 + 6 * k2 + e) = d6[e];
                }
                row0 += 6 * n_m;
              }
            }
            LMC_WTICK(5);
            __syncwarp();
#if LMC_WARP_LANE_TREE
            // ---- 3. one transposed reduction: lane r sums row r ---------------------------------------------------------------
            double lE = 0.0, lLogp = 0.0, ldE = 0.0;  // the leaf of this lane (lane < Bc)
            XF lw = xf_zero();
            {
              // energies: rows s (kinetic) and B + s (log-density sums), s < Bc
              double sum = 0.0;
              if (lane < Bc || (lane >= B && lane < B + Bc)) sum = row_sum(lane);
              const double lp_s = __shfl_sync(FULL, sum, (lane + B) & 31);
              if (lane < Bc) {
                double pre[2] = {0.0, 0.0};
                if constexpr (Target::kPre > 0) {
                  pre[0] = vPre[2 * lane];
                  pre[1] = vPre[2 * lane + 1];
                }
                lLogp = tgt.finish(lp_s, pre);
                lE = 0.5 * sum - lLogp;
                ldE = lE - E0;                               // nuts.py:352
                if (isnan(ldE)) ldE = CUDART_INF;            // :353-354
                if (fabs(ldE) < a.Emax) lw = xf_exp(-ldE);   // log_size = -dE (:359)
              }
              // merge dot products: rows 2B .. 2B + 4 Bc - 7
              const int n_dot = Bc >= 2 ? 4 * Bc - 6 : 0;
              for (int r = lane; r < n_dot; r += 32) vDot[r] = row_sum(2 * B + r);
            }
            __syncwarp();
            // the merge this lane answers for (lane < Bc - 1): level ml, index midx, its U-turn flag
            bool mflag = false;
            int ml = 0, midx = 0;
            if (lane < Bc - 1) {
              int cnt = Bc / 2;
              midx = lane;
              while (midx >= cnt) {
                midx -= cnt;
                cnt >>= 1;
                ++ml;
              }
              if (ml == 0) {
                mflag = (vDot[2 * midx] <= 0) || (vDot[2 * midx + 1] <= 0);  // :391
              } else {
                int r0 = Bc;  // first row of level ml
                for (int ll = 1; ll < ml; ++ll) r0 += 6 * (Bc >> (ll + 1));
                const double* dd = vDot + r0 + 6 * midx;
                mflag = (dd[0] <= 0) || (dd[1] <= 0) || (dd[2] <= 0) || (dd[3] <= 0) || (dd[4] <= 0) || (dd[5] <= 0);  // :391-398
              }
            }
            LMC_WTICK(6);
            // ---- 4. the chunk's leaves and merges, resolved by all lanes at once -----------------------------------------------
            // The reference walks them in post-order (leaf 0, leaf 1, merge(0,1), leaf 2, ...) and stops at the first
            // diverging leaf or turning merge.  Event numbers in that order: leaf s -> 2s - popc(s); the merge at level l
            // completed by leaf i -> number(i) + 1 + l; its uniform is number i - popc(i) + l of the chunk.  So: the first
            // failing event is a warp minimum, what the walk would have counted up to it are ballots, and without a failure
            // the chunk's subtree is a log2(Bc)-step tree over lanes -- same values, same uniforms, same decisions.
            {
              const unsigned kNone = 0xffffffffu;
              const unsigned ulane = (unsigned)lane;
              const unsigned seq_leaf = 2u * ulane - (unsigned)__popc(ulane);
              const unsigned m_i = ((unsigned)(midx + 1) << (ml + 1)) - 1u;  // leaf that completes this lane's merge
              const unsigned m_rank = m_i - (unsigned)__popc(m_i) + (unsigned)ml;
              const unsigned seq_merge = m_i + m_rank + 1u;                    // = 2 m_i - popc(m_i) + 1 + ml
              const bool leaf_bad = lane < Bc && !(fabs(ldE) < a.Emax);      // :358 / :370-375
              unsigned my = leaf_bad ? seq_leaf : kNone;
              if (lane < Bc - 1 && mflag && seq_merge < my) my = seq_merge;
              const unsigned F = __reduce_min_sync(FULL, my);
              const bool leaf_in = lane < Bc && seq_leaf <= F;        // leaves the walk reaches
              const bool merge_in = lane < Bc - 1 && seq_merge <= F;  // merges it completes (a turning one draws its uniform)
              n_leaves += __popc(__ballot_sync(FULL, leaf_in));
              const unsigned n_u = (unsigned)__popc(__ballot_sync(FULL, merge_in));
              {  // max_energy_change over those leaves: the first one with the largest |dE| (:356-357)
                const unsigned long long key = leaf_in ? (unsigned long long)__double_as_longlong(fabs(ldE)) : 0ull;
                const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                const unsigned mh = __reduce_max_sync(FULL, hi);
                const unsigned mlo = __reduce_max_sync(FULL, (leaf_in && hi == mh) ? lo : 0u);
                const unsigned who = __ballot_sync(FULL, leaf_in && hi == mh && lo == mlo);
                const double cand = __shfl_sync(FULL, ldE, __ffs(who) - 1);
                if (fabs(cand) > fabs(tr.max_dE)) tr.max_dE = cand;
              }
              const unsigned uc0 = uc;
              if constexpr (TAPE) {
                if (__any_sync(FULL, merge_in && (long long)(uc0 + m_rank) >= a.rng.u_stride)) status |= LMC_STATUS_TAPE_EXHAUSTED;
              }
              if (F != kNone) {
                uc = uc0 + n_u;
                fail = __ballot_sync(FULL, leaf_bad && seq_leaf == F) ? 1 : 2;
                break;
              }
              // uniforms uc0 .. uc0 + Bc - 2 of this transition's stream
              const unsigned blk = uc0 & ~31u;
              double u_lo = 0.0, u_hi = 0.0;
              if constexpr (!TAPE) {
                if (Bc >= 2) {
                  if ((uc0 & 31u) == 0u) u_lane = philox_uniform_cold(seed, it, uc0 + ulane);  // the block starts here
                  u_lo = u_lane;
                  if (uc0 + (unsigned)Bc - 2u >= blk + 32u) u_hi = philox_uniform_cold(seed, it, blk + 32u + ulane);
                }
              }
              auto chunk_uniform = [&](unsigned r) -> double {  // called by all lanes
                const unsigned x = uc0 + r;
                if constexpr (TAPE) {
                  return (long long)x < a.rng.u_stride ? a.rng.uniforms[row * a.rng.u_stride + x] : 0.5;
                } else {
                  const double ua = __shfl_sync(FULL, u_lo, (int)(x & 31u)), ub = __shfl_sync(FULL, u_hi, (int)(x & 31u));
                  return x >= blk + 32u ? ub : ua;
                }
              };
              // every group of Bc lanes builds the same tree: lane L starts from leaf L mod Bc
              const int src = lane & (Bc - 1);
              XF nw{__shfl_sync(FULL, lw.m, src), __shfl_sync(FULL, lw.e, src)};
              const double sdE = __shfl_sync(FULL, ldE, src);
              XF na = (-sdE < 0.0) ? xf_sqr(nw) : nw;  // log_p_accept_weighted = -dE + min(0, -dE)  (:363)
              double npE = __shfl_sync(FULL, lE, src), nplogp = __shfl_sync(FULL, lLogp, src);
              int npidx = src;
              for (int l = 0; l < bc; ++l) {
                const int bit = 1 << l;
                const XF pw{__shfl_xor_sync(FULL, nw.m, bit), __shfl_xor_sync(FULL, nw.e, bit)};
                const XF pa{__shfl_xor_sync(FULL, na.m, bit), __shfl_xor_sync(FULL, na.e, bit)};
                const double ppE = __shfl_xor_sync(FULL, npE, bit), pplogp = __shfl_xor_sync(FULL, nplogp, bit);
                const int ppidx = __shfl_xor_sync(FULL, npidx, bit);
                const bool right = (lane & bit) != 0;  // this lane's node is tree2 of the merge
                const XF t2w = right ? nw : pw;
                const unsigned i_done = ((((unsigned)src >> (l + 1)) + 1u) << (l + 1)) - 1u;
                const double u = chunk_uniform(i_done - (unsigned)__popc(i_done) + (unsigned)l);
                const XF sw = xf_add(right ? pw : nw, t2w);         // log_size = logaddexp(...)            (:400)
                const XF sa = xf_add(right ? pa : na, right ? na : pa);  // log_weighted_accept_sum          (:401-403)
                const bool mine = xf_u_less(u, sw, t2w) == right;   // logbern(tree2.log_size - log_size)   (:404)
                npE = mine ? npE : ppE;
                nplogp = mine ? nplogp : pplogp;
                npidx = mine ? npidx : ppidx;
                nw = sw;
                na = sa;
              }
              cur.w = nw;
              cur.a = na;
              cur.pE = npE;
              cur.plogp = nplogp;
              cur.pslot = kLeafProp;
              pidx = npidx;
              uc = uc0 + (unsigned)Bc - 1u;
              if constexpr (!TAPE) {
                if (Bc >= 2 && uc > blk + 32u) u_lane = u_hi;  // the stream continues in the next block of 32
              }
            }
#else
            // ---- 3. one transposed reduction: lane r sums row r ---------------------------------------------------------------
            {
              // energies: rows s (kinetic) and B + s (log-density sums), s < Bc
              double sum = 0.0;
              if (lane < Bc || (lane >= B && lane < B + Bc)) sum = row_sum(lane);
              const double lp_s = __shfl_sync(FULL, sum, (lane + B) & 31);
              if (lane < Bc) {
                double pre[2] = {0.0, 0.0};
                if constexpr (Target::kPre > 0) {
                  pre[0] = vPre[2 * lane];
                  pre[1] = vPre[2 * lane + 1];
                }
                const double logp = tgt.finish(lp_s, pre);
                const double E = 0.5 * sum - logp;
                double dE = E - E0;                       // nuts.py:352
                if (isnan(dE)) dE = CUDART_INF;           // :353-354
                XF w = xf_zero();
                if (fabs(dE) < a.Emax) w = xf_exp(-dE);   // log_size = -dE (:359)
                vE[lane] = E;
                vLogp[lane] = logp;
                vdE[lane] = dE;
                vWm[lane] = w.m;
                vWe[lane] = w.e;
              }
              // merge dot products: rows 2B .. 2B + 4 Bc - 7
              const int n_dot = Bc >= 2 ? 4 * Bc - 6 : 0;
              for (int r = lane; r < n_dot; r += 32) vDot[r] = row_sum(2 * B + r);
            }
            __syncwarp();
            // U-turn flags of the chunk's merges, one lane per merge: id = (Bc - (Bc >> l)) + k2 for level l
            unsigned turnmask = 0u;
            if (Bc >= 2) {
              bool flag = false;
              if (lane < Bc - 1) {
                int l = 0, idx = lane, cnt = Bc / 2;
                while (idx >= cnt) {
                  idx -= cnt;
                  cnt >>= 1;
                  ++l;
                }
                if (l == 0) {
                  flag = (vDot[2 * idx] <= 0) || (vDot[2 * idx + 1] <= 0);  // :391
                } else {
                  // first row of level l: Bc + 6 * (Bc/4 + .. + Bc/2^l) = Bc + 6 * (Bc/2 - (Bc >> (l + 1))) ... per level sizes
                  int r0 = Bc;
                  for (int ll = 1; ll < l; ++ll) r0 += 6 * (Bc >> (ll + 1));
                  const double* dd = vDot + r0 + 6 * idx;
                  flag = (dd[0] <= 0) || (dd[1] <= 0) || (dd[2] <= 0) || (dd[3] <= 0) || (dd[4] <= 0) || (dd[5] <= 0);  // :391-398
                }
              }
              turnmask = __ballot_sync(FULL, flag);
            }
            LMC_WTICK(6);
            // ---- 4. the chunk's leaves and merges in the reference's post-order (uniform across the warp) -------------------
            for (int s = 0; s < Bc; ++s) {
              ++n_leaves;
              const double dE = vdE[s];
              if (fabs(dE) > fabs(tr.max_dE)) tr.max_dE = dE;  // :356-357
              if (!(fabs(dE) < a.Emax)) {                      // :358 / :370-375
                fail = 1;
                break;
              }
              cur.w = XF{vWm[s], vWe[s]};
              cur.a = (-dE < 0.0) ? xf_sqr(cur.w) : cur.w;     // log_p_accept_weighted = -dE + min(0, -dE)  (:363)
              cur.pE = vE[s];
              cur.plogp = vLogp[s];
              cur.pslot = kLeafProp;
              pidx = s;
              int lvl = 0;
              for (int bits = s; bits & 1; bits >>= 1, ++lvl) {
                const int mid = (Bc - (Bc >> lvl)) + (s >> (lvl + 1));
                const double u = next_uniform();
                const double* ls = lstk + 5 * lvl;
                const XF t1w{ls[0], lstk_i[3 * lvl]}, t1a{ls[1], lstk_i[3 * lvl + 1]};
                const XF nw = xf_add(t1w, cur.w);   // log_size = logaddexp(...)            (:400)
                const XF na = xf_add(t1a, cur.a);   // log_weighted_accept_sum              (:401-403)
                if (!xf_u_less(u, nw, cur.w)) {     // logbern(tree2.log_size - log_size)   (:404): keep tree1's proposal
                  cur.pE = ls[2];
                  cur.plogp = ls[3];
                  pidx = lstk_i[3 * lvl + 2];
                }
                cur.w = nw;
                cur.a = na;
                if ((turnmask >> mid) & 1u) {
                  fail = 2;
                  break;
                }
              }
              if (fail) break;
              if (s + 1 < Bc) {  // push on the chunk-local stack (scalars only: the vectors are the ring)
                __syncwarp();
                if (lane == 0) {
                  double* ls = lstk + 5 * lvl;
                  ls[0] = cur.w.m;
                  ls[1] = cur.a.m;
                  ls[2] = cur.pE;
                  ls[3] = cur.plogp;
                  lstk_i[3 * lvl] = cur.w.e;
                  lstk_i[3 * lvl + 1] = cur.a.e;
                  lstk_i[3 * lvl + 2] = pidx;
                }
                __syncwarp();
              }
            }
#endif
            if (fail) break;
            LMC_WTICK(7);
            // ---- 5. the chunk is a subtree of level bc: merge it with the stack levels >= bc (generic path) -------------------
            if (n_chunks > 1) {
              double2 var[NP], cur_lp[NP], cur_ps[NP];
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                var[k] = s_var[k * 32];
                cur_lp[k] = ring_p[k * 32];   // left edge = first leaf of the chunk
                cur_ps[k] = psbuf[k * 32];    // p_sum of the whole chunk (Bc >= 2 here)
              }
              int lvl = bc;
              for (unsigned cb = c; cb & 1u; cb >>= 1, ++lvl) {
                if (merge_upper<G, NP>(sc, tgrp, ss, lvl, PairArray<NP>{var}, PairArray<NP>{p}, PairArray<NP>{cur_lp},
                                       PairArray<NP>{cur_ps}, PairArrayOut<NP>{cur_ps}, PairArrayOut<NP>{cur_lp}, cur,
                                       free_slots, next_uniform)) {
                  fail = 2;
                  break;
                }
              }
              if (fail) break;
              if (c + 1 < n_chunks) {
                __builtin_assume(lvl >= 1);
                if (cur.pslot == kLeafProp) {  // the proposal leaves the position ring before the next chunk overwrites it
                  cur.pslot = __ffs(free_slots) - 1;
                  free_slots &= ~(1u << cur.pslot);
#pragma unroll
                  for (int k = 0; k < NP; ++k)
                    sc.st(vid_prop(cur.pslot), k, ring_q[(size_t)(pidx * NP + k) * 32]);
                }
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  sc.st(vid_stack(lvl, 0), k, cur_lp[k]);
                  sc.st(vid_stack(lvl, 1), k, p[k]);
                  sc.st(vid_stack(lvl, 2), k, cur_ps[k]);
                }
                __syncwarp();
                if (lane == 0) {
                  ss->wm[lvl] = cur.w.m;
                  ss->we[lvl] = cur.w.e;
                  ss->am[lvl] = cur.a.m;
                  ss->ae[lvl] = cur.a.e;
                  ss->pE[lvl] = cur.pE;
                  ss->plogp[lvl] = cur.plogp;
                  ss->pslot[lvl] = cur.pslot;
                }
                __syncwarp();
              } else {
                // last chunk: its merged result is the doubling's subtree; keep its vectors for extend_top
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  psbuf[k * 32] = cur_ps[k];
                  ring_p[k * 32] = cur_lp[k];
                }
              }
            }
            LMC_WTICK(8);
          }
          ++tr.depth;            // nuts.py:315
          tr.n_prop += n_leaves;  // :316
          if (fail) {            // :318-319 -> :216-217
            diverging = (fail == 1);
            break;
          }
          LMC_WTICK(8);
          // ---- top of _Tree.extend (nuts.py:321-340): T.left.p = ring_p[0] (or p), T.p_sum = psbuf[0] (or p) ------------------
          {
            double2 var[NP], cur_lp[NP], cur_ps[NP], qprop[NP];
            const bool single = n_leaf_total == 1u;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              var[k] = s_var[k * 32];
              cur_lp[k] = single ? p[k] : ring_p[k * 32];
              cur_ps[k] = single ? p[k] : psbuf[k * 32];
              qprop[k] = q[k];
            }
            if (cur.pslot == kLeafProp) {
#pragma unroll
              for (int k = 0; k < NP; ++k) qprop[k] = ring_q[(size_t)(pidx * NP + k) * 32];
            }
            if (extend_top_f<G, NP>(sc, tgrp, tail, dir, PairArray<NP>{var}, PairArray<NP>{qprop}, PairArray<NP>{p},
                                    PairArray<NP>{cur_lp}, PairArray<NP>{cur_ps}, cur, tr, next_uniform()))
              break;  // :340
          }
          LMC_WTICK(9);
          if (d + 1 < max_depth) {
            const int eb = (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              sc.st(tvid(tail, eb + 0), k, q[k]);
              sc.st(tvid(tail, eb + 1), k, p[k]);
              sc.st(tvid(tail, eb + 2), k, g[k]);
            }
            reg_edge = dir;
          } else {
            reached_max = true;
          }
        }
        LMC_WTICK(10);
        const double accept_stat = mean_tree_accept(tr);
#pragma unroll
        for (int k = 0; k < NP; ++k) q[k] = sc.ld(tvid(tail, T_PROPQ), k);  // hmc_step.end.q

        double* const ad = a.adapt + (size_t)chain * LMC_ADAPT_STRIDE;
        const bool kept = sticky && t > 0;  // the scalars of a sticky chain are on chip since its previous transition
        const double* const adr = kept ? keep : ad;
        static_assert(LMC_ADAPT_LOG_STEP < 10 && LMC_ADAPT_WINDOW < 10, "keep[] mirrors the adapt row");
        DualAvg da{kept ? adr[LMC_ADAPT_LOG_STEP] : __ldcg(ad + LMC_ADAPT_LOG_STEP),
                   kept ? adr[LMC_ADAPT_LOG_BAR] : __ldcg(ad + LMC_ADAPT_LOG_BAR),
                   kept ? adr[LMC_ADAPT_HBAR] : __ldcg(ad + LMC_ADAPT_HBAR),
                   kept ? adr[LMC_ADAPT_COUNT] : __ldcg(ad + LMC_ADAPT_COUNT),
                   kept ? adr[LMC_ADAPT_MU] : __ldcg(ad + LMC_ADAPT_MU)};
        WelfordScalars wel{kept ? adr[LMC_ADAPT_W_FG] : __ldcg(ad + LMC_ADAPT_W_FG),
                           kept ? adr[LMC_ADAPT_W_BG] : __ldcg(ad + LMC_ADAPT_W_BG),
                           (long long)(kept ? adr[LMC_ADAPT_NSAMPLES] : __ldcg(ad + LMC_ADAPT_NSAMPLES)),
                           (long long)(kept ? adr[LMC_ADAPT_WINDOW] : __ldcg(ad + LMC_ADAPT_WINDOW))};
        if (adapt_step) dual_average_update(da, accept_stat, a.target_accept, a.gamma, a.k, a.t0);
        if (tune && a.adapt_mass) {
          const size_t off = (size_t)chain * a.ld;
          double2 var[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k) var[k] = s_var[k * 32];
          welford_update<G, NP>(lane, D, ldh, a.mean_fg + off, a.rawvar_fg + off, a.mean_bg + off, a.rawvar_bg + off, q,
                                var, wel, a.window_multiplier);
          store_row<G, NP>(a.var + off, lane, ldh, var);
#pragma unroll
          for (int k = 0; k < NP; ++k) s_var[k * 32] = var[k];  // sticky launches keep the mass matrix on chip
        }
        double* const trow = trace_row();
        double* const srow = stats_row();
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const int j = lane + k * G;
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
          srow[LMC_STAT_N_UNIFORMS] = (double)uc;
          srow[LMC_STAT_REACHED_MAX_TREEDEPTH] = reached_max ? 1.0 : 0.0;
        }
        store_row<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
        __syncwarp();  // every lane has read the adaptation scalars before lane 0 overwrites them
        if (lane == 0) {
          // (after tuning the step sizes do not move: a sticky chain reuses the pair it formed last time)
          const bool same = kept && !adapt_step;
          const double e_step = same ? keep[10] : exp_cold(da.log_step), e_bar = same ? keep[11] : exp_cold(da.log_bar);
          srow[LMC_STAT_STEP_SIZE] = e_step;
          srow[LMC_STAT_STEP_SIZE_BAR] = e_bar;
          if (sticky) {
            keep[LMC_ADAPT_LOG_STEP] = da.log_step;
            keep[LMC_ADAPT_LOG_BAR] = da.log_bar;
            keep[LMC_ADAPT_HBAR] = da.hbar;
            keep[LMC_ADAPT_COUNT] = da.count;
            keep[LMC_ADAPT_MU] = da.mu;
            keep[LMC_ADAPT_W_FG] = wel.w_fg;
            keep[LMC_ADAPT_W_BG] = wel.w_bg;
            keep[LMC_ADAPT_NSAMPLES] = (double)wel.n_samples;
            keep[LMC_ADAPT_WINDOW] = (double)wel.window;
            keep[10] = e_step;
            keep[11] = e_bar;
          }
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
    if (sticky) {
      sticky_dead = dead;
      if (completes_block(a, t)) {  // uniform across the warp
        __threadfence();
        __syncwarp();
        if (lane == 0) report_block(a, t);
      }
      __syncwarp();
    } else {
      __threadfence();
      __syncwarp();
      if (lane == 0 && completes_block(a, t)) report_block(a, t);
      if (lane == 0 && t + 1 < a.n_trans) sched_push(sv, (unsigned)a.n_chains, chain, t + 1, dead);
    }
    LMC_WTICK(11);
  }
#ifdef LMC_WARP_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("warp-timing kcycles: pop %lld load+draw %lld init|edgesave %lld dir+reload %lld leap %lld dots %lld reduce %lld "
           "scalar %lld merge %lld extend %lld exit %lld epilogue+push %lld\n", s_tacc[0] / 1000, s_tacc[1] / 1000,
           s_tacc[2] / 1000, s_tacc[3] / 1000, s_tacc[4] / 1000, s_tacc[5] / 1000, s_tacc[6] / 1000, s_tacc[7] / 1000,
           s_tacc[8] / 1000, s_tacc[9] / 1000, s_tacc[10] / 1000, s_tacc[11] / 1000);
#endif
}

#ifndef __CUDACC_RTC__
// `kern`: a sampler_warp_kernel instantiation of this library or a cudaKernel_t compiled at run time for a user target
// (lmc_user.cu); `tgt`: host pointer to the kernel's by-value target argument.
template <int NP, int B, int WPB>
int launch_warp_kernel(const void* kern, const lmc_sampler_args& a, const void* tgt, int stage_vecs = 0) {
  using LY = WarpLayout<NP, B>;
  constexpr int VS = LY::VS;
  int dev = 0, n_sm = 0, smem_optin = 0;
  LMC_CUDA(cudaGetDevice(&dev));
  LMC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  LMC_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const int sdepth = scratch_depth(a);
  WarpCfg cfg;
  cfg.ws_vecs = ws_vecs_warp(sdepth, B);
  // tree-scratch vectors in shared memory: the three trajectory vectors every doubling touches (ids 0..2) by default
  int n_smem = a.tune_smem_vecs >= 0 ? a.tune_smem_vecs : 3;
  const int hot = vid_tail(sdepth);
  if (n_smem > hot) n_smem = hot;
  cfg.n_smem_vecs = n_smem;
  cfg.slot_bytes = LY::kFixedBytes + (stage_vecs + n_smem) * VS * 16;
  const size_t smem = (size_t)WPB * cfg.slot_bytes;
  if (smem > (size_t)smem_optin) return LMC_ERR_UNSUPPORTED;
  LMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * WPB, smem));
  if (occ < 1) return LMC_ERR_UNSUPPORTED;
  long long blocks_needed = ((long long)a.n_chains + WPB - 1) / WPB;
  long long grid = (long long)n_sm * occ;  // persistent: every CTA resident, a warp waiting on the ring never deadlocks
  if (a.tune_max_slots > 0 && grid * WPB > a.tune_max_slots) grid = (a.tune_max_slots + WPB - 1) / WPB;
  if (grid > blocks_needed) grid = blocks_needed;
  if (grid < 1) grid = 1;
  const long long need = (long long)sched_bytes(a.n_chains) + grid * WPB * (long long)cfg.ws_vecs * VS * 16;
  if (need > a.workspace_bytes) return LMC_ERR_WORKSPACE;
  if ((long long)a.n_chains * a.n_trans >= (1ll << 31)) return LMC_ERR_UNSUPPORTED;
  sched_init_kernel<<<(a.n_chains + 255) / 256, 256, 0, (cudaStream_t)a.stream>>>(a.workspace, a.n_chains, a.n_trans);
  void* kargs[] = {const_cast<lmc_sampler_args*>(&a), const_cast<void*>(tgt), &cfg};
  LMC_CUDA(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(32 * WPB), kargs, smem, (cudaStream_t)a.stream));
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

template <class Target, int NP, int B, int WPB, int MINB, bool TAPE>
int launch_warp_mode(const lmc_sampler_args& a, const Target& tgt) {
  return launch_warp_kernel<NP, B, WPB>(reinterpret_cast<const void*>(sampler_warp_kernel<Target, NP, B, WPB, MINB, TAPE>),
                                        a, &tgt, warp_stage_vecs<Target, NP, B>());
}

template <class Target, int NP, int B, int WPB, int MINB>
int launch_warp(const lmc_sampler_args& a, const Target& tgt) {
  return a.rng.mode == LMC_RNG_TAPE ? launch_warp_mode<Target, NP, B, WPB, MINB, true>(a, tgt)
                                    : launch_warp_mode<Target, NP, B, WPB, MINB, false>(a, tgt);
}
#endif  // !__CUDACC_RTC__

}  // namespace lmc
