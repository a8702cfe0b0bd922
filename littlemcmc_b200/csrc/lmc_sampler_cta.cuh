// Chunked CTA-per-chain NUTS kernel: the sampler for 257 .. 1024 dimensions (BASELINE configs 3, 5 and the headline).
//
// The chunking of lmc_sampler_warp.cuh carried over to one CTA of G = 128 threads per chain.  The lean kernel
// (lmc_sampler_lean.cuh) pays, per leaf, a block-wide energy reduction (butterfly + shared-memory exchange + barrier), an
// exp and the scalar bookkeeping on the critical path, plus half a level-0 and a quarter of a level-1 merge reduction --
// 870 instructions per warp and leapfrog of which ~100 are the leapfrog (profiles/r02z_headline_ncu_full.md: issue slots
// 38%, the rest is waiting on those dependent chains).  Inside one subtree the leapfrog trajectory does not depend on any
// tree decision, so here the tree is built in chunks of B = 2^b consecutive leaves (B = 4 by default):
//   1. B leapfrogs back to back, state in registers, no reduction and NO BARRIER in between (a target whose gradient
//      needs a sum keeps that one); every leaf's momentum goes to a shared-memory ring, its position to an L2-resident
//      ring, the per-thread partials of its kinetic energy / log density to a shared-memory table;
//   2. the dot products of the B - 1 merges inside the chunk (levels 0 .. b-1 of reference nuts.py:387-398) in ONE pass
//      over the ring: every thread on its own columns, the velocities var*p formed once, partials into the same table;
//   3. one transposed reduction of all 6B - 6 rows: after a barrier thread (warp w, lane r) sums columns 32w .. 32w+31 of
//      row r (16 LDS.128, no shuffles); after a second barrier EVERY warp finishes the rows for itself (4 quarter sums),
//      B lanes take the energies' exp in parallel, the U-turn flags of all merges come out of one ballot;
//   4. every warp walks the chunk's leaves and merges in the reference's post-order on its own copy of the scalars
//      (identical inputs, identical results: no barrier, same uniforms in the same order as the recursion);
//   5. the chunk is a subtree of level b and merges with the stack levels >= b through lmc_tree.cuh's merge_upper.
// Two barriers per chunk instead of ~1.75 per leaf, the level-0/1 stack traffic stays in the ring, the scalar pass is off
// the vector critical path.  Everything a thread reads from the ring, the running p_sum and the mass matrix it wrote
// itself (thread-private columns), so those need no synchronisation at all.
//
// Registers hold q, p, grad only (as in the lean kernel): var, the chunk's p_sum and its left-edge momentum are shared
// memory operands of merge_upper / extend_top_f (accessor functors), updated in place.
// Same arithmetic per element and the same decisions as the reference; dot products are summed in a fixed order that
// differs from the other kernels' (like BLAS ddot, unspecified): results agree with the oracle to the same ~1e-13.
#pragma once
#include "lmc_sampler.cuh"

namespace lmc {

template <int G, int NP, int B>
struct CtaLayout {
  static_assert(G == 128, "four warps per chain: the quarter-row reduction assumes it");
  static_assert(B == 2 || B == 4, "chunk of 2 or 4 leaves");
  static constexpr int VS = G * NP;                  // pairs per vector
  static constexpr int kWarps = G / 32;
  static constexpr int kLog = (B == 2 ? 1 : 2);
  static constexpr int kRows = 6 * B - 6;            // table rows: 2B energy partials + (4B - 6) merge dot products
  static constexpr int kDots = 4 * B - 6;
  static constexpr int kRowLd = G + 2;               // doubles per table row (conflict-free 128-bit reads down a column block)
  static_assert(kRows <= 32, "one lane per table row");
  static constexpr int a16(int x) { return (x + 15) & ~15; }
  // byte offsets
  static constexpr int oRingP = 0;                                  // [B][NP][G] double2: momenta of the chunk's leaves
  static constexpr int oPs = oRingP + B * VS * 16;                  // [NP][G] p_sum of the subtree being assembled
  static constexpr int oVar = oPs + VS * 16;                        // [NP][G] mass-matrix diagonal
  static constexpr int oPart = oVar + VS * 16;                      // [kRows][kRowLd] per-thread partials
  static constexpr int oQsum = a16(oPart + kRows * kRowLd * 8);     // [kRows][kWarps] quarter-row sums
  static constexpr int oPre = oQsum + kRows * kWarps * 8;           // [B][2] reduced pre-sums of the leaves
  static constexpr int oRed = oPre + 2 * B * 8;                     // Group<G>: [2][kWarps][kRedSlots]
  static constexpr int oPriv = oRed + 2 * kWarps * kRedSlots * 8;   // per-warp copies of the chunk scalars
  // one warp's copy: doubles E[B], logp[B], dE[B], wm[B], dot[kDots], local stack (kLog + 1) x 5; ints we[B], (kLog + 1) x 3
  static constexpr int nPrivD = 4 * B + kDots + 5 * (kLog + 1);
  static constexpr int nPrivI = B + 3 * (kLog + 1);
  static constexpr int kPrivBytes = a16(nPrivD * 8 + nPrivI * 4);
  static constexpr int oSS = oPriv + kWarps * kPrivBytes;
  static constexpr int oPop = oSS + (int)sizeof(StackScalars);
  static constexpr int kFixedBytes = a16(oPop + 8);                 // tree-scratch vectors kept on chip follow
};

// All-reduce of up to 8 values over the CTA through the partial-sum table instead of shuffle butterflies: thread c stores
// value n into column c of row n, thread (warp w, lane r < N) sums columns 32w .. 32w+31 of row r, everybody adds the four
// quarter sums.  Two barriers like the butterfly + exchange it replaces, but ~60 instructions per call site instead of
// ~230: the five call sites of this kernel were 18 KB of SHFL code, and the kernel's footprint is what bounds it
// (profiles/r02y_cta2_ncu_full.md: 12% of the warp samples wait for instructions once the code passes ~100 KB).
template <int LD>
struct TableGroupCta {
  int lane;
  double* rows;  // [8][LD] partials, idle rows of the table
  double* qs;    // [8][4] quarter sums
  template <int N>
  __device__ __forceinline__ void allreduce(double (&v)[N]) {
    static_assert(N <= 8, "TableGroupCta reduces up to 8 values");
#pragma unroll
    for (int n = 0; n < N; ++n) rows[n * LD + lane] = v[n];
    __syncthreads();
    const int wl = lane & 31, wid = lane >> 5;
    if (wl < N) {
      const double2* r = reinterpret_cast<const double2*>(rows + wl * LD + 32 * wid);
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const double2 x = r[k], y = r[k + 1];
        s0 += x.x;
        s1 += x.y;
        s2 += y.x;
        s3 += y.y;
      }
      qs[wl * 4 + wid] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const double2* t = reinterpret_cast<const double2*>(qs + n * 4);
      const double2 x = t[0], y = t[1];
      v[n] = (x.x + x.y) + (y.x + y.y);
    }
  }
};

// LMC_CTA_TIMING: thread 0 of every CTA accumulates clock64() deltas per phase (index = the phase that just ENDED) and
// block 0 prints its totals -- a development probe (tools/quick_bench.py under a -DLMC_CTA_TIMING=1 variant build)
#ifdef LMC_CTA_TIMING
#include <stdio.h>
#define LMC_TICK(i)                                  \
  if (threadIdx.x == 0) {                            \
    const long long now_ = clock64();                \
    s_tacc[i] += now_ - s_tacc[15];                  \
    s_tacc[15] = now_;                               \
  }
#else
#define LMC_TICK(i)
#endif

template <class Target, int G, int NP, int B, int MINB, bool TAPE>
__global__ void __launch_bounds__(G, MINB) sampler_cta_kernel(const lmc_sampler_args a, const Target tgt,
                                                              const KernelCfg cfg) {
  using LY = CtaLayout<G, NP, B>;
  constexpr int VS = LY::VS;
  constexpr int KW = LY::kWarps;
  constexpr int LD = LY::kRowLd;
  constexpr unsigned FULL = 0xffffffffu;
  // scratch words of several pairs loaded before anything is stored (lmc_tree.cuh: BATCH): only with the registers of
  // three resident CTAs per SM
  constexpr bool kBatch = MINB < 4;
  extern __shared__ double2 smem2[];
  const int lane = threadIdx.x;
  const int wl = lane & 31;
  const int wid = lane >> 5;
  const int slot = blockIdx.x;
  char* const base = reinterpret_cast<char*>(smem2);
  double2* const ring_p = reinterpret_cast<double2*>(base + LY::oRingP) + lane;  // [B][NP][G]
  double2* const psbuf = reinterpret_cast<double2*>(base + LY::oPs) + lane;      // [NP][G]
  double2* const s_var = reinterpret_cast<double2*>(base + LY::oVar) + lane;     // [NP][G]
  double* const part = reinterpret_cast<double*>(base + LY::oPart);              // [kRows][LD]: row = value, column = thread
  double* const qsum = reinterpret_cast<double*>(base + LY::oQsum);              // [kRows][KW]
  double* const vPre = reinterpret_cast<double*>(base + LY::oPre);               // [B][2]
  double* const vE = reinterpret_cast<double*>(base + LY::oPriv + wid * LY::kPrivBytes);  // this warp's copy
  double* const vLogp = vE + B;
  double* const vdE = vLogp + B;
  double* const vWm = vdE + B;
  double* const vDot = vWm + B;            // [kDots]
  double* const lstk = vDot + LY::kDots;   // local stack: [kLog + 1][5] = wm, am, pE, plogp, (unused)
  int* const vWe = reinterpret_cast<int*>(lstk + 5 * (LY::kLog + 1));  // [B]
  int* const lstk_i = vWe + B;                                         // [kLog + 1][3] = we, ae, pidx
  StackScalars* const ss = reinterpret_cast<StackScalars*>(base + LY::oSS);
  int* const s_pop = reinterpret_cast<int*>(base + LY::oPop);
  Scratch<G, NP> sc;
  sc.sm = reinterpret_cast<double2*>(base + LY::kFixedBytes);
  sc.ws = reinterpret_cast<double2*>(reinterpret_cast<char*>(a.workspace) + sched_bytes(a.n_chains)) +
          (size_t)slot * cfg.ws_vecs * VS;
  sc.n_smem = cfg.n_smem_vecs;
  sc.lane = lane;
  // every all-reduce goes through the table, on rows that are idle at that point:  grp  two values inside a leapfrog (a
  // target's pre-sums; phase 1 does not touch the dot-product rows) and the initial energy;  tgrp  the six values of the
  // stack merges and of extend (after the chunk's table has been consumed): the first rows.  Quarter sums in `qsum`.
  TableGroupCta<LD> grp{lane, part + 2 * B * LD, qsum};
  TableGroupCta<LD> tgrp{lane, part, qsum};
  const SchedView sv = sched_view(a.workspace, a.n_chains, a.n_trans);
  const unsigned total_units = (unsigned)a.n_chains * (unsigned)a.n_trans;

  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  const int sdepth = scratch_depth(a);
  const int tail = vid_tail(sdepth);
  // position ring of the current chunk: B vectors after the tree scratch of this slot (global, written once per leaf,
  // read back only for the one position per chunk that survives as a proposal)
  double2* const ring_q = sc.ws + (size_t)ws_vecs_nuts(sdepth) * VS + lane;
  // operands of the stack merges and of extend, where this kernel keeps them
  const auto var_f = [=](int k) { return s_var[k * G]; };
  const auto lp_f = [=](int k) { return ring_p[k * G]; };   // left edge of the subtree being assembled: ring entry 0
  const auto ps_f = [=](int k) { return psbuf[k * G]; };
  const auto set_ps = [=](int k, double2 v) { psbuf[k * G] = v; };
  const auto set_lp = [=](int k, double2 v) { ring_p[k * G] = v; };
  auto skew = [&](int row) -> double* { return part + row * LD + lane; };
  // total of table row `row` from its four quarter sums (fixed order)
  auto row_total = [&](int row) -> double {
    static_assert(KW == 4, "two 128-bit loads per row");
    const double2* t = reinterpret_cast<const double2*>(qsum + row * KW);
    const double2 x = t[0], y = t[1];
    return (x.x + x.y) + (y.x + y.y);
  };

#ifdef LMC_CTA_TIMING
  __shared__ long long s_tacc[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 15; ++i) s_tacc[i] = 0;
    s_tacc[15] = clock64();
  }
#endif
  // "sticky" launches (lmc_sampler_warp.cuh): with a resident CTA per chain there is nothing to schedule
  const bool sticky = (unsigned)a.n_chains <= gridDim.x;
  int t_next = 0;
  bool sticky_dead = false;
  double2 q[NP];

  for (;;) {
    // ---- the next (chain, transition) unit: own chain (sticky) or popped from the FIFO (lmc_sampler.cuh: scheduler) ----
    int chain = -1, t = 0;
    if (sticky) {
      if (slot < a.n_chains && t_next < a.n_trans) {
        chain = slot | (sticky_dead ? (int)kDeadBit : 0);
        t = t_next++;
      }
    } else {
      if (lane == 0) {
        sched_pop(sv, total_units, (unsigned)a.n_chains, chain, t);
        s_pop[0] = chain;
        s_pop[1] = t;
      }
      __syncthreads();
      chain = s_pop[0];
      t = s_pop[1];
    }
    if (chain == -1) break;
    LMC_TICK(0);
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
        for (int k = 0; k < NP; ++k) s_var[k * G] = var[k];
      }
      const uint64_t seed = TAPE ? 0ull : a.rng.seeds[chain];
      const long long it = a.iter0 + t;
      const bool tune = it < a.n_tune;
      const bool adapt_step = tune && a.adapt_step_size;
      unsigned uc = 0;
      double u_lane = 0.0;
      auto next_uniform = [&]() -> double {  // every warp keeps the stream for itself (same values, no exchange)
        double u;
        if constexpr (TAPE) {
          if ((long long)uc < a.rng.u_stride) {
            u = a.rng.uniforms[row * a.rng.u_stride + uc];
          } else {
            u = 0.5;
            status |= LMC_STATUS_TAPE_EXHAUSTED;
          }
        } else {
          if ((uc & 31u) == 0u) u_lane = philox_uniform_cold(seed, it, uc + (unsigned)wl);
          u = __shfl_sync(FULL, u_lane, (int)(uc & 31u));
        }
        ++uc;
        return u;
      };

      // ---- p0 = potential.random()  (quadpotential.py:221-224 / 374-376) ------------------------------------------
      {
        double2 nrm[NP], vv[NP], is[NP];
        if constexpr (TAPE) {
          const double* normals_row = a.rng.normals + row * D;
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            const int j = lane + k * G;
            nrm[k] = make_double2(2 * j < D ? normals_row[2 * j] : 0.0, 2 * j + 1 < D ? normals_row[2 * j + 1] : 0.0);
          }
        } else {
          // (one out-of-line copy called NP times: an unrolled, interleaved version is 15% faster for a chain running
          //  alone and 5% slower with three resident chains per SM -- 15 KB of straight-line code per transition)
#pragma unroll
          for (int k = 0; k < NP; ++k) nrm[k] = philox_normal_pair(seed, it, (uint32_t)(lane + k * G));
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) vv[k] = s_var[k * G];
        inv_sqrt_pairs<NP>(vv, is);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const int j = lane + k * G;
          p[k].x = (2 * j < D) ? mul_rn(is[k].x, nrm[k].x) : 0.0;
          p[k].y = (2 * j + 1 < D) ? mul_rn(is[k].y, nrm[k].y) : 0.0;
        }
      }

      LMC_TICK(1);
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
        for (int k = 0; k < NP; ++k) acc[0] = dot2(acc[0], p[k], mul2(s_var[k * G], p[k]));
        grp.allreduce(acc);
        logp0 = tgt.finish(acc[1], pre);
        E0 = 0.5 * acc[0] - logp0;
      }
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
          LMC_TICK(2);
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
            LMC_TICK(3);
            // ---- 1. Bc leapfrogs (integration.py:100-121), no reduction between them ------------------------------------
            for (int s = 0; s < Bc; ++s) {
              double pre[2] = {0.0, 0.0};
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                p[k] = axpy2(p[k], dt, g[k]);
                q[k] = axpy2(q[k], eps_d, mul2(s_var[k * G], p[k]));
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
                k_part = dot2(k_part, p[k], mul2(s_var[k * G], p[k]));
                ring_p[(s * NP + k) * G] = p[k];
                ring_q[(size_t)(s * NP + k) * G] = q[k];
              }
              *skew(s) = k_part;
              *skew(B + s) = lp_part;
            }
            LMC_TICK(4);
            // ---- 2. dot products of the merges inside the chunk (nuts.py:387-398), every thread on its own columns ----------
            if (Bc == 2) {
              double d0 = 0.0, d1 = 0.0;
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                const double2 pa = ring_p[k * G], pb = p[k];  // the chunk's last leaf is still in registers
                const double2 vk = s_var[k * G];
                const double2 ps = add2(pa, pb);          // p_sum = tree1.p_sum + tree2.p_sum (:390)
                d0 = dot2(d0, ps, mul2(vk, pa));          // p_sum . left.v
                d1 = dot2(d1, ps, mul2(vk, pb));          // p_sum . right.v
                psbuf[k * G] = ps;
              }
              *skew(2 * B) = d0;
              *skew(2 * B + 1) = d1;
            }
            if constexpr (B == 4) {
              if (Bc == 4) {
                // leaves 0..3; level 0: (0,1) and (2,3), single leaves: left.p == right.p == p_sum; level 1: tree1 = (0,1)
                // with left.p = p0, right.p = p1, tree2 = (2,3) with left.p = p2, right.p = p3.  var*p formed once per leaf.
                double e[10] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  const double2 p0 = ring_p[k * G], p1 = ring_p[(NP + k) * G], p2 = ring_p[(2 * NP + k) * G], p3 = p[k];
                  const double2 vk = s_var[k * G];
                  const double2 v0 = mul2(vk, p0), v1 = mul2(vk, p1), v2 = mul2(vk, p2), v3 = mul2(vk, p3);
                  const double2 s01 = add2(p0, p1), s23 = add2(p2, p3);  // :390 at level 0
                  e[0] = dot2(e[0], s01, v0);
                  e[1] = dot2(e[1], s01, v1);
                  e[2] = dot2(e[2], s23, v2);
                  e[3] = dot2(e[3], s23, v3);
                  const double2 ps = add2(s01, s23);   // :390
                  const double2 ps1 = add2(s01, p2);   // tree1.p_sum + tree2.left.p (:394)
                  const double2 ps2 = add2(p1, s23);   // tree1.right.p + tree2.p_sum (:396)
                  e[4] = dot2(e[4], ps, v0);
                  e[5] = dot2(e[5], ps, v3);
                  e[6] = dot2(e[6], ps1, v0);
                  e[7] = dot2(e[7], ps1, v2);
                  e[8] = dot2(e[8], ps2, v1);
                  e[9] = dot2(e[9], ps2, v3);
                  psbuf[k * G] = ps;
                }
#pragma unroll
                for (int i = 0; i < 10; ++i) *skew(2 * B + i) = e[i];
              }
            }
            LMC_TICK(5);
            __syncthreads();
            // ---- 3. one transposed reduction: thread (warp w, lane r) sums columns 32w .. 32w+31 of row r ---------------------
            const int n_dot = Bc >= 2 ? 4 * Bc - 6 : 0;
            if (wl < Bc || (wl >= B && wl < B + Bc) || (wl >= 2 * B && wl < 2 * B + n_dot)) {
              const double2* r = reinterpret_cast<const double2*>(part + wl * LD + 32 * wid);
              double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
              for (int k = 0; k < 16; k += 2) {
                const double2 x = r[k], y = r[k + 1];
                s0 += x.x;
                s1 += x.y;
                s2 += y.x;
                s3 += y.y;
              }
              qsum[wl * KW + wid] = (s0 + s1) + (s2 + s3);
            }
            __syncthreads();
            // every warp finishes the rows for itself: energies -> exp() on Bc lanes, dot products -> U-turn flags
            if (wl < Bc) {
              double pre[2] = {0.0, 0.0};
              if constexpr (Target::kPre > 0) {
                pre[0] = vPre[2 * wl];
                pre[1] = vPre[2 * wl + 1];
              }
              const double logp = tgt.finish(row_total(B + wl), pre);
              const double E = 0.5 * row_total(wl) - logp;
              double dE = E - E0;                       // nuts.py:352
              if (isnan(dE)) dE = CUDART_INF;           // :353-354
              XF w = xf_zero();
              if (fabs(dE) < a.Emax) w = xf_exp(-dE);   // log_size = -dE (:359)
              vE[wl] = E;
              vLogp[wl] = logp;
              vdE[wl] = dE;
              vWm[wl] = w.m;
              vWe[wl] = w.e;
            }
            if (wl < n_dot) vDot[wl] = row_total(2 * B + wl);
            __syncwarp();
            // U-turn flags of the chunk's merges, one lane per merge: id = (Bc - (Bc >> l)) + k2 for level l
            unsigned turnmask = 0u;
            if (Bc >= 2) {
              bool flag = false;
              if (wl < Bc - 1) {
                if (wl < Bc / 2) {
                  flag = (vDot[2 * wl] <= 0) || (vDot[2 * wl + 1] <= 0);  // :391
                } else {
                  const double* dd = vDot + Bc;  // the one level-1 merge of a chunk of four
                  flag = (dd[0] <= 0) || (dd[1] <= 0) || (dd[2] <= 0) || (dd[3] <= 0) || (dd[4] <= 0) || (dd[5] <= 0);  // :391-398
                }
              }
              turnmask = __ballot_sync(FULL, flag);
            }
            LMC_TICK(6);
            // ---- 4. the chunk's leaves and merges in the reference's post-order (every warp on its own copy) -----------------
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
                if (wl == 0) {
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
            if (fail) break;
            LMC_TICK(7);
            // ---- 5. the chunk is a subtree of level bc: merge it with the stack levels >= bc (lmc_tree.cuh) -------------------
            if (n_chunks > 1) {
              int lvl = bc;
              for (unsigned cb = c; cb & 1u; cb >>= 1, ++lvl) {
                if (merge_upper<G, NP, kBatch>(sc, tgrp, ss, lvl, var_f, PairArray<NP>{p}, lp_f, ps_f, set_ps, set_lp, cur,
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
                  double2 t[NP];  // loads first: a scratch store may alias the next load, one round trip instead of NP
#pragma unroll
                  for (int k = 0; k < NP; ++k) t[k] = ring_q[(size_t)(pidx * NP + k) * G];
#pragma unroll
                  for (int k = 0; k < NP; ++k) sc.st(vid_prop(cur.pslot), k, t[k]);
                }
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  sc.st(vid_stack(lvl, 0), k, ring_p[k * G]);
                  sc.st(vid_stack(lvl, 1), k, p[k]);
                  sc.st(vid_stack(lvl, 2), k, psbuf[k * G]);
                }
                // (no barrier: the next reader of these scalars is a merge_upper behind its reduction's barrier)
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
            LMC_TICK(8);
          }
          ++tr.depth;            // nuts.py:315
          tr.n_prop += n_leaves;  // :316
          if (fail) {            // :318-319 -> :216-217
            diverging = (fail == 1);
            break;
          }
          LMC_TICK(8);
          // ---- top of _Tree.extend (nuts.py:321-340): T.left.p = ring entry 0, T.p_sum = psbuf (or p for a single leaf) ------
          {
            const bool single = n_leaf_total == 1u;
            const int pi = pidx;
            if (extend_top_f<G, NP, kBatch>(sc, tgrp, tail, dir, var_f,
                                    [&](int k) { return ring_q[(size_t)(pi * NP + k) * G]; },  // read when the proposal is a leaf
                                    PairArray<NP>{p}, lp_f, [&](int k) { return single ? p[k] : psbuf[k * G]; }, cur, tr,
                                    next_uniform()))
              break;  // :340
          }
          LMC_TICK(9);
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
        LMC_TICK(10);
        const double accept_stat = mean_tree_accept(tr);
#pragma unroll
        for (int k = 0; k < NP; ++k) q[k] = sc.ld(tvid(tail, T_PROPQ), k);  // hmc_step.end.q

        double* const ad = a.adapt + (size_t)chain * LMC_ADAPT_STRIDE;
        DualAvg da{__ldcg(ad + LMC_ADAPT_LOG_STEP), __ldcg(ad + LMC_ADAPT_LOG_BAR), __ldcg(ad + LMC_ADAPT_HBAR),
                   __ldcg(ad + LMC_ADAPT_COUNT), __ldcg(ad + LMC_ADAPT_MU)};
        WelfordScalars wel{__ldcg(ad + LMC_ADAPT_W_FG), __ldcg(ad + LMC_ADAPT_W_BG),
                           (long long)__ldcg(ad + LMC_ADAPT_NSAMPLES), (long long)__ldcg(ad + LMC_ADAPT_WINDOW)};
        // only thread 0 uses the step-size state (statistics row, write-back): the other warps skip its ~300 instructions
        if (adapt_step && wid == 0) dual_average_update(da, accept_stat, a.target_accept, a.gamma, a.k, a.t0);
        if (tune && a.adapt_mass) {
          const size_t off = (size_t)chain * a.ld;
          double2 var[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k) var[k] = s_var[k * G];
          if (kBatch)
            welford_update_batched<G, NP>(lane, D, ldh, a.mean_fg + off, a.rawvar_fg + off, a.mean_bg + off,
                                          a.rawvar_bg + off, q, var, wel, a.window_multiplier);
          else
            welford_update<G, NP>(lane, D, ldh, a.mean_fg + off, a.rawvar_fg + off, a.mean_bg + off, a.rawvar_bg + off, q,
                                  var, wel, a.window_multiplier);
          store_row<G, NP>(a.var + off, lane, ldh, var);
#pragma unroll
          for (int k = 0; k < NP; ++k) s_var[k * G] = var[k];  // sticky launches keep the mass matrix on chip
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
    if (sticky) {
      sticky_dead = dead;
      if (completes_block(a, t)) {  // uniform across the CTA
        LMC_TICK(11);
        __threadfence();
        __syncthreads();
        if (lane == 0) report_block(a, t);
      }
      __syncthreads();  // lane 0's adaptation scalars are in place before the next transition reads them
    } else {
      LMC_TICK(11);
      __threadfence();
      __syncthreads();
      if (lane == 0 && completes_block(a, t)) report_block(a, t);
      if (lane == 0 && t + 1 < a.n_trans) sched_push(sv, (unsigned)a.n_chains, chain, t + 1, dead);
    }
    LMC_TICK(12);
  }
#ifdef LMC_CTA_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 7)
    printf("cta-timing kcycles: pop %lld load+draw %lld init|edgesave %lld dir+reload %lld leap %lld dots %lld reduce %lld "
           "scalar %lld merge %lld extend %lld exit %lld epilogue %lld fence+push %lld\n", s_tacc[0] / 1000, s_tacc[1] / 1000,
           s_tacc[2] / 1000, s_tacc[3] / 1000, s_tacc[4] / 1000, s_tacc[5] / 1000, s_tacc[6] / 1000, s_tacc[7] / 1000,
           s_tacc[8] / 1000, s_tacc[9] / 1000, s_tacc[10] / 1000, s_tacc[11] / 1000, s_tacc[12] / 1000);
#endif
}

#ifndef __CUDACC_RTC__
// vectors of global scratch per slot: tree stack + trajectory (as the other kernels) + the position ring of one chunk
__host__ __device__ constexpr int ws_vecs_cta(int sdepth, int chunk) { return ws_vecs_nuts(sdepth) + chunk; }

// `kern`: a sampler_cta_kernel instantiation; `tgt`: host pointer to the kernel's by-value target argument.
template <int G, int NP, int B, int MINB>
int launch_cta_kernel(const void* kern, const lmc_sampler_args& a, const void* tgt) {
  using LY = CtaLayout<G, NP, B>;
  constexpr int VS = LY::VS;
  int dev = 0, n_sm = 0, smem_optin = 0, smem_sm = 0;
  LMC_CUDA(cudaGetDevice(&dev));
  LMC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  LMC_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  LMC_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  const int sdepth = scratch_depth(a);
  KernelCfg cfg;
  cfg.ws_vecs = ws_vecs_cta(sdepth, B);
  const size_t vec_bytes = (size_t)VS * sizeof(double2);
  const size_t fixed = LY::kFixedBytes;
  // tree-scratch vectors kept on chip (hottest first: the three trajectory vectors every doubling reads): as many as
  // fit beside MINB resident CTAs per SM (1 KB per CTA is the system's)
  const size_t per_cta = (size_t)smem_sm / MINB - 1024;
  const size_t cap = per_cta < (size_t)smem_optin ? per_cta : (size_t)smem_optin;
  if (fixed > cap) return LMC_ERR_UNSUPPORTED;
  int n_smem = (int)((cap - fixed) / vec_bytes);
  const int hot = vid_tail(sdepth);
  if (n_smem > hot) n_smem = hot;
  if (a.tune_smem_vecs >= 0) n_smem = a.tune_smem_vecs < hot ? a.tune_smem_vecs : hot;
  cfg.n_smem_vecs = n_smem;
  const size_t smem = fixed + (size_t)n_smem * vec_bytes;
  if (smem > (size_t)smem_optin) return LMC_ERR_UNSUPPORTED;
  LMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  LMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G, smem));
  if (occ < 1) return LMC_ERR_UNSUPPORTED;
  long long grid = (long long)n_sm * occ;  // persistent: every CTA resident, a CTA waiting on the ring never deadlocks
  if (a.tune_max_slots > 0 && grid > a.tune_max_slots) grid = a.tune_max_slots;
  if (grid > a.n_chains) grid = a.n_chains;
  if (grid < 1) grid = 1;
  const long long need = (long long)sched_bytes(a.n_chains) + grid * (long long)cfg.ws_vecs * (long long)vec_bytes;
  if (need > a.workspace_bytes) return LMC_ERR_WORKSPACE;
  if ((long long)a.n_chains * a.n_trans >= (1ll << 31)) return LMC_ERR_UNSUPPORTED;
  sched_init_kernel<<<(a.n_chains + 255) / 256, 256, 0, (cudaStream_t)a.stream>>>(a.workspace, a.n_chains, a.n_trans);
  void* kargs[] = {const_cast<lmc_sampler_args*>(&a), const_cast<void*>(tgt), &cfg};
  LMC_CUDA(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(G), kargs, smem, (cudaStream_t)a.stream));
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

template <class Target, int G, int NP, int B, int MINB>
int launch_cta(const lmc_sampler_args& a, const Target& tgt) {
  return a.rng.mode == LMC_RNG_TAPE
             ? launch_cta_kernel<G, NP, B, MINB>(
                   reinterpret_cast<const void*>(sampler_cta_kernel<Target, G, NP, B, MINB, true>), a, &tgt)
             : launch_cta_kernel<G, NP, B, MINB>(
                   reinterpret_cast<const void*>(sampler_cta_kernel<Target, G, NP, B, MINB, false>), a, &tgt);
}
#endif  // !__CUDACC_RTC__

}  // namespace lmc
