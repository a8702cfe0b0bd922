// lmc_nuts_sample / lmc_hmc_sample: whole MCMC transitions (momentum draw, initial state, trajectory, both
// adaptations, trace + statistics) for thousands of independent chains in ONE launch.
//
// Execution model.  A chain is owned by a thread group (lmc_device.cuh) for all `n_trans` transitions of the
// call; its live phase-space point (q, p, grad), its mass-matrix diagonal and every tree scalar stay in
// registers for the whole call.  Chains never synchronise with each other, so there is no lock-step loss: a
// chain that needs 1023 leapfrogs for a draw does not hold up one that needs 3.  The grid is persistent
// (`n_slots` resident groups striding over the chains) so the tree scratch is per resident slot, not per chain.
//
// NUTS tree (reference nuts.py:251-435) is built with the iterative binary-counter stack of SURVEY.md A.1:
// leaf i is merged with stack level 0,1,.. for every trailing 1-bit of i, which visits merges in exactly the
// post-order of the reference's recursion and therefore consumes uniforms in the same order.  A stack entry
// keeps only left.p, right.p, p_sum (velocities are recomputed as var*p, bit-identical to the stored ones)
// and an index into a small pool of proposal-position vectors, so choosing a proposal moves an index, never
// a vector.  The hottest scratch vectors (level-0/1 entries and the first proposal slots) live in shared
// memory, the rest in an L2-resident global workspace.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lmc_common.h"
#include "lmc_device.cuh"

namespace lmc {

enum { KIND_NUTS = 0, KIND_HMC = 1 };
constexpr int kLeafProp = -1;  // "the proposal is the current leaf, still in registers"

// scratch-vector ids (ordered hottest first; ids < n_smem_vecs are shared-memory resident)
//   0                 stack level 0: p  (left.p == right.p == p_sum for a single leaf)
//   1, 2              proposal slots 0, 1
//   3+4(l-1)+{0,1,2}  stack level l >= 1: left.p, right.p, p_sum
//   3+4(l-1)+3        proposal slot l+1
//   tail              trajectory edges L(q,p,g), R(q,p,g), trajectory p_sum, trajectory proposal q
__host__ __device__ constexpr int vid_stack(int level, int which) { return level == 0 ? 0 : 3 + 4 * (level - 1) + which; }
__host__ __device__ constexpr int vid_prop(int slot) { return slot < 2 ? 1 + slot : 4 * slot - 2; }
__host__ __device__ constexpr int vid_tail(int max_depth) { return max_depth < 1 ? 3 : 4 * max_depth - 1; }
enum { T_LQ = 0, T_LP, T_LG, T_RQ, T_RP, T_RG, T_PSUM, T_PROPQ, T_COUNT };
__host__ __device__ constexpr int ws_vecs_nuts(int max_depth) { return vid_tail(max_depth) + T_COUNT; }

// per-level scalars of the subtree stack, one copy per chain in shared memory (written by lane 0 only; every read
// is separated from the write by a group barrier / __syncwarp, see the push below)
struct StackScalars {
  double wm[kMaxDepth], am[kMaxDepth];   // mantissas of exp(log_size), exp(log_weighted_accept_sum)
  double pE[kMaxDepth], plogp[kMaxDepth];  // proposal energy / model_logp
  int we[kMaxDepth], ae[kMaxDepth];      // exponents
  int pslot[kMaxDepth];                  // proposal slot
  int pad[kMaxDepth];
};

struct KernelCfg {
  int n_smem_vecs;  // scratch vectors per group kept in shared memory
  int ws_vecs;      // scratch vectors per slot in the global workspace (ids are absolute: smem ids unused there)
};

template <int G>
__host__ __device__ constexpr int block_threads() { return G >= 64 ? G : 128; }

// Instantiated (threads per chain, pairs per thread, min resident CTAs per SM).  The third column caps registers:
// 65536 / (block_threads * min_ctas) per thread.
#define LMC_SHAPES(X) \
  X(32, 1, 4) X(32, 2, 4) X(32, 4, 3) X(64, 4, 4) X(64, 8, 4) X(128, 2, 4) X(128, 4, 3) X(256, 2, 2) X(256, 4, 1) \
  X(512, 2, 1) X(512, 4, 1) X(1024, 4, 1)

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
  const int n_slots = gridDim.x * CPB;
  double2* const sm = smem2 + (size_t)gib * cfg.n_smem_vecs * VS;
  double* const red = reinterpret_cast<double*>(smem2 + (size_t)CPB * cfg.n_smem_vecs * VS) +
                      gib * (2 * Group<G>::kWarps * kRedSlots);
  StackScalars* const ss = reinterpret_cast<StackScalars*>(
      reinterpret_cast<double*>(smem2 + (size_t)CPB * cfg.n_smem_vecs * VS) + CPB * (2 * Group<G>::kWarps * kRedSlots)) + gib;
  double2* const ws = reinterpret_cast<double2*>(a.workspace) + (size_t)slot * cfg.ws_vecs * VS;
  Group<G> grp(lane, red);

  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  const int n_smem = cfg.n_smem_vecs;
  // this thread's word of scratch vector `id`, pair k
  auto vec = [&](int id) -> double2* { return (id < n_smem ? sm : ws) + (size_t)id * VS + lane; };
  const int tail = vid_tail(a.max_treedepth);

  for (int chain = slot; chain < a.n_chains; chain += n_slots) {
    double2 q[NP], p[NP], g[NP], var[NP];
    load_row<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
    load_row<G, NP>(a.var + (size_t)chain * a.ld, lane, ldh, var);
    mask_tail<G, NP>(lane, D, q);
    mask_tail<G, NP>(lane, D, var);

    double* const ad = a.adapt + (size_t)chain * LMC_ADAPT_STRIDE;
    double log_step = ad[LMC_ADAPT_LOG_STEP], log_bar = ad[LMC_ADAPT_LOG_BAR], hbar = ad[LMC_ADAPT_HBAR];
    double da_count = ad[LMC_ADAPT_COUNT];
    const double da_mu = ad[LMC_ADAPT_MU];
    double w_fg = ad[LMC_ADAPT_W_FG], w_bg = ad[LMC_ADAPT_W_BG];
    long long n_samples = (long long)ad[LMC_ADAPT_NSAMPLES];
    long long window = (long long)ad[LMC_ADAPT_WINDOW];
    const uint64_t seed = (a.rng.mode == LMC_RNG_PHILOX) ? a.rng.seeds[chain] : 0ull;
    int status = 0;

    for (int t = 0; t < a.n_trans; ++t) {
      const long long it = a.iter0 + t;  // BaseHMC.iter_count
      const bool tune = it < a.n_tune;
      const bool adapt_step = tune && a.adapt_step_size;  // base_hmc.py:151
      unsigned uc = 0;                                    // uniforms consumed by this transition
      const size_t row = (size_t)chain * a.n_trans + t;
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
          u = philox_uniform(seed, it, uc);
        }
        ++uc;
        return u;
      };

      // ---- p0 = potential.random()  (quadpotential.py:221-224 / 374-376) ---------------------------------------
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int j = lane + k * G;
        double2 n = make_double2(0.0, 0.0);
        if (a.rng.mode == LMC_RNG_TAPE) {
          const double* nr = a.rng.normals + row * D;
          if (2 * j < D) n.x = nr[2 * j];
          if (2 * j + 1 < D) n.y = nr[2 * j + 1];
        } else if (2 * j < D) {
          n = philox_normal_pair(seed, it, (uint32_t)j);
        }
        // inv_stds * vals with inv_stds = 1/sqrt(var): IEEE sqrt and divide, identical to the stored arrays
        p[k].x = (2 * j < D) ? mul_rn(1.0 / sqrt(var[k].x), n.x) : 0.0;
        p[k].y = (2 * j + 1 < D) ? mul_rn(1.0 / sqrt(var[k].y), n.y) : 0.0;
      }

      // ---- start = integrator.compute_state(q0, p0)  (integration.py:52-66) -------------------------------------
      double E0, logp0;
      eval_energy<false>(tgt, grp, D, ldh, q, p, g, var, 0.0, E0, logp0);
      double* const srow = a.stats + row * LMC_NSTATS;
      if (!isfinite(E0)) {  // base_hmc.py:145-148: the reference raises; we flag the chain and stop it
        status |= LMC_STATUS_BAD_INITIAL_ENERGY;
        const double nan = CUDART_NAN;
        for (int tt = t; tt < a.n_trans; ++tt) {
          double* tr = a.trace + (size_t)chain * a.trace_chain_stride + (size_t)tt * a.trace_draw_stride;
          for (int e = lane; e < D; e += G) tr[e] = nan;
          if (lane == 0) {
            double* s2 = a.stats + ((size_t)chain * a.n_trans + tt) * LMC_NSTATS;
            for (int s = 0; s < LMC_NSTATS; ++s) s2[s] = nan;
          }
        }
        break;
      }
      const double eps = exp(adapt_step ? log_step : log_bar);  // step_sizes.py:58-69

      double accept_stat, stat_a, stat_b, stat_energy, stat_energy_error, stat_c, stat_logp;
      bool diverging = false;

      if constexpr (KIND == KIND_NUTS) {
        // ---- NUTS._hamiltonian_step + _Tree  (nuts.py:204-224, 251-435) -----------------------------------------
        const int max_depth = (tune && it < 200) ? a.early_max_treedepth : a.max_treedepth;  // nuts.py:205-208
        // trajectory state (nuts.py:267-282)
        XF Wp = xf_zero();  // exp(log_size) - 1: total weight of the accepted subtrees (the start point has weight 1)
        XF Acc = xf_zero(); // exp(log_weighted_accept_sum)
        double max_dE = 0.0;
        double prop_E = E0, prop_logp = logp0;
        int depth = 0;
        long long n_prop = 0;
        bool turning = false;
        int reg_edge = 0;  // which trajectory edge (q,p,g) currently sits in registers: 0 both (start), +1 R, -1 L
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          vec(tail + T_LQ)[k * G] = q[k];
          vec(tail + T_LP)[k * G] = p[k];
          vec(tail + T_LG)[k * G] = g[k];
          vec(tail + T_RQ)[k * G] = q[k];
          vec(tail + T_RP)[k * G] = p[k];
          vec(tail + T_RG)[k * G] = g[k];
          vec(tail + T_PSUM)[k * G] = p[k];   // p_sum = start.p.copy()
          vec(tail + T_PROPQ)[k * G] = q[k];  // proposal = start
        }
        for (int d = 0; d < max_depth; ++d) {  // nuts.py:212
          // logbern(log 0.5): log(u) < log(0.5) <=> u < 0.5 (log is monotone; the two can only disagree for the single
          // double adjacent to 0.5)                                                                   nuts.py:213
          const int dir = (next_uniform() < 0.5) ? 1 : -1;
          if (reg_edge != 0 && reg_edge != dir) {  // fetch the edge we extend from (nuts.py:297 / 306)
            const int base = tail + (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              q[k] = vec(base + 0)[k * G];
              p[k] = vec(base + 1)[k * G];
              g[k] = vec(base + 2)[k * G];
            }
          }
          const double eps_d = dir > 0 ? eps : -eps;
          unsigned free_slots = 0xffffffffu;  // proposal-slot pool: bit s set = slot s free
          int fail = 0;                       // 1 = diverging, 2 = turning
          long long n_leaves = 0;
          // summary of the subtree being assembled on top of the stack ("cur"); its right edge is always z
          double2 cur_lp[NP], cur_ps[NP];
          XF cur_w = xf_zero(), cur_a = xf_zero();  // exp(log_size), exp(log_weighted_accept_sum) of "cur"
          double cur_pE = 0.0, cur_plogp = 0.0;
          int cur_pslot = kLeafProp;

          const unsigned n_leaf_total = 1u << d;
          for (unsigned i = 0; i < n_leaf_total; ++i) {  // leaves of _build_subtree in integration order
            double E, logp;
            leapfrog(tgt, grp, D, ldh, eps_d, q, p, g, var, E, logp);  // nuts.py:347
            double dE = E - E0;                                        // :352
            if (isnan(dE)) dE = CUDART_INF;                            // :353-354
            if (fabs(dE) > fabs(max_dE)) max_dE = dE;                  // :356-357
            ++n_leaves;
            if (!(fabs(dE) < a.Emax)) {  // :358 / :370-375
              fail = 1;
              break;
            }
            cur_w = xf_exp(-dE);                             // log_size = -dE
            cur_a = (-dE < 0.0) ? xf_sqr(cur_w) : cur_w;     // log_p_accept_weighted = -dE + min(0, -dE)  (:363)
            cur_pE = E;
            cur_plogp = logp;
            cur_pslot = kLeafProp;
#pragma unroll
            for (int k = 0; k < NP; ++k) cur_lp[k] = cur_ps[k] = p[k];

            unsigned jbits = i;
            int lvl = 0;
            while (jbits & 1u) {  // merge with the stack entry of this level (nuts.py:387-417)
              double2 t1_lp[NP], t1_rp[NP], t1_ps[NP];
              if (lvl == 0) {
#pragma unroll
                for (int k = 0; k < NP; ++k) t1_lp[k] = t1_rp[k] = t1_ps[k] = vec(vid_stack(0, 0))[k * G];
              } else {
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  t1_lp[k] = vec(vid_stack(lvl, 0))[k * G];
                  t1_rp[k] = vec(vid_stack(lvl, 1))[k * G];
                  t1_ps[k] = vec(vid_stack(lvl, 2))[k * G];
                }
              }
              double dots[6] = {0.0, 0.0, 1.0, 1.0, 1.0, 1.0};
              if (lvl == 0) {
                double d2[2] = {0.0, 0.0};
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  const double2 ps = add2(t1_ps[k], cur_ps[k]);  // p_sum = tree1.p_sum + tree2.p_sum (:390)
                  d2[0] = dot2(d2[0], ps, mul2(var[k], t1_lp[k]));   // p_sum . left.v
                  d2[1] = dot2(d2[1], ps, mul2(var[k], p[k]));       // p_sum . right.v
                  cur_ps[k] = ps;
                }
                grp.allreduce(d2);
                dots[0] = d2[0];
                dots[1] = d2[1];
              } else {
                double d6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  const double2 ps = add2(t1_ps[k], cur_ps[k]);    // :390
                  const double2 ps1 = add2(t1_ps[k], cur_lp[k]);   // tree1.p_sum + tree2.left.p (:394)
                  const double2 ps2 = add2(t1_rp[k], cur_ps[k]);   // tree1.right.p + tree2.p_sum (:396)
                  const double2 v1l = mul2(var[k], t1_lp[k]), v1r = mul2(var[k], t1_rp[k]);
                  const double2 v2l = mul2(var[k], cur_lp[k]), v2r = mul2(var[k], p[k]);
                  d6[0] = dot2(d6[0], ps, v1l);
                  d6[1] = dot2(d6[1], ps, v2r);
                  d6[2] = dot2(d6[2], ps1, v1l);
                  d6[3] = dot2(d6[3], ps1, v2l);
                  d6[4] = dot2(d6[4], ps2, v1r);
                  d6[5] = dot2(d6[5], ps2, v2r);
                  cur_ps[k] = ps;
                }
                grp.allreduce(d6);
#pragma unroll
                for (int n = 0; n < 6; ++n) dots[n] = d6[n];
              }
#pragma unroll
              for (int k = 0; k < NP; ++k) cur_lp[k] = t1_lp[k];  // left edge of the merged tree
              const bool turn = (dots[0] <= 0) || (dots[1] <= 0) || (dots[2] <= 0) || (dots[3] <= 0) ||
                                (dots[4] <= 0) || (dots[5] <= 0);  // :391-398 (dots 2..5 preset to 1 at level 0)
              const XF nw = xf_add(XF{ss->wm[lvl], ss->we[lvl]}, cur_w);   // log_size = logaddexp(...)        (:400)
              const XF na = xf_add(XF{ss->am[lvl], ss->ae[lvl]}, cur_a);   // log_weighted_accept_sum      (:401-403)
              // logbern(tree2.log_size - log_size) <=> u * size < size2; the uniform is drawn even when turning (:404)
              const int t1_pslot = ss->pslot[lvl];
              if (xf_u_less(next_uniform(), nw, cur_w)) {
                free_slots |= 1u << t1_pslot;  // keep tree2's proposal, drop tree1's
              } else {
                if (cur_pslot != kLeafProp) free_slots |= 1u << cur_pslot;
                cur_pslot = t1_pslot;
                cur_pE = ss->pE[lvl];
                cur_plogp = ss->plogp[lvl];
              }
              cur_w = nw;
              cur_a = na;
              if (turn) {
                fail = 2;
                break;
              }
              jbits >>= 1;
              ++lvl;
            }
            if (fail) break;
            if (i + 1 < n_leaf_total) {  // push "cur" at level lvl (the last leaf's result stays in registers)
              if (cur_pslot == kLeafProp) {
                cur_pslot = __ffs(free_slots) - 1;
                free_slots &= ~(1u << cur_pslot);
#pragma unroll
                for (int k = 0; k < NP; ++k) vec(vid_prop(cur_pslot))[k * G] = q[k];
              }
              if (lvl == 0) {
#pragma unroll
                for (int k = 0; k < NP; ++k) vec(vid_stack(0, 0))[k * G] = p[k];
              } else {
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                  vec(vid_stack(lvl, 0))[k * G] = cur_lp[k];
                  vec(vid_stack(lvl, 1))[k * G] = p[k];
                  vec(vid_stack(lvl, 2))[k * G] = cur_ps[k];
                }
              }
              // One writer.  Readers see it after at least one group barrier (the next leaf's energy reduction) and
              // finished reading the previous occupant before the barrier that preceded this point.
              if (lane == 0) {
                ss->wm[lvl] = cur_w.m;
                ss->we[lvl] = cur_w.e;
                ss->am[lvl] = cur_a.m;
                ss->ae[lvl] = cur_a.e;
                ss->pE[lvl] = cur_pE;
                ss->plogp[lvl] = cur_plogp;
                ss->pslot[lvl] = cur_pslot;
              }
              if constexpr (G == 32) __syncwarp();
            }
          }
          ++depth;              // nuts.py:315
          n_prop += n_leaves;   // :316
          if (fail) {           // :318-319 -> :216-217 (the subtree is discarded, no uniform is drawn)
            diverging = (fail == 1);
            turning = (fail == 2);
            break;
          }
          // ---- top of _Tree.extend (nuts.py:321-340): T = cur, T.left.p = cur_lp, T.right = z, T.p_sum = cur_ps
          if (xf_u_less(next_uniform(), xf_add(Wp, xf_one()), cur_w)) {  // logbern(tree.log_size - self.log_size) :321-323
            prop_E = cur_pE;
            prop_logp = cur_plogp;
            if (cur_pslot == kLeafProp) {
#pragma unroll
              for (int k = 0; k < NP; ++k) vec(tail + T_PROPQ)[k * G] = q[k];
            } else {
#pragma unroll
              for (int k = 0; k < NP; ++k) vec(tail + T_PROPQ)[k * G] = vec(vid_prop(cur_pslot))[k * G];
            }
          }
          Wp = xf_add(Wp, cur_w);    // log_size = logaddexp(log_size, tree.log_size)                     (:325)
          Acc = xf_add(Acc, cur_a);  // log_weighted_accept_sum                                          (:326-328)
          double d6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            const double2 psum = add2(vec(tail + T_PSUM)[k * G], cur_ps[k]);  // self.p_sum[:] += tree.p_sum (:329)
            vec(tail + T_PSUM)[k * G] = psum;
            const double2 oLp = vec(tail + T_LP)[k * G], oRp = vec(tail + T_RP)[k * G];  // old edges' momenta
            const double2 voL = mul2(var[k], oLp), voR = mul2(var[k], oRp);
            const double2 vTl = mul2(var[k], cur_lp[k]), vTr = mul2(var[k], p[k]);
            if (dir > 0) {
              // left = old left, right = T.right; leftmost = old trajectory with the ALIASED (already
              // updated) p_sum, rightmost = T                                         (:300-303, :333-339)
              const double2 ps1 = add2(psum, cur_lp[k]);   // leftmost_p_sum + rightmost_begin.p
              const double2 ps2 = add2(oRp, cur_ps[k]);    // leftmost_end.p + rightmost_p_sum
              d6[0] = dot2(d6[0], psum, voL);
              d6[1] = dot2(d6[1], psum, vTr);
              d6[2] = dot2(d6[2], ps1, voL);
              d6[3] = dot2(d6[3], ps1, vTl);
              d6[4] = dot2(d6[4], ps2, voR);
              d6[5] = dot2(d6[5], ps2, vTr);
            } else {
              // left = T.right, right = old right; leftmost = T (begin = T.right, end = T.left), rightmost =
              // old trajectory with the aliased p_sum                                  (:309-312, :333-339)
              const double2 ps1 = add2(cur_ps[k], oLp);    // leftmost_p_sum + rightmost_begin.p
              const double2 ps2 = add2(cur_lp[k], psum);   // leftmost_end.p + rightmost_p_sum
              d6[0] = dot2(d6[0], psum, vTr);
              d6[1] = dot2(d6[1], psum, voR);
              d6[2] = dot2(d6[2], ps1, vTr);
              d6[3] = dot2(d6[3], ps1, voL);
              d6[4] = dot2(d6[4], ps2, vTl);
              d6[5] = dot2(d6[5], ps2, voR);
            }
          }
          grp.allreduce(d6);
          if ((d6[0] <= 0) || (d6[1] <= 0) || (d6[2] <= 0) || (d6[3] <= 0) || (d6[4] <= 0) || (d6[5] <= 0)) {
            turning = true;  // :340
            break;
          }
          if (d + 1 < max_depth) {  // self.right / self.left = tree.right (:304 / :313)
            const int base = tail + (dir > 0 ? T_RQ : T_LQ);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              vec(base + 0)[k * G] = q[k];
              vec(base + 1)[k * G] = p[k];
              vec(base + 2)[k * G] = g[k];
            }
            reg_edge = dir;
          }
        }
        (void)turning;
        // _Tree.stats (nuts.py:419-435)
        double mta = 0.0;
        // log_size > 0 <=> exp(log_size) - 1 > 0 in double; exp(lwas - logdiffexp(log_size, 0)) = Acc / (exp(log_size) - 1)
        if (xf_value(Wp) > 0.0) mta = xf_ratio(Acc, Wp);
        accept_stat = mta;
        stat_a = (double)depth;
        stat_b = (double)n_prop;
        stat_energy = prop_E;
        stat_energy_error = prop_E - E0;
        stat_c = max_dE;
        stat_logp = prop_logp;
#pragma unroll
        for (int k = 0; k < NP; ++k) q[k] = vec(tail + T_PROPQ)[k * G];  // hmc_step.end.q
      } else {
        // ---- HamiltonianMC._hamiltonian_step (hmc.py:140-182) ---------------------------------------------------
        double2 q0[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) q0[k] = q[k];
        const double path_length = next_uniform() * a.path_length;  // :141
        const double ratio = path_length / eps;
        int n_steps = ratio >= (double)a.max_steps ? a.max_steps : (int)ratio;  // :142-143 (int() truncates)
        if (n_steps < 1) n_steps = 1;
        double E = E0, logp = logp0;
        for (int s = 0; s < n_steps; ++s) leapfrog(tgt, grp, D, ldh, eps, q, p, g, var, E, logp);  // :149-150
        if (!isfinite(E)) diverging = true;       // :154-155
        double dE = E0 - E;                        // :156
        if (isnan(dE)) dE = -CUDART_INF;           // :157-158
        if (fabs(dE) > a.Emax) diverging = true;   // :159-162
        accept_stat = fmin(1.0, exp(dE));          // :164
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

      // ---- step_adapt.update(accept_stat, adapt_step)  (step_sizes.py:71-92) -----------------------------------
      if (adapt_step) {
        const double w = 1.0 / (da_count + a.t0);
        hbar = (1.0 - w) * hbar + w * (a.target_accept - accept_stat);
        log_step = da_mu - hbar * sqrt(da_count) / a.gamma;
        const double mk = pow(da_count, -a.k);
        log_bar = mk * log_step + (1.0 - mk) * log_bar;
        da_count += 1.0;
      }
      // ---- potential.update(end.q, end.q_grad, tune)  (quadpotential.py:231-245, 322-338) ----------------------
      if (tune && a.adapt_mass) {
        const size_t off = (size_t)chain * a.ld;
        double2* mfg = reinterpret_cast<double2*>(a.mean_fg + off);
        double2* rfg = reinterpret_cast<double2*>(a.rawvar_fg + off);
        double2* mbg = reinterpret_cast<double2*>(a.mean_bg + off);
        double2* rbg = reinterpret_cast<double2*>(a.rawvar_bg + off);
        w_fg += 1.0;
        w_bg += 1.0;
        const double prop_fg = 1.0 / w_fg, prop_bg = 1.0 / w_bg;
        const bool sw = n_samples > 0 && window > 0 && (n_samples % window) == 0;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const int j = lane + k * G;
          if (j < ldh) {
            double2 m = mfg[j], r = rfg[j];
            double2 od = make_double2(add_rn(q[k].x, -m.x), add_rn(q[k].y, -m.y));   // old_diff = x - mean
            m = axpy2(m, prop_fg, od);                                               // mean += prop * old_diff
            double2 nd = make_double2(add_rn(q[k].x, -m.x), add_rn(q[k].y, -m.y));   // new_diff = x - mean
            r = add2(r, mul2(od, nd));                                               // raw_var += old*new
            double2 m2 = mbg[j], r2 = rbg[j];
            od = make_double2(add_rn(q[k].x, -m2.x), add_rn(q[k].y, -m2.y));
            m2 = axpy2(m2, prop_bg, od);
            nd = make_double2(add_rn(q[k].x, -m2.x), add_rn(q[k].y, -m2.y));
            r2 = add2(r2, mul2(od, nd));
            var[k] = make_double2(r.x / w_fg, r.y / w_fg);  // _update_from_weightvar(foreground) (:226-229)
            if (2 * j >= D) var[k].x = 0.0;
            if (2 * j + 1 >= D) var[k].y = 0.0;
            if (sw) {  // foreground <- background, background <- fresh (:240-243)
              mfg[j] = m2;
              rfg[j] = r2;
              mbg[j] = make_double2(0.0, 0.0);
              rbg[j] = make_double2(0.0, 0.0);
            } else {
              mfg[j] = m;
              rfg[j] = r;
              mbg[j] = m2;
              rbg[j] = r2;
            }
          }
        }
        if (sw) {
          w_fg = w_bg;
          w_bg = 0.0;
          window = (long long)((double)window * a.window_multiplier);
        }
        ++n_samples;
      }

      // ---- outputs: trace[:, i] = q (sampling.py:513) and the stats dict (base_hmc.py:185-188) ------------------
      {
        double* tr = a.trace + (size_t)chain * a.trace_chain_stride + (size_t)t * a.trace_draw_stride;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const int j = lane + k * G;
          if (2 * j < D) tr[2 * j] = q[k].x;
          if (2 * j + 1 < D) tr[2 * j + 1] = q[k].y;
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
          srow[LMC_STAT_STEP_SIZE] = exp(log_step);
          srow[LMC_STAT_STEP_SIZE_BAR] = exp(log_bar);
          srow[LMC_STAT_N_UNIFORMS] = (double)uc;
        }
      }
    }

    // ---- write the chain's state back ---------------------------------------------------------------------------
    store_row<G, NP>(a.q + (size_t)chain * a.ld, lane, ldh, q);
    store_row<G, NP>(a.var + (size_t)chain * a.ld, lane, ldh, var);
    if (lane == 0) {
      ad[LMC_ADAPT_LOG_STEP] = log_step;
      ad[LMC_ADAPT_LOG_BAR] = log_bar;
      ad[LMC_ADAPT_HBAR] = hbar;
      ad[LMC_ADAPT_COUNT] = da_count;
      ad[LMC_ADAPT_W_FG] = w_fg;
      ad[LMC_ADAPT_W_BG] = w_bg;
      ad[LMC_ADAPT_NSAMPLES] = (double)n_samples;
      ad[LMC_ADAPT_WINDOW] = (double)window;
      if (status) atomicOr(a.status + chain, status);
    }
  }
}

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

template <class Target, int G, int NP, int KIND>
int launch(const lmc_sampler_args& a, const Target& tgt) {
  constexpr int BLOCK = block_threads<G>();
  constexpr int CPB = BLOCK / G;
  constexpr int VS = G * NP;
  auto kern = sampler_kernel<Target, G, NP, KIND>;
  int dev = 0, n_sm = 0, smem_optin = 0;
  LMC_CUDA(cudaGetDevice(&dev));
  LMC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  LMC_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

  KernelCfg cfg;
  cfg.ws_vecs = KIND == KIND_NUTS ? ws_vecs_nuts(a.max_treedepth) : 0;
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
    const int hot = vid_tail(a.max_treedepth);
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
  long long grid = (long long)n_sm * occ;
  if (a.tune_max_slots > 0 && grid * CPB > a.tune_max_slots) grid = (a.tune_max_slots + CPB - 1) / CPB;
  if (grid > blocks_needed) grid = blocks_needed;
  if (grid < 1) grid = 1;
  const long long need = grid * CPB * (long long)cfg.ws_vecs * (long long)vec_bytes;
  if (need > a.workspace_bytes) return LMC_ERR_WORKSPACE;
  kern<<<(unsigned)grid, BLOCK, smem, (cudaStream_t)a.stream>>>(a, tgt, cfg);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
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


}  // namespace lmc
