// NUTS tree bookkeeping shared by the fused sampler kernel (lmc_sampler.cuh) and the callback-mode state machine
// (lmc_callback.cu): the pieces of reference nuts.py:284-435 that sit between two leapfrog steps.
//
// The recursive _build_subtree (nuts.py:377-417) is run as the iterative binary-counter stack of SURVEY.md A.1: leaf i
// is merged with stack level 0, 1, .. for every trailing 1-bit of i, which visits the merges in the post-order of the
// reference's recursion, so uniforms are consumed in the same order.  A stack entry keeps left.p, right.p, p_sum
// (velocities are recomputed as var*p, bit-identical to the stored ones) and an index into a small pool of
// proposal-position vectors, so choosing a proposal moves an index, never a vector.
#pragma once
#include "lmc_device.cuh"

namespace lmc {

constexpr int kLeafProp = -1;  // "the proposal is the current leaf, still in registers"

// scratch-vector ids (ordered hottest first; ids < n_smem_vecs are shared-memory resident in the fused kernel)
//   0, 1, 2               trajectory p_sum, left edge p, right edge p  (read by every doubling's U-turn checks)
//   3 + 0                 stack level 0: p  (left.p == right.p == p_sum for a single leaf)
//   3 + {1, 2}            proposal slots 0, 1
//   3 + 3+4(l-1)+{0,1,2}  stack level l >= 1: left.p, right.p, p_sum
//   3 + 3+4(l-1)+3        proposal slot l+1
//   tail + T_*            the rest of the trajectory state: edges' q and grad, trajectory proposal q
// ids 0..2: the three trajectory vectors every doubling reads (p_sum, left.p, right.p) -- measured +5% at 1024 x 1000
// over keeping them in the L2-resident tail
constexpr int kHotTail = 3;
__host__ __device__ constexpr int vid_stack(int level, int which) {
  return kHotTail + (level == 0 ? 0 : 3 + 4 * (level - 1) + which);
}
__host__ __device__ constexpr int vid_prop(int slot) { return kHotTail + (slot < 2 ? 1 + slot : 4 * slot - 2); }
__host__ __device__ constexpr int vid_tail(int max_depth) { return kHotTail + (max_depth < 1 ? 3 : 4 * max_depth - 1); }
enum { T_LQ = 0, T_LP, T_LG, T_RQ, T_RP, T_RG, T_PSUM, T_PROPQ, T_COUNT };
// id of trajectory vector `t` (T_*), `tail` = vid_tail(max_treedepth)
__host__ __device__ constexpr int tvid(int tail, int t) {
  return t == T_PSUM ? 0 : t == T_LP ? 1 : t == T_RP ? 2 : tail + t;
}
__host__ __device__ constexpr int ws_vecs_nuts(int max_depth) { return vid_tail(max_depth) + T_COUNT; }
// Depth the scratch is sized for.  Trees grow to early_max_treedepth while tune && iter_count < 200 (nuts.py:205-208),
// and NUTS(max_treedepth=5) keeps the default early_max_treedepth=8 in the reference: both caps bound the stack.
// The *_bytes entry points take this value as their `max_treedepth` argument.
__host__ __device__ inline int scratch_depth(const lmc_sampler_args& a) {
  return a.max_treedepth > a.early_max_treedepth ? a.max_treedepth : a.early_max_treedepth;
}

// per-level scalars of the subtree stack (written by lane 0 only; every read is separated from the write by a group
// barrier / __syncwarp or by a kernel boundary)
struct StackScalars {
  double wm[kMaxDepth], am[kMaxDepth];     // mantissas of exp(log_size), exp(log_weighted_accept_sum)
  double pE[kMaxDepth], plogp[kMaxDepth];  // proposal energy / model_logp
  int we[kMaxDepth], ae[kMaxDepth];        // exponents
  int pslot[kMaxDepth];                    // proposal slot
  int pad[kMaxDepth];
};

// Where a group's scratch vectors live.  Vector `id` occupies VS = G*NP pairs; this thread owns words lane + k*G.
template <int G, int NP>
struct Scratch {
  double2* sm;  // shared-memory part (ids < n_smem), may be null when n_smem == 0
  double2* ws;  // global part (ids are absolute: the first n_smem slots are unused there)
  int n_smem;
  int lane;
  // Pair k of vector `id` (this thread's word lane + k * G).  LMC_SCRATCH_MODE 0 (default): one generic LD / ST through
  // a pointer selected at run time.  Modes 1 / 2 address the two homes in their own address spaces (LDS / STS for the
  // shared-memory ids; plain or .cg global accesses for the workspace): measured SLOWER at every shape (1024 x 1000:
  // 1.06e8 generic, 0.96e8 mode 1, 0.90e8 mode 2 leapfrog/s on the same box, profiles/r02g_scratch_modes.log) -- the
  // second address computation per access costs registers the 128-register kernels do not have (spills), which outweighs
  // the shorter LDS latency.  Kept for the record and for kernels with register headroom.
#ifndef LMC_SCRATCH_MODE
#define LMC_SCRATCH_MODE 0
#endif
  __device__ __forceinline__ double2 ld(int id, int k) const {
#if LMC_SCRATCH_MODE == 0
    return ((id < n_smem ? sm : ws) + (size_t)id * (G * NP) + lane)[k * G];
#else
    if (id < n_smem) {
      const double2* p = sm + (size_t)id * (G * NP) + lane + k * G;
      __builtin_assume(__isShared(p));
      return *p;
    }
#if LMC_SCRATCH_MODE == 1
    return ws[(size_t)id * (G * NP) + lane + k * G];
#else
    return __ldcg(ws + (size_t)id * (G * NP) + lane + k * G);
#endif
#endif
  }
  __device__ __forceinline__ void st(int id, int k, double2 v) const {
#if LMC_SCRATCH_MODE == 0
    ((id < n_smem ? sm : ws) + (size_t)id * (G * NP) + lane)[k * G] = v;
#else
    if (id < n_smem) {
      double2* p = sm + (size_t)id * (G * NP) + lane + k * G;
      __builtin_assume(__isShared(p));
      *p = v;
    } else {
#if LMC_SCRATCH_MODE == 1
      ws[(size_t)id * (G * NP) + lane + k * G] = v;
#else
      __stcg(ws + (size_t)id * (G * NP) + lane + k * G, v);
#endif
    }
#endif
  }
};

// summary of the subtree being assembled on top of the stack ("cur"); its right edge is always the current state z
struct CurTree {
  XF w, a;           // exp(log_size), exp(log_weighted_accept_sum)
  double pE, plogp;  // proposal energy / model_logp
  int pslot;         // proposal slot or kLeafProp
};

// trajectory-level scalars of _Tree (nuts.py:267-282)
struct TrajScalars {
  XF Wp;   // exp(log_size) - 1: total weight of the accepted subtrees (the start point has weight 1)
  XF Acc;  // exp(log_weighted_accept_sum)
  double max_dE, prop_E, prop_logp;
  int depth;
  long long n_prop;
};

// _Tree.__init__ (nuts.py:267-282): both edges and the proposal are the start state, p_sum = start.p.copy()
template <int G, int NP>
__device__ __forceinline__ void tree_init(const Scratch<G, NP>& sc, int tail, const double2 (&q)[NP],
                                          const double2 (&p)[NP], const double2 (&g)[NP]) {
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    sc.st(tvid(tail, T_LQ), k, q[k]);
    sc.st(tvid(tail, T_LP), k, p[k]);
    sc.st(tvid(tail, T_LG), k, g[k]);
    sc.st(tvid(tail, T_RQ), k, q[k]);
    sc.st(tvid(tail, T_RP), k, p[k]);
    sc.st(tvid(tail, T_RG), k, g[k]);
    sc.st(tvid(tail, T_PSUM), k, p[k]);
    sc.st(tvid(tail, T_PROPQ), k, q[k]);
  }
}

// _single_step after the leapfrog (nuts.py:352-375).  Returns false when the leaf diverges.
template <int NP>
__device__ __forceinline__ bool leaf_init(double E, double logp, double E0, double Emax, const double2 (&p)[NP],
                                          double& max_dE, CurTree& cur, double2 (&cur_lp)[NP], double2 (&cur_ps)[NP]) {
  double dE = E - E0;                         // :352
  if (isnan(dE)) dE = CUDART_INF;             // :353-354
  if (fabs(dE) > fabs(max_dE)) max_dE = dE;   // :356-357
  if (!(fabs(dE) < Emax)) return false;       // :358 / :370-375
  cur.w = xf_exp(-dE);                                 // log_size = -dE
  cur.a = (-dE < 0.0) ? xf_sqr(cur.w) : cur.w;         // log_p_accept_weighted = -dE + min(0, -dE)  (:363)
  cur.pE = E;
  cur.plogp = logp;
  cur.pslot = kLeafProp;
#pragma unroll
  for (int k = 0; k < NP; ++k) cur_lp[k] = cur_ps[k] = p[k];
  return true;
}

// The vector operands of the tree arithmetic are reached through small accessor functors, so that ONE copy of every
// formula (and of the reference's p_sum aliasing quirk) serves kernels that keep them in registers (arrays), in shared
// memory (the lean kernel) or behind an id into the scratch (a left edge that is never copied): a functor `f(k)` returns
// pair k of this thread; `set_ps(k, v)` / `set_lp(k, v)` store the merged subtree's p_sum / left-edge momentum.
template <int NP>
struct PairArray {
  const double2 (&a)[NP];
  __device__ __forceinline__ double2 operator()(int k) const { return a[k]; }
};
template <int NP>
struct PairArrayOut {
  double2 (&a)[NP];
  __device__ __forceinline__ void operator()(int k, double2 v) const { a[k] = v; }
};

// scalar half of one merge (nuts.py:400-407): sizes, weighted accept sums, proposal choice with the uniform `u`
__device__ __forceinline__ void merge_scalars(const StackScalars* ss, int lvl, CurTree& cur, unsigned& free_slots, double u) {
  const XF nw = xf_add(XF{ss->wm[lvl], ss->we[lvl]}, cur.w);  // log_size = logaddexp(...)             (:400)
  const XF na = xf_add(XF{ss->am[lvl], ss->ae[lvl]}, cur.a);  // log_weighted_accept_sum           (:401-403)
  // logbern(tree2.log_size - log_size) <=> u * size < size2                                            (:404)
  const int t1_pslot = ss->pslot[lvl];
  if (xf_u_less(u, nw, cur.w)) {
    free_slots |= 1u << t1_pslot;  // keep tree2's proposal, drop tree1's
  } else {
    if (cur.pslot != kLeafProp) free_slots |= 1u << cur.pslot;
    cur.pslot = t1_pslot;
    cur.pE = ss->pE[lvl];
    cur.plogp = ss->plogp[lvl];
  }
  cur.w = nw;
  cur.a = na;
}

// One merge of _build_subtree at stack level lvl >= 1 (nuts.py:387-417): tree1 = stack entry `lvl`, tree2 = cur (left
// edge momentum lp(k), p_sum ps(k), right edge momentum p(k)).  `u` is the uniform of logbern (:404), drawn by the
// caller even when the merged tree turns.  Returns `turning`.
// BATCH: all of tree1's scratch words are loaded before anything is stored.  Scratch accesses are generic loads / stores
// that the compiler must keep in program order (a store may alias the next load), so the plain form pays one dependent
// L2 round trip per pair; kernels with the registers to hold 3 x NP pairs for a moment ask for the batched form.
template <int G, int NP, bool BATCH = false, class Grp, class VarF, class PF, class LpF, class PsF, class SetPsF,
          class SetLpF, class UF>
__device__ __forceinline__ bool merge_upper(const Scratch<G, NP>& sc, Grp& grp, const StackScalars* ss, int lvl,
                                            VarF var, PF p, LpF lp, PsF ps, SetPsF set_ps, SetLpF set_lp, CurTree& cur,
                                            unsigned& free_slots, UF next_u) {
  double d6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double2 b_lp[BATCH ? NP : 1], b_rp[BATCH ? NP : 1], b_ps[BATCH ? NP : 1];
  if constexpr (BATCH) {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      b_lp[k] = sc.ld(vid_stack(lvl, 0), k);
      b_rp[k] = sc.ld(vid_stack(lvl, 1), k);
      b_ps[k] = sc.ld(vid_stack(lvl, 2), k);
    }
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const double2 t1_lp = BATCH ? b_lp[BATCH ? k : 0] : sc.ld(vid_stack(lvl, 0), k);
    const double2 t1_rp = BATCH ? b_rp[BATCH ? k : 0] : sc.ld(vid_stack(lvl, 1), k);
    const double2 t1_ps = BATCH ? b_ps[BATCH ? k : 0] : sc.ld(vid_stack(lvl, 2), k);
    const double2 c_lp = lp(k), c_ps = ps(k), vk = var(k), pk = p(k);
    const double2 nps = add2(t1_ps, c_ps);   // p_sum = tree1.p_sum + tree2.p_sum (:390)
    const double2 ps1 = add2(t1_ps, c_lp);   // tree1.p_sum + tree2.left.p (:394)
    const double2 ps2 = add2(t1_rp, c_ps);   // tree1.right.p + tree2.p_sum (:396)
    const double2 v1l = mul2(vk, t1_lp), v1r = mul2(vk, t1_rp);
    const double2 v2l = mul2(vk, c_lp), v2r = mul2(vk, pk);
    d6[0] = dot2(d6[0], nps, v1l);
    d6[1] = dot2(d6[1], nps, v2r);
    d6[2] = dot2(d6[2], ps1, v1l);
    d6[3] = dot2(d6[3], ps1, v2l);
    d6[4] = dot2(d6[4], ps2, v1r);
    d6[5] = dot2(d6[5], ps2, v2r);
    set_ps(k, nps);
    set_lp(k, t1_lp);  // left edge of the merged tree
  }
  grp.allreduce(d6);
  const bool turn = (d6[0] <= 0) || (d6[1] <= 0) || (d6[2] <= 0) || (d6[3] <= 0) || (d6[4] <= 0) || (d6[5] <= 0);  // :391-398
  merge_scalars(ss, lvl, cur, free_slots, next_u());  // the uniform is drawn whether or not the merged tree turns
  return turn;
}

// The same at level 0: tree1 is a single leaf (left.p == right.p == p_sum == the stored momentum), and so is tree2 when
// it comes straight from the leapfrog (lp == ps == p).
template <int G, int NP, class Grp, class VarF, class PF, class PsF, class SetPsF, class SetLpF, class UF>
__device__ __forceinline__ bool merge_leaves(const Scratch<G, NP>& sc, Grp& grp, const StackScalars* ss, VarF var,
                                             PF p, PsF ps, SetPsF set_ps, SetLpF set_lp, CurTree& cur,
                                             unsigned& free_slots, UF next_u) {
  double d2[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const double2 t1p = sc.ld(vid_stack(0, 0), k);
    const double2 vk = var(k);
    const double2 nps = add2(t1p, ps(k));           // p_sum = tree1.p_sum + tree2.p_sum (:390)
    d2[0] = dot2(d2[0], nps, mul2(vk, t1p));        // p_sum . left.v
    d2[1] = dot2(d2[1], nps, mul2(vk, p(k)));       // p_sum . right.v
    set_ps(k, nps);
    set_lp(k, t1p);
  }
  grp.allreduce(d2);
  const bool turn = (d2[0] <= 0) || (d2[1] <= 0);  // :391
  merge_scalars(ss, 0, cur, free_slots, next_u());
  return turn;
}

// One merge of _build_subtree (nuts.py:387-417) with every operand in registers: tree1 = stack entry `lvl`, tree2 = cur
// (right edge momentum p).  Returns `turning`.
template <int G, int NP>
__device__ __forceinline__ bool merge_level(const Scratch<G, NP>& sc, Group<G>& grp, const StackScalars* ss, int lvl,
                                            const double2 (&var)[NP], const double2 (&p)[NP], double2 (&cur_lp)[NP],
                                            double2 (&cur_ps)[NP], CurTree& cur, unsigned& free_slots, double u) {
  if (lvl == 0)
    return merge_leaves<G, NP>(sc, grp, ss, PairArray<NP>{var}, PairArray<NP>{p}, PairArray<NP>{cur_ps},
                               PairArrayOut<NP>{cur_ps}, PairArrayOut<NP>{cur_lp}, cur, free_slots, [u] { return u; });
  return merge_upper<G, NP>(sc, grp, ss, lvl, PairArray<NP>{var}, PairArray<NP>{p}, PairArray<NP>{cur_lp},
                            PairArray<NP>{cur_ps}, PairArrayOut<NP>{cur_ps}, PairArrayOut<NP>{cur_lp}, cur, free_slots,
                            [u] { return u; });
}

// ---- level-0 specialisations used by the fused kernel ------------------------------------------------------------------
// A single leaf has left.p == right.p == p_sum == p, so it needs no cur_lp / cur_ps registers: an even leaf goes straight
// to stack level 0 (push_leaf), an odd leaf is merged with it (merge_leaf_pair), which is where cur_lp / cur_ps are first
// defined.  Same arithmetic, same order of operations and of uniforms as leaf_init + merge_level(lvl = 0) + push_cur.

// scalar part of _single_step after the leapfrog (nuts.py:352-368).  Returns false when the leaf diverges.
__device__ __forceinline__ bool leaf_scalars(double E, double logp, double E0, double Emax, double& max_dE, CurTree& cur) {
  double dE = E - E0;                         // :352
  if (isnan(dE)) dE = CUDART_INF;             // :353-354
  if (fabs(dE) > fabs(max_dE)) max_dE = dE;   // :356-357
  if (!(fabs(dE) < Emax)) return false;       // :358 / :370-375
  cur.w = xf_exp(-dE);                                 // log_size = -dE
  cur.a = (-dE < 0.0) ? xf_sqr(cur.w) : cur.w;         // log_p_accept_weighted = -dE + min(0, -dE)  (:363)
  cur.pE = E;
  cur.plogp = logp;
  cur.pslot = kLeafProp;
  return true;
}

// push the current leaf (state q, p) as stack entry 0
template <int G, int NP>
__device__ __forceinline__ void push_leaf(const Scratch<G, NP>& sc, StackScalars* ss, const double2 (&q)[NP],
                                          const double2 (&p)[NP], CurTree& cur, unsigned& free_slots) {
  cur.pslot = __ffs(free_slots) - 1;
  free_slots &= ~(1u << cur.pslot);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    sc.st(vid_prop(cur.pslot), k, q[k]);
    sc.st(vid_stack(0, 0), k, p[k]);
  }
  if (sc.lane == 0) {
    ss->wm[0] = cur.w.m;
    ss->we[0] = cur.w.e;
    ss->am[0] = cur.a.m;
    ss->ae[0] = cur.a.e;
    ss->pE[0] = cur.pE;
    ss->plogp[0] = cur.plogp;
    ss->pslot[0] = cur.pslot;
  }
}

// merge stack entry 0 (a leaf) with the current leaf (momentum p): defines cur_lp, cur_ps.  Returns `turning`.
template <int G, int NP>
__device__ __forceinline__ bool merge_leaf_pair(const Scratch<G, NP>& sc, Group<G>& grp, const StackScalars* ss,
                                                const double2 (&var)[NP], const double2 (&p)[NP], double2 (&cur_lp)[NP],
                                                double2 (&cur_ps)[NP], CurTree& cur, unsigned& free_slots, double u) {
  return merge_leaves<G, NP>(sc, grp, ss, PairArray<NP>{var}, PairArray<NP>{p}, PairArray<NP>{p}, PairArrayOut<NP>{cur_ps},
                             PairArrayOut<NP>{cur_lp}, cur, free_slots, [u] { return u; });
}

// Push "cur" (right edge state q, p) as stack entry `lvl`.  One writer for the scalars (lane 0).
template <int G, int NP>
__device__ __forceinline__ void push_cur(const Scratch<G, NP>& sc, StackScalars* ss, int lvl, const double2 (&q)[NP],
                                         const double2 (&p)[NP], const double2 (&cur_lp)[NP],
                                         const double2 (&cur_ps)[NP], CurTree& cur, unsigned& free_slots) {
  if (cur.pslot == kLeafProp) {
    cur.pslot = __ffs(free_slots) - 1;
    free_slots &= ~(1u << cur.pslot);
#pragma unroll
    for (int k = 0; k < NP; ++k) sc.st(vid_prop(cur.pslot), k, q[k]);
  }
  if (lvl == 0) {
#pragma unroll
    for (int k = 0; k < NP; ++k) sc.st(vid_stack(0, 0), k, p[k]);
  } else {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      sc.st(vid_stack(lvl, 0), k, cur_lp[k]);
      sc.st(vid_stack(lvl, 1), k, p[k]);
      sc.st(vid_stack(lvl, 2), k, cur_ps[k]);
    }
  }
  if (sc.lane == 0) {
    ss->wm[lvl] = cur.w.m;
    ss->we[lvl] = cur.w.e;
    ss->am[lvl] = cur.a.m;
    ss->ae[lvl] = cur.a.e;
    ss->pE[lvl] = cur.pE;
    ss->plogp[lvl] = cur.plogp;
    ss->pslot[lvl] = cur.pslot;
  }
}

// Top of _Tree.extend after a completed subtree T = cur (nuts.py:321-340): T.left.p = lp(k), T.right = z = (q, p),
// T.p_sum = ps(k).  `u` is the uniform of the biased progressive accept (:321-323).  Returns `turning`.
// BATCH: as in merge_upper -- the old trajectory's words (and a proposal that moves) are loaded before anything is stored.
template <int G, int NP, bool BATCH = false, class Grp, class VarF, class QF, class PF, class LpF, class PsF>
__device__ __forceinline__ bool extend_top_f(const Scratch<G, NP>& sc, Grp& grp, int tail, int dir, VarF var, QF q,
                                             PF p, LpF lp, PsF ps, const CurTree& cur, TrajScalars& tr, double u) {
  double2 b_ps[BATCH ? NP : 1], b_lp[BATCH ? NP : 1], b_rp[BATCH ? NP : 1];
  if constexpr (BATCH) {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      b_ps[k] = sc.ld(tvid(tail, T_PSUM), k);
      b_lp[k] = sc.ld(tvid(tail, T_LP), k);
      b_rp[k] = sc.ld(tvid(tail, T_RP), k);
    }
  }
  if (xf_u_less(u, xf_add(tr.Wp, xf_one()), cur.w)) {  // logbern(tree.log_size - self.log_size) :321-323
    tr.prop_E = cur.pE;
    tr.prop_logp = cur.plogp;
    if constexpr (BATCH) {
      double2 t[NP];
      if (cur.pslot == kLeafProp) {
#pragma unroll
        for (int k = 0; k < NP; ++k) t[k] = q(k);
      } else {
#pragma unroll
        for (int k = 0; k < NP; ++k) t[k] = sc.ld(vid_prop(cur.pslot), k);
      }
#pragma unroll
      for (int k = 0; k < NP; ++k) sc.st(tvid(tail, T_PROPQ), k, t[k]);
    } else if (cur.pslot == kLeafProp) {
#pragma unroll
      for (int k = 0; k < NP; ++k) sc.st(tvid(tail, T_PROPQ), k, q(k));
    } else {
#pragma unroll
      for (int k = 0; k < NP; ++k) sc.st(tvid(tail, T_PROPQ), k, sc.ld(vid_prop(cur.pslot), k));
    }
  }
  tr.Wp = xf_add(tr.Wp, cur.w);    // log_size = logaddexp(log_size, tree.log_size)                     (:325)
  tr.Acc = xf_add(tr.Acc, cur.a);  // log_weighted_accept_sum                                          (:326-328)
  double d6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const double2 c_ps = ps(k), c_lp = lp(k), vk = var(k), pk = p(k);
    const double2 psum = add2(BATCH ? b_ps[BATCH ? k : 0] : sc.ld(tvid(tail, T_PSUM), k), c_ps);  // self.p_sum[:] += tree.p_sum (:329)
    sc.st(tvid(tail, T_PSUM), k, psum);
    const double2 oLp = BATCH ? b_lp[BATCH ? k : 0] : sc.ld(tvid(tail, T_LP), k);  // old edges' momenta
    const double2 oRp = BATCH ? b_rp[BATCH ? k : 0] : sc.ld(tvid(tail, T_RP), k);
    const double2 voL = mul2(vk, oLp), voR = mul2(vk, oRp);
    const double2 vTl = mul2(vk, c_lp), vTr = mul2(vk, pk);
    if (dir > 0) {
      // left = old left, right = T.right; leftmost = old trajectory with the ALIASED (already updated) p_sum,
      // rightmost = T                                                          (:300-303, :333-339)
      const double2 ps1 = add2(psum, c_lp);  // leftmost_p_sum + rightmost_begin.p
      const double2 ps2 = add2(oRp, c_ps);   // leftmost_end.p + rightmost_p_sum
      d6[0] = dot2(d6[0], psum, voL);
      d6[1] = dot2(d6[1], psum, vTr);
      d6[2] = dot2(d6[2], ps1, voL);
      d6[3] = dot2(d6[3], ps1, vTl);
      d6[4] = dot2(d6[4], ps2, voR);
      d6[5] = dot2(d6[5], ps2, vTr);
    } else {
      // left = T.right, right = old right; leftmost = T (begin = T.right, end = T.left), rightmost = old trajectory
      // with the aliased p_sum                                                 (:309-312, :333-339)
      const double2 ps1 = add2(c_ps, oLp);   // leftmost_p_sum + rightmost_begin.p
      const double2 ps2 = add2(c_lp, psum);  // leftmost_end.p + rightmost_p_sum
      d6[0] = dot2(d6[0], psum, vTr);
      d6[1] = dot2(d6[1], psum, voR);
      d6[2] = dot2(d6[2], ps1, vTr);
      d6[3] = dot2(d6[3], ps1, voL);
      d6[4] = dot2(d6[4], ps2, vTl);
      d6[5] = dot2(d6[5], ps2, voR);
    }
  }
  grp.allreduce(d6);
  return (d6[0] <= 0) || (d6[1] <= 0) || (d6[2] <= 0) || (d6[3] <= 0) || (d6[4] <= 0) || (d6[5] <= 0);  // :333-340
}
template <int G, int NP>
__device__ __forceinline__ bool extend_top(const Scratch<G, NP>& sc, Group<G>& grp, int tail, int dir,
                                           const double2 (&var)[NP], const double2 (&q)[NP], const double2 (&p)[NP],
                                           const double2 (&cur_lp)[NP], const double2 (&cur_ps)[NP],
                                           const CurTree& cur, TrajScalars& tr, double u) {
  return extend_top_f<G, NP>(sc, grp, tail, dir, PairArray<NP>{var}, PairArray<NP>{q}, PairArray<NP>{p},
                             PairArray<NP>{cur_lp}, PairArray<NP>{cur_ps}, cur, tr, u);
}

// _Tree.stats: mean_tree_accept (nuts.py:419-425).  log_size > 0 <=> exp(log_size) - 1 > 0 in double;
// exp(lwas - logdiffexp(log_size, 0)) = Acc / (exp(log_size) - 1)
__device__ __forceinline__ double mean_tree_accept(const TrajScalars& tr) {
  return xf_value(tr.Wp) > 0.0 ? xf_ratio(tr.Acc, tr.Wp) : 0.0;
}

// ---- HamiltonianMC._hamiltonian_step scalars (hmc.py:141-164) -----------------------------------------------------------
// n_steps = max(1, int(path_length / step_size)); n_steps = min(max_steps, n_steps)   (:142-143, int() truncates)
__device__ __forceinline__ int hmc_n_steps(double path_length, double eps, int max_steps) {
  const double ratio = path_length / eps;
  int n = ratio >= (double)max_steps ? max_steps : (int)ratio;
  return n < 1 ? 1 : n;
}
// energy_change = start.energy - end.energy with the divergence rules of :154-162; accept_stat = min(1, exp(dE)) (:164).
// Returns `diverging`.
__device__ __forceinline__ bool hmc_energy_check(double E0, double E, double Emax, double& dE, double& accept_stat) {
  bool diverging = !isfinite(E);          // :154-155
  dE = E0 - E;                            // :156
  if (isnan(dE)) dE = -CUDART_INF;        // :157-158
  if (fabs(dE) > Emax) diverging = true;  // :159-162
  accept_stat = fmin(1.0, exp(dE));       // :164
  return diverging;
}

// ---- the two adaptations that close BaseHMC._astep (base_hmc.py:161-162) ---------------------------------------------
struct DualAvg {
  double log_step, log_bar, hbar, count, mu;
};
// DualAverageAdaptation.update (step_sizes.py:71-92)
LMC_COLD DualAvg dual_average_step(DualAvg s, double accept_stat, double target, double gamma, double k, double t0) {
  const double w = 1.0 / (s.count + t0);
  s.hbar = (1.0 - w) * s.hbar + w * (target - accept_stat);
  s.log_step = s.mu - s.hbar * sqrt(s.count) / gamma;
  // count ** -k (step_sizes.py:86) as exp(-k log count): count >= 1, |k log count| < 20, so the result agrees with a
  // correctly rounded pow to ~1e-15 relative (the inlined double-precision pow is 500 instructions)
  const double mk = exp(-k * log(s.count));
  s.log_bar = mk * s.log_step + (1.0 - mk) * s.log_bar;
  s.count += 1.0;
  return s;
}
__device__ __forceinline__ void dual_average_update(DualAvg& s, double accept_stat, double target, double gamma,
                                                    double k, double t0) {
  s = dual_average_step(s, accept_stat, target, gamma, k, t0);
}
// exp() for once-per-transition statistics
LMC_COLD double exp_cold(double x) { return exp(x); }

struct WelfordScalars {
  double w_fg, w_bg;
  long long n_samples, window;
};
// QuadPotentialDiagAdapt.update (quadpotential.py:231-245) with _WeightedVariance.add_sample (:322-338) on both
// estimators, the var refresh from the foreground (:226-229) and the window switch (:240-243).  Rows live in HBM and
// may have been written by another SM (ld.cg).
template <int G, int NP>
__device__ __forceinline__ void welford_update(int lane, int D, int ldh, double* mean_fg, double* rawvar_fg,
                                               double* mean_bg, double* rawvar_bg, const double2 (&q)[NP],
                                               double2 (&var)[NP], WelfordScalars& ws, double window_multiplier) {
  double2* mfg = reinterpret_cast<double2*>(mean_fg);
  double2* rfg = reinterpret_cast<double2*>(rawvar_fg);
  double2* mbg = reinterpret_cast<double2*>(mean_bg);
  double2* rbg = reinterpret_cast<double2*>(rawvar_bg);
  ws.w_fg += 1.0;
  ws.w_bg += 1.0;
  const double prop_fg = 1.0 / ws.w_fg, prop_bg = 1.0 / ws.w_bg;
  const bool sw = ws.n_samples > 0 && ws.window > 0 && (ws.n_samples % ws.window) == 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    if (j < ldh) {
      double2 m = ldcg2(mfg + j), r = ldcg2(rfg + j);
      double2 od = make_double2(add_rn(q[k].x, -m.x), add_rn(q[k].y, -m.y));  // old_diff = x - mean
      m = axpy2(m, prop_fg, od);                                              // mean += prop * old_diff
      double2 nd = make_double2(add_rn(q[k].x, -m.x), add_rn(q[k].y, -m.y));  // new_diff = x - mean
      r = add2(r, mul2(od, nd));                                              // raw_var += old*new
      double2 m2 = ldcg2(mbg + j), r2 = ldcg2(rbg + j);
      od = make_double2(add_rn(q[k].x, -m2.x), add_rn(q[k].y, -m2.y));
      m2 = axpy2(m2, prop_bg, od);
      nd = make_double2(add_rn(q[k].x, -m2.x), add_rn(q[k].y, -m2.y));
      r2 = add2(r2, mul2(od, nd));
      var[k] = make_double2(div_cold(r.x, ws.w_fg), div_cold(r.y, ws.w_fg));  // _update_from_weightvar(fg) (:226-229)
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
    ws.w_fg = ws.w_bg;
    ws.w_bg = 0.0;
    ws.window = (long long)((double)ws.window * window_multiplier);
  }
  ++ws.n_samples;
}

// The same update with every row loaded before anything is computed or stored (one HBM round trip for the four rows
// instead of a dependent chain per pair): for kernels with the registers to hold 4 x NP pairs for a moment.  Identical
// results.
template <int G, int NP>
__device__ __forceinline__ void welford_update_batched(int lane, int D, int ldh, double* mean_fg, double* rawvar_fg,
                                                       double* mean_bg, double* rawvar_bg, const double2 (&q)[NP],
                                                       double2 (&var)[NP], WelfordScalars& ws, double window_multiplier) {
  double2* mfg = reinterpret_cast<double2*>(mean_fg);
  double2* rfg = reinterpret_cast<double2*>(rawvar_fg);
  double2* mbg = reinterpret_cast<double2*>(mean_bg);
  double2* rbg = reinterpret_cast<double2*>(rawvar_bg);
  double2 m[NP], r[NP], m2[NP], r2[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    const double2 z = make_double2(0.0, 0.0);
    m[k] = j < ldh ? ldcg2(mfg + j) : z;
    r[k] = j < ldh ? ldcg2(rfg + j) : z;
    m2[k] = j < ldh ? ldcg2(mbg + j) : z;
    r2[k] = j < ldh ? ldcg2(rbg + j) : z;
  }
  ws.w_fg += 1.0;
  ws.w_bg += 1.0;
  const double prop_fg = 1.0 / ws.w_fg, prop_bg = 1.0 / ws.w_bg;
  const bool sw = ws.n_samples > 0 && ws.window > 0 && (ws.n_samples % ws.window) == 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    double2 od = make_double2(add_rn(q[k].x, -m[k].x), add_rn(q[k].y, -m[k].y));  // old_diff = x - mean
    m[k] = axpy2(m[k], prop_fg, od);                                            // mean += prop * old_diff
    double2 nd = make_double2(add_rn(q[k].x, -m[k].x), add_rn(q[k].y, -m[k].y));  // new_diff = x - mean
    r[k] = add2(r[k], mul2(od, nd));                                            // raw_var += old*new
    od = make_double2(add_rn(q[k].x, -m2[k].x), add_rn(q[k].y, -m2[k].y));
    m2[k] = axpy2(m2[k], prop_bg, od);
    nd = make_double2(add_rn(q[k].x, -m2[k].x), add_rn(q[k].y, -m2[k].y));
    r2[k] = add2(r2[k], mul2(od, nd));
  }
  div_pairs<NP>(r, ws.w_fg, var);  // _update_from_weightvar(fg) (:226-229)
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    if (2 * j >= D) var[k].x = 0.0;
    if (2 * j + 1 >= D) var[k].y = 0.0;
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    if (j < ldh) {
      if (sw) {  // foreground <- background, background <- fresh (:240-243)
        mfg[j] = m2[k];
        rfg[j] = r2[k];
        mbg[j] = make_double2(0.0, 0.0);
        rbg[j] = make_double2(0.0, 0.0);
      } else {
        mfg[j] = m[k];
        rfg[j] = r[k];
        mbg[j] = m2[k];
        rbg[j] = r2[k];
      }
    }
  }
  if (sw) {
    ws.w_fg = ws.w_bg;
    ws.w_bg = 0.0;
    ws.window = (long long)((double)ws.window * window_multiplier);
  }
  ++ws.n_samples;
}

// p0 = potential.random() (quadpotential.py:221-224 / 374-376): inv_stds * normals, inv_stds = 1/sqrt(var) with IEEE
// sqrt and divide, identical to the reference's stored arrays.  TAPE: `normals_row` = this transition's D normals.
template <int G, int NP>
__device__ __forceinline__ void draw_momentum(int lane, int D, const double* normals_row, uint64_t seed, long long it,
                                              const double2 (&var)[NP], double2 (&p)[NP]) {
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
    p[k].x = (2 * j < D) ? mul_rn(inv_sqrt_cold(var[k].x), n.x) : 0.0;
    p[k].y = (2 * j + 1 < D) ? mul_rn(inv_sqrt_cold(var[k].y), n.y) : 0.0;
  }
}

}  // namespace lmc
