// Dense-mass mode: NUTS / HMC transitions with a dense mass matrix (reference quadpotential.py:390-615:
// QuadPotentialFull / QuadPotentialFullInv / QuadPotentialFullAdapt) -- the state machine of lmc_callback.cu with the
// velocity v = M^-1 p and the momentum draw p0 = potential.random() as EXTERNAL batched operations next to the
// gradient, plus the two kernels those operations are made of for per-chain matrices: a batched symmetric-agnostic
// matrix-vector product (the HBM-bound hot op of this mode) and the Welford covariance update.
//
// Phases of a chain (include/lmc_b200.h, lmc_dense_args):
//     START    needs GRAD (at the transition's start position) and MOM (p0 from this transition's normals)
//     START_V  needs VEL  of (p0, g0)           -> E0, tree init, first half step
//     LEAF_G   needs GRAD at the new position    -> p' = p_half + dt g'
//     LEAF_V   needs VEL  of (p', g')            -> energy, leaf + merges (+ end of doubling / transition), half step
// The current phase-space point between launches lives in the evaluation buffers themselves:
//     q_eval = q,  x_eval[0] = p,  x_eval[1] = g,  v_eval[0] = v = M^-1 p,  v_eval[1] = w = M^-1 g.
// The half-kicked velocity of integration.py:108-111 is v + dt w (velocity is linear), so a leapfrog costs one GRAD
// and one two-vector VEL.  Tree arithmetic follows lmc_tree.cuh (same order of operations, of uniforms and of
// decisions); the difference is that velocities are stored with every momentum instead of recomputed as var * p.
#include <cuda_runtime.h>
#include <stdint.h>

#include "lmc_common.h"
#include "lmc_device.cuh"
#include "lmc_tree.cuh"

namespace lmc {

enum { DK_NUTS = 0, DK_HMC = 1 };
enum { DPH_START = 0, DPH_START_V = 1, DPH_LEAF_G = 2, DPH_LEAF_V = 3, DPH_DONE = 4 };

struct DnScalars {
  int phase, t, d, dir;
  int max_depth, last_dir, n_steps, pad0;
  unsigned i, uc, free_slots, pad1;
  long long n_leaves;
  double E0, logp0, eps, path_length, logp_leaf;
  TrajScalars tr;
};
struct DnMachine {
  DnScalars s;
  StackScalars ss;
};
constexpr size_t kDnMachineBytes = (sizeof(DnMachine) + 255) & ~(size_t)255;

// scratch-vector ids (all in the per-chain global workspace)
enum { S_LP = 0, S_LV, S_RP, S_RV, S_PS, S_COUNT };
enum { DT_LQ = 0, DT_LP, DT_LG, DT_LV, DT_LW, DT_RQ, DT_RP, DT_RG, DT_RV, DT_RW, DT_PSUM, DT_PROPQ, DT_COUNT };
__host__ __device__ constexpr int dv_stack(int level, int which) { return S_COUNT * level + which; }
__host__ __device__ constexpr int dv_prop(int max_depth, int slot) { return S_COUNT * max_depth + slot; }
__host__ __device__ constexpr int dv_tail(int max_depth) { return S_COUNT * max_depth + max_depth + 1; }
__host__ __device__ constexpr int dn_vecs(int kind, int max_depth) { return kind == DK_NUTS ? dv_tail(max_depth) + DT_COUNT : 1; }

template <int G>
__host__ __device__ constexpr int dn_block() { return G >= 64 ? G : 128; }

static bool dn_pick_shape(int ndim, int* G, int* NP) {
  const int pairs = (ndim + 1) / 2;
  static const int table[][2] = {{32, 1}, {64, 1}, {128, 1}, {256, 1}, {256, 2}, {512, 2}, {512, 4}, {1024, 4}};
  for (auto& s : table)
    if (s[0] * s[1] >= pairs) { *G = s[0]; *NP = s[1]; return true; }
  return false;
}

// standard normals of the momentum draw of transition `t` of this call -> n_eval row (tape or Philox;
// quadpotential.py:416,455), written by `nthreads` cooperating threads
__device__ __forceinline__ void write_normals(const lmc_sampler_args& a, int chain, int t, int tid, int nthreads,
                                              double* n_row) {
  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  const long long it = a.iter0 + t;
  const uint64_t seed = (a.rng.mode == LMC_RNG_PHILOX) ? a.rng.seeds[chain] : 0ull;
  const double* tape = a.rng.mode == LMC_RNG_TAPE ? a.rng.normals + ((size_t)chain * a.n_trans + t) * D : nullptr;
  for (int j = tid; j < ldh; j += nthreads) {
    double2 n = make_double2(0.0, 0.0);
    if (tape) {
      if (2 * j < D) n.x = tape[2 * j];
      if (2 * j + 1 < D) n.y = tape[2 * j + 1];
    } else if (2 * j < D) {
      n = philox_normal_pair(seed, it, (uint32_t)j);
      if (2 * j + 1 >= D) n.y = 0.0;
    }
    reinterpret_cast<double2*>(n_row)[j] = n;
  }
}

__global__ void dn_begin_kernel(const lmc_dense_args c) {
  const lmc_sampler_args& a = c.base;
  const int chain = blockIdx.x;
  DnMachine* M = reinterpret_cast<DnMachine*>(reinterpret_cast<char*>(c.machine) + (size_t)chain * kDnMachineBytes);
  const bool run = a.n_trans > 0;
  if (threadIdx.x == 0) {
    M->s.phase = run ? DPH_START : DPH_DONE;
    M->s.t = 0;
    c.need[chain] = run ? (LMC_NEED_GRAD | LMC_NEED_MOM) : 0;
    if (chain == 0) *c.n_running = run ? a.n_chains : 0;
  }
  for (int e = threadIdx.x; e < (int)a.ld; e += blockDim.x)
    c.q_eval[(size_t)chain * a.ld + e] = e < a.ndim ? a.q[(size_t)chain * a.ld + e] : 0.0;
  if (run) write_normals(a, chain, 0, threadIdx.x, blockDim.x, c.n_eval + (size_t)chain * a.ld);
}

// ---- tree pieces with stored velocities (lmc_tree.cuh: merge_level / push_cur / extend_top) --------------------------
template <int G, int NP>
__device__ __forceinline__ bool dn_merge_level(const Scratch<G, NP>& sc, Group<G>& grp, const StackScalars* ss, int lvl,
                                               const double2 (&p)[NP], const double2 (&v)[NP], double2 (&cur_lp)[NP],
                                               double2 (&cur_lv)[NP], double2 (&cur_ps)[NP], CurTree& cur,
                                               unsigned& free_slots, double u) {
  double2 t1_lp[NP], t1_lv[NP], t1_rp[NP], t1_rv[NP], t1_ps[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    t1_lp[k] = sc.ld(dv_stack(lvl, S_LP), k);
    t1_lv[k] = sc.ld(dv_stack(lvl, S_LV), k);
    t1_rp[k] = sc.ld(dv_stack(lvl, S_RP), k);
    t1_rv[k] = sc.ld(dv_stack(lvl, S_RV), k);
    t1_ps[k] = sc.ld(dv_stack(lvl, S_PS), k);
  }
  bool turn;
  if (lvl == 0) {
    double d2[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const double2 ps = add2(t1_ps[k], cur_ps[k]);  // p_sum = tree1.p_sum + tree2.p_sum (nuts.py:390)
      d2[0] = dot2(d2[0], ps, t1_lv[k]);             // p_sum . left.v
      d2[1] = dot2(d2[1], ps, v[k]);                 // p_sum . right.v
      cur_ps[k] = ps;
    }
    grp.allreduce(d2);
    turn = (d2[0] <= 0) || (d2[1] <= 0);  // :391
  } else {
    double d6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const double2 ps = add2(t1_ps[k], cur_ps[k]);   // :390
      const double2 ps1 = add2(t1_ps[k], cur_lp[k]);  // tree1.p_sum + tree2.left.p (:394)
      const double2 ps2 = add2(t1_rp[k], cur_ps[k]);  // tree1.right.p + tree2.p_sum (:396)
      d6[0] = dot2(d6[0], ps, t1_lv[k]);
      d6[1] = dot2(d6[1], ps, v[k]);
      d6[2] = dot2(d6[2], ps1, t1_lv[k]);
      d6[3] = dot2(d6[3], ps1, cur_lv[k]);
      d6[4] = dot2(d6[4], ps2, t1_rv[k]);
      d6[5] = dot2(d6[5], ps2, v[k]);
      cur_ps[k] = ps;
    }
    grp.allreduce(d6);
    turn = (d6[0] <= 0) || (d6[1] <= 0) || (d6[2] <= 0) || (d6[3] <= 0) || (d6[4] <= 0) || (d6[5] <= 0);  // :391-398
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    cur_lp[k] = t1_lp[k];  // left edge of the merged tree
    cur_lv[k] = t1_lv[k];
  }
  const XF nw = xf_add(XF{ss->wm[lvl], ss->we[lvl]}, cur.w);  // :400
  const XF na = xf_add(XF{ss->am[lvl], ss->ae[lvl]}, cur.a);  // :401-403
  const int t1_pslot = ss->pslot[lvl];
  if (xf_u_less(u, nw, cur.w)) {  // :404-407
    free_slots |= 1u << t1_pslot;
  } else {
    if (cur.pslot != kLeafProp) free_slots |= 1u << cur.pslot;
    cur.pslot = t1_pslot;
    cur.pE = ss->pE[lvl];
    cur.plogp = ss->plogp[lvl];
  }
  cur.w = nw;
  cur.a = na;
  return turn;
}

template <int G, int NP>
__device__ __forceinline__ void dn_push_cur(const Scratch<G, NP>& sc, StackScalars* ss, int lvl, int max_depth,
                                            const double2 (&q)[NP], const double2 (&p)[NP], const double2 (&v)[NP],
                                            const double2 (&cur_lp)[NP], const double2 (&cur_lv)[NP],
                                            const double2 (&cur_ps)[NP], CurTree& cur, unsigned& free_slots) {
  if (cur.pslot == kLeafProp) {
    cur.pslot = __ffs(free_slots) - 1;
    free_slots &= ~(1u << cur.pslot);
#pragma unroll
    for (int k = 0; k < NP; ++k) sc.st(dv_prop(max_depth, cur.pslot), k, q[k]);
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    sc.st(dv_stack(lvl, S_LP), k, cur_lp[k]);
    sc.st(dv_stack(lvl, S_LV), k, cur_lv[k]);
    sc.st(dv_stack(lvl, S_RP), k, p[k]);
    sc.st(dv_stack(lvl, S_RV), k, v[k]);
    sc.st(dv_stack(lvl, S_PS), k, cur_ps[k]);
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

template <int G, int NP>
__device__ __forceinline__ bool dn_extend_top(const Scratch<G, NP>& sc, Group<G>& grp, int tail, int max_depth, int dir,
                                              const double2 (&q)[NP], const double2 (&p)[NP], const double2 (&v)[NP],
                                              const double2 (&cur_lp)[NP], const double2 (&cur_lv)[NP],
                                              const double2 (&cur_ps)[NP], const CurTree& cur, TrajScalars& tr, double u) {
  if (xf_u_less(u, xf_add(tr.Wp, xf_one()), cur.w)) {  // logbern(tree.log_size - self.log_size) nuts.py:321-323
    tr.prop_E = cur.pE;
    tr.prop_logp = cur.plogp;
    if (cur.pslot == kLeafProp) {
#pragma unroll
      for (int k = 0; k < NP; ++k) sc.st(tail + DT_PROPQ, k, q[k]);
    } else {
#pragma unroll
      for (int k = 0; k < NP; ++k) sc.st(tail + DT_PROPQ, k, sc.ld(dv_prop(max_depth, cur.pslot), k));
    }
  }
  tr.Wp = xf_add(tr.Wp, cur.w);    // :325
  tr.Acc = xf_add(tr.Acc, cur.a);  // :326-328
  double d6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const double2 psum = add2(sc.ld(tail + DT_PSUM, k), cur_ps[k]);  // self.p_sum[:] += tree.p_sum (:329)
    sc.st(tail + DT_PSUM, k, psum);
    const double2 oLp = sc.ld(tail + DT_LP, k), oRp = sc.ld(tail + DT_RP, k);
    const double2 voL = sc.ld(tail + DT_LV, k), voR = sc.ld(tail + DT_RV, k);
    const double2 vTl = cur_lv[k], vTr = v[k];
    if (dir > 0) {  // (:300-303, :333-339, with the aliased p_sum)
      const double2 ps1 = add2(psum, cur_lp[k]);
      const double2 ps2 = add2(oRp, cur_ps[k]);
      d6[0] = dot2(d6[0], psum, voL);
      d6[1] = dot2(d6[1], psum, vTr);
      d6[2] = dot2(d6[2], ps1, voL);
      d6[3] = dot2(d6[3], ps1, vTl);
      d6[4] = dot2(d6[4], ps2, voR);
      d6[5] = dot2(d6[5], ps2, vTr);
    } else {        // (:309-312, :333-339)
      const double2 ps1 = add2(cur_ps[k], oLp);
      const double2 ps2 = add2(cur_lp[k], psum);
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

template <int G, int NP, int KIND>
__global__ void __launch_bounds__(dn_block<G>()) dn_advance_kernel(const lmc_dense_args c, size_t vec_off, int n_vecs) {
  constexpr int CPB = dn_block<G>() / G;
  constexpr int VS = G * NP;
  const lmc_sampler_args& a = c.base;
  __shared__ double red_s[CPB * 2 * Group<G>::kWarps * kRedSlots];
  const int gib = threadIdx.x / G;
  const int lane = threadIdx.x - gib * G;
  const int chain = blockIdx.x * CPB + gib;
  if (chain >= a.n_chains) return;
  DnMachine* const M = reinterpret_cast<DnMachine*>(reinterpret_cast<char*>(c.machine) + (size_t)chain * kDnMachineBytes);
  // the caller has not served this chain yet (it batches potential.update over chains): nothing moves, need stays
  if (c.need[chain] & LMC_NEED_HOLD) return;
  DnScalars s = M->s;
  if (s.phase == DPH_DONE) {  // a finished chain may still have asked for its last potential.update: served by now
    if (lane == 0) c.need[chain] = 0;
    return;
  }
  Group<G> grp(lane, red_s + gib * (2 * Group<G>::kWarps * kRedSlots));
  Scratch<G, NP> sc;
  sc.sm = nullptr;
  sc.ws = reinterpret_cast<double2*>(reinterpret_cast<char*>(c.machine) + vec_off) + (size_t)chain * n_vecs * VS;
  sc.n_smem = 0;
  sc.lane = lane;
  StackScalars* const ss = &M->ss;
  const int MD = scratch_depth(a);  // stack sized for both depth caps (lmc_tree.cuh)
  const int tail = dv_tail(MD);
  const int V_Q0 = 0;  // HMC: the transition's start position

  const int D = a.ndim;
  const int ldh = (int)(a.ld >> 1);
  const size_t off = (size_t)chain * a.ld;
  double* const xp = c.x_eval + 2 * off;           // momentum row
  double* const xg = c.x_eval + 2 * off + a.ld;    // gradient row
  const double* const vp = c.v_eval + 2 * off;     // v = M^-1 p
  const double* const vg = c.v_eval + 2 * off + a.ld;  // w = M^-1 g

  double* const ad = a.adapt + (size_t)chain * LMC_ADAPT_STRIDE;
  const long long it = a.iter0 + s.t;
  const bool tune = it < a.n_tune;
  const bool adapt_step = tune && a.adapt_step_size;
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

  // ---- the two "ask for a velocity" phases are pure data movement ---------------------------------------------------
  if (s.phase == DPH_START) {
    // gradient at the start position and p0 arrived (base_hmc.py:142-143): ask for v0 = velocity(p0), w0 = velocity(g0)
    double2 g[NP], p0[NP];
    load_row<G, NP>(c.g_eval + off, lane, ldh, g);
    load_row<G, NP>(c.p0_eval + off, lane, ldh, p0);
    mask_tail<G, NP>(lane, D, g);
    mask_tail<G, NP>(lane, D, p0);
    store_row<G, NP>(xp, lane, ldh, p0);
    store_row<G, NP>(xg, lane, ldh, g);
    s.logp0 = c.logp_eval[chain];
    s.uc = 0;
    s.phase = DPH_START_V;
    if (lane == 0) {
      M->s = s;
      c.need[chain] = LMC_NEED_VEL;
    }
    return;
  }
  if (s.phase == DPH_LEAF_G) {
    // p' = p_half + dt g' (integration.py:116): ask for v' = velocity(p'), w' = velocity(g')
    const double dt = 0.5 * (s.dir > 0 ? s.eps : -s.eps);
    double2 g[NP], p[NP];
    load_row<G, NP>(c.g_eval + off, lane, ldh, g);
    load_row<G, NP>(xp, lane, ldh, p);
    mask_tail<G, NP>(lane, D, g);
#pragma unroll
    for (int k = 0; k < NP; ++k) p[k] = axpy2(p[k], dt, g[k]);
    store_row<G, NP>(xp, lane, ldh, p);
    store_row<G, NP>(xg, lane, ldh, g);
    s.logp_leaf = c.logp_eval[chain];
    s.phase = DPH_LEAF_V;
    if (lane == 0) {
      M->s = s;
      c.need[chain] = LMC_NEED_VEL;
    }
    return;
  }

  // ---- START_V / LEAF_V: the full state z = (q, p, g, v, w) is available ---------------------------------------------
  double2 q[NP], p[NP], g[NP], v[NP], w[NP];
  load_row<G, NP>(c.q_eval + off, lane, ldh, q);
  load_row<G, NP>(xp, lane, ldh, p);
  load_row<G, NP>(xg, lane, ldh, g);
  load_row<G, NP>(vp, lane, ldh, v);
  load_row<G, NP>(vg, lane, ldh, w);
  mask_tail<G, NP>(lane, D, q);
  mask_tail<G, NP>(lane, D, v);
  mask_tail<G, NP>(lane, D, w);
  double k1[1] = {0.0};
#pragma unroll
  for (int k = 0; k < NP; ++k) k1[0] = dot2(k1[0], p[k], v[k]);
  grp.allreduce(k1);

  bool new_doubling = false, trans_end = false, dead = false, diverging = false, reached_max = false;
  double accept_stat = 0.0, stat_a = 0.0, stat_b = 0.0, stat_energy = 0.0, stat_energy_error = 0.0, stat_c = 0.0,
         stat_logp = 0.0;

  if (s.phase == DPH_START_V) {
    s.E0 = 0.5 * k1[0] - s.logp0;  // integration.py:63-65
    if (!isfinite(s.E0)) {
      status |= LMC_STATUS_BAD_INITIAL_ENERGY;
      dead = true;
    } else {
      s.eps = exp(adapt_step ? ad[LMC_ADAPT_LOG_STEP] : ad[LMC_ADAPT_LOG_BAR]);  // step_sizes.py:58-69
      if (a.step_size_override) s.eps = a.step_size_override[chain];             // base_hmc.py:154-155
      if constexpr (KIND == DK_NUTS) {
        s.max_depth = (tune && it < 200) ? a.early_max_treedepth : a.max_treedepth;  // nuts.py:205-208
        s.tr = TrajScalars{xf_zero(), xf_zero(), 0.0, s.E0, s.logp0, 0, 0};
        s.d = 0;
        s.last_dir = 0;
#pragma unroll
        for (int k = 0; k < NP; ++k) {  // _Tree.__init__ (nuts.py:267-282)
          sc.st(tail + DT_LQ, k, q[k]);
          sc.st(tail + DT_LP, k, p[k]);
          sc.st(tail + DT_LG, k, g[k]);
          sc.st(tail + DT_LV, k, v[k]);
          sc.st(tail + DT_LW, k, w[k]);
          sc.st(tail + DT_RQ, k, q[k]);
          sc.st(tail + DT_RP, k, p[k]);
          sc.st(tail + DT_RG, k, g[k]);
          sc.st(tail + DT_RV, k, v[k]);
          sc.st(tail + DT_RW, k, w[k]);
          sc.st(tail + DT_PSUM, k, p[k]);
          sc.st(tail + DT_PROPQ, k, q[k]);
        }
        new_doubling = s.max_depth > 0;
        trans_end = !new_doubling;
        reached_max = trans_end;
      } else {
        s.path_length = next_uniform() * a.path_length;             // hmc.py:141
        s.n_steps = hmc_n_steps(s.path_length, s.eps, a.max_steps);  // :142-143
        s.i = 0;
        s.dir = 1;
#pragma unroll
        for (int k = 0; k < NP; ++k) sc.st(V_Q0, k, q[k]);
      }
    }
  } else {  // DPH_LEAF_V
    const double logp = s.logp_leaf;
    const double E = 0.5 * k1[0] - logp;  // integration.py:118-119
    if constexpr (KIND == DK_NUTS) {
      double2 cur_lp[NP], cur_lv[NP], cur_ps[NP];
      CurTree cur{xf_zero(), xf_zero(), 0.0, 0.0, kLeafProp};
      int fail = 0, lvl = 0;
      ++s.n_leaves;
      if (!leaf_scalars(E, logp, s.E0, a.Emax, s.tr.max_dE, cur)) {
        fail = 1;
      } else {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          cur_lp[k] = cur_ps[k] = p[k];
          cur_lv[k] = v[k];
        }
        unsigned jbits = s.i;
        while (jbits & 1u) {
          if (dn_merge_level<G, NP>(sc, grp, ss, lvl, p, v, cur_lp, cur_lv, cur_ps, cur, s.free_slots, next_uniform())) {
            fail = 2;
            break;
          }
          jbits >>= 1;
          ++lvl;
        }
      }
      if (!fail && s.i + 1 < (1u << s.d)) {
        dn_push_cur<G, NP>(sc, ss, lvl, MD, q, p, v, cur_lp, cur_lv, cur_ps, cur, s.free_slots);
        ++s.i;
      } else {  // the doubling is over (nuts.py:315-340)
        ++s.tr.depth;
        s.tr.n_prop += s.n_leaves;
        if (fail) {
          diverging = (fail == 1);
          trans_end = true;
        } else if (dn_extend_top<G, NP>(sc, grp, tail, MD, s.dir, q, p, v, cur_lp, cur_lv, cur_ps, cur, s.tr,
                                        next_uniform())) {
          trans_end = true;
        } else if (s.d + 1 < s.max_depth) {
          const int base = tail + (s.dir > 0 ? DT_RQ : DT_LQ);  // self.right / self.left = tree.right (:304 / :313)
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            sc.st(base + 0, k, q[k]);
            sc.st(base + 1, k, p[k]);
            sc.st(base + 2, k, g[k]);
            sc.st(base + 3, k, v[k]);
            sc.st(base + 4, k, w[k]);
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

  if constexpr (KIND == DK_NUTS) {
    if (trans_end) {  // _Tree.stats (nuts.py:419-435)
      accept_stat = mean_tree_accept(s.tr);
      stat_a = (double)s.tr.depth;
      stat_b = (double)s.tr.n_prop;
      stat_energy = s.tr.prop_E;
      stat_energy_error = s.tr.prop_E - s.E0;
      stat_c = s.tr.max_dE;
      stat_logp = s.tr.prop_logp;
#pragma unroll
      for (int k = 0; k < NP; ++k) q[k] = sc.ld(tail + DT_PROPQ, k);  // hmc_step.end.q
    }
  }

  int need = 0;
  double* const srow = a.stats + row * LMC_NSTATS;
  double* const trow = a.trace + (size_t)chain * a.trace_chain_stride + (size_t)s.t * a.trace_draw_stride;
  if (trans_end) {
    // ---- close BaseHMC._astep (base_hmc.py:161-190); potential.update is the caller's (LMC_NEED_UPDATE) ------------
    DualAvg da{ad[LMC_ADAPT_LOG_STEP], ad[LMC_ADAPT_LOG_BAR], ad[LMC_ADAPT_HBAR], ad[LMC_ADAPT_COUNT], ad[LMC_ADAPT_MU]};
    if (adapt_step) dual_average_update(da, accept_stat, a.target_accept, a.gamma, a.k, a.t0);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      if (2 * j < D) trow[2 * j] = q[k].x;
      if (2 * j + 1 < D) trow[2 * j + 1] = q[k].y;
    }
    store_row<G, NP>(a.q + off, lane, ldh, q);
    store_row<G, NP>(c.q_eval + off, lane, ldh, q);  // the next transition's gradient is evaluated here
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
    }
    if (tune && a.adapt_mass) need |= LMC_NEED_UPDATE;  // potential.update(end.q, end.q_grad, tune), base_hmc.py:162
    ++s.t;
    if (s.t < a.n_trans) {
      s.phase = DPH_START;
      need |= LMC_NEED_GRAD | LMC_NEED_MOM;
      write_normals(a, chain, s.t, lane, G, c.n_eval + off);
    } else {
      s.phase = DPH_DONE;
    }
  } else if (dead) {
    const double nan = CUDART_NAN;
    for (int tt = s.t; tt < a.n_trans; ++tt) {
      double* tr2 = a.trace + (size_t)chain * a.trace_chain_stride + (size_t)tt * a.trace_draw_stride;
      for (int e = lane; e < D; e += G) tr2[e] = nan;
      if (lane == 0) {
        double* s2 = a.stats + ((size_t)chain * a.n_trans + tt) * LMC_NSTATS;
        for (int n = 0; n < LMC_NSTATS; ++n) s2[n] = nan;
      }
    }
    s.phase = DPH_DONE;
  } else {
    if (new_doubling) {  // nuts.py:213 and the edge the new subtree grows from (:297 / :306)
      s.dir = (next_uniform() < 0.5) ? 1 : -1;
      if (s.last_dir != 0 && s.last_dir != s.dir) {
        const int base = tail + (s.dir > 0 ? DT_RQ : DT_LQ);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          q[k] = sc.ld(base + 0, k);
          p[k] = sc.ld(base + 1, k);
          g[k] = sc.ld(base + 2, k);
          v[k] = sc.ld(base + 3, k);
          w[k] = sc.ld(base + 4, k);
        }
      }
      s.i = 0;
      s.n_leaves = 0;
      s.free_slots = 0xffffffffu;
    }
    // ---- first half of the next leapfrog (integration.py:105-112): p_half = p + dt g, v_half = v + dt w -------------
    const double e = s.dir > 0 ? s.eps : -s.eps;
    const double dt = 0.5 * e;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      p[k] = axpy2(p[k], dt, g[k]);
      const double2 vh = axpy2(v[k], dt, w[k]);
      q[k] = axpy2(q[k], e, vh);
    }
    store_row<G, NP>(xp, lane, ldh, p);
    store_row<G, NP>(c.q_eval + off, lane, ldh, q);
    s.phase = DPH_LEAF_G;
    need = LMC_NEED_GRAD;
  }
  if (lane == 0) {
    M->s = s;
    c.need[chain] = need;
    if (status) atomicOr(a.status + chain, status);
    if (s.phase == DPH_DONE) atomicSub(c.n_running, 1);
  }
}

template <int KIND>
static int dn_launch(const lmc_dense_args& c, size_t vec_off, int n_vecs, int G, int NP) {
  const lmc_sampler_args& a = c.base;
  cudaStream_t st = (cudaStream_t)a.stream;
#define LMC_CASE(gg, np)                                                                                          \
  if (G == gg && NP == np) {                                                                                      \
    constexpr int CPB = dn_block<gg>() / gg;                                                                      \
    dn_advance_kernel<gg, np, KIND><<<(a.n_chains + CPB - 1) / CPB, dn_block<gg>(), 0, st>>>(c, vec_off, n_vecs);  \
    LMC_CUDA(cudaGetLastError());                                                                                 \
    return LMC_OK;                                                                                                \
  }
  LMC_CASE(32, 1) LMC_CASE(64, 1) LMC_CASE(128, 1) LMC_CASE(256, 1) LMC_CASE(256, 2) LMC_CASE(512, 2) LMC_CASE(512, 4)
  LMC_CASE(1024, 4)
#undef LMC_CASE
  return LMC_ERR_UNSUPPORTED;
}

static int dn_check(int kind, const lmc_dense_args* c, int* G, int* NP, size_t* vec_off, int* n_vecs) {
  if (!c || (kind != DK_NUTS && kind != DK_HMC)) return LMC_ERR_BADARG;
  const lmc_sampler_args& a = c->base;
  if (a.abi_version != LMC_ABI_VERSION) return LMC_ERR_BADARG;
  if (a.n_chains < 0 || a.ndim < 1 || a.n_trans < 0 || a.ld < a.ndim || (a.ld & 1)) return LMC_ERR_BADARG;
  if (a.ld > 8192) return LMC_ERR_UNSUPPORTED;
  if (!a.q || !a.adapt || !a.trace || !a.stats || !a.status) return LMC_ERR_BADARG;
  if (!c->q_eval || !c->g_eval || !c->logp_eval || !c->x_eval || !c->v_eval || !c->n_eval || !c->p0_eval || !c->need ||
      !c->machine || !c->n_running)
    return LMC_ERR_BADARG;
  if (((uintptr_t)a.q | (uintptr_t)c->q_eval | (uintptr_t)c->g_eval | (uintptr_t)c->x_eval | (uintptr_t)c->v_eval |
       (uintptr_t)c->n_eval | (uintptr_t)c->p0_eval | (uintptr_t)c->machine) & 15)
    return LMC_ERR_BADARG;
  if (a.rng.mode == LMC_RNG_TAPE) {
    if (!a.rng.normals || !a.rng.uniforms || a.rng.u_stride < 1) return LMC_ERR_BADARG;
  } else if (a.rng.mode == LMC_RNG_PHILOX) {
    if (!a.rng.seeds) return LMC_ERR_BADARG;
  } else {
    return LMC_ERR_BADARG;
  }
  if (kind == DK_NUTS) {
    if (a.max_treedepth < 1 || a.max_treedepth > kMaxDepth) return LMC_ERR_UNSUPPORTED;
    if (a.early_max_treedepth < 0 || a.early_max_treedepth > kMaxDepth) return LMC_ERR_UNSUPPORTED;
  } else if (a.max_steps < 1) {
    return LMC_ERR_BADARG;
  }
  if (a.trace_skip != 0 || a.progress) return LMC_ERR_UNSUPPORTED;  // single-launch host traces: fused kernels only
  if (!dn_pick_shape(a.ndim, G, NP)) return LMC_ERR_UNSUPPORTED;
  *n_vecs = dn_vecs(kind, scratch_depth(a));
  *vec_off = (size_t)a.n_chains * kDnMachineBytes;
  const size_t need = *vec_off + (size_t)a.n_chains * *n_vecs * (size_t)(*G * *NP) * sizeof(double2);
  if ((size_t)c->machine_bytes < need) return LMC_ERR_WORKSPACE;
  return LMC_OK;
}

// ---- y = A x for listed chains: one warp per matrix row, lanes stride along the row with 128-bit loads ---------------
// Each CTA owns ROWS_PER_CTA consecutive rows of one chain's matrix; x (both right-hand sides) is staged in shared
// memory once per CTA.  Traffic: the matrix is read exactly once (8 ndim^2 bytes per chain), x from L2.
constexpr int kMvWarps = 8;
constexpr int kMvRowsPerWarp = 4;
template <int NRHS>
__global__ void __launch_bounds__(kMvWarps * 32) dense_matvec_kernel(const int32_t* __restrict__ idx,
                                                                      const double* __restrict__ A, long long chain_stride,
                                                                      long long lda, int D, long long ld,
                                                                      const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ double2 xs2[];  // [NRHS][ld/2]
  const int chain = idx ? idx[blockIdx.y] : (int)blockIdx.y;
  const int ldh = (int)(ld >> 1);
  const double2* xrow = reinterpret_cast<const double2*>(x + (size_t)chain * NRHS * ld);
  for (int e = threadIdx.x; e < NRHS * ldh; e += blockDim.x) {
    double2 val = xrow[e];
    const int j = e % ldh;
    if (2 * j >= D) val.x = 0.0;
    if (2 * j + 1 >= D) val.y = 0.0;
    xs2[e] = val;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* Ac = A + (size_t)chain * chain_stride;
  const int row0 = (blockIdx.x * kMvWarps + warp) * kMvRowsPerWarp;
  const int Dh = (D + 1) >> 1;  // pairs per row actually holding data (lda even, element D of an odd row is padding)
  if (row0 >= D) return;
  // the warp's rows advance together: kMvRowsPerWarp independent 16-byte loads in flight per lane
  const double2* arow[kMvRowsPerWarp];
#pragma unroll
  for (int r = 0; r < kMvRowsPerWarp; ++r)
    arow[r] = reinterpret_cast<const double2*>(Ac + (size_t)min(row0 + r, D - 1) * lda);
  double acc[kMvRowsPerWarp][NRHS];
#pragma unroll
  for (int r = 0; r < kMvRowsPerWarp; ++r)
#pragma unroll
    for (int n = 0; n < NRHS; ++n) acc[r][n] = 0.0;
  for (int j = lane; j < Dh; j += 32) {
    double2 av[kMvRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kMvRowsPerWarp; ++r) av[r] = __ldcs(arow[r] + j);  // streamed once
    const bool odd_tail = 2 * j + 1 >= D;
#pragma unroll
    for (int n = 0; n < NRHS; ++n) {
      const double2 xv = xs2[n * ldh + j];
#pragma unroll
      for (int r = 0; r < kMvRowsPerWarp; ++r) {
        if (odd_tail) av[r].y = 0.0;
        acc[r][n] = dot2(acc[r][n], av[r], xv);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < kMvRowsPerWarp; ++r)
#pragma unroll
      for (int n = 0; n < NRHS; ++n) acc[r][n] += __shfl_xor_sync(0xffffffffu, acc[r][n], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < kMvRowsPerWarp; ++r) {
      if (row0 + r < D) {
#pragma unroll
        for (int n = 0; n < NRHS; ++n) y[((size_t)chain * NRHS + n) * ld + row0 + r] = acc[r][n];
      }
    }
  }
}

// ---- _WeightedCovariance.add_sample on fg and bg, then cov = raw_fg / (n_fg - 1) -------------------------------------
__global__ void __launch_bounds__(256) dense_cov_update_kernel(const int32_t* __restrict__ idx, int D, long long ld,
                                                               long long lda, const double* __restrict__ x,
                                                               double* mean_fg, double* raw_fg, double* mean_bg,
                                                               double* raw_bg, double* nsamp, double* cov) {
  extern __shared__ double sm[];  // old_fg[D], new_fg[D], old_bg[D], new_bg[D]
  const int chain = idx ? idx[blockIdx.y] : (int)blockIdx.y;
  double* old_fg = sm;
  double* new_fg = sm + D;
  double* old_bg = sm + 2 * D;
  double* new_bg = sm + 3 * D;
  const double n_fg = nsamp[2 * chain] + 1.0, n_bg = nsamp[2 * chain + 1] + 1.0;  // n_samples += 1 (:609)
  const size_t voff = (size_t)chain * ld;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const double xv = x[voff + j];
    const double m1 = mean_fg[voff + j], m2 = mean_bg[voff + j];
    const double o1 = add_rn(xv, -m1), o2 = add_rn(xv, -m2);  // old_diff = x - mean           (:610)
    const double mn1 = add_rn(m1, o1 / n_fg), mn2 = add_rn(m2, o2 / n_bg);  // mean += old_diff / n (:611)
    old_fg[j] = o1;
    old_bg[j] = o2;
    new_fg[j] = add_rn(xv, -mn1);  // new_diff = x - mean (:612)
    new_bg[j] = add_rn(xv, -mn2);
  }
  __syncthreads();
  // raw_cov += weight * new_diff[:, None] * old_diff[None, :]  (:613), rows split over blockIdx.x
  const size_t moff = (size_t)chain * D * lda;
  const int rows_per_block = (D + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(D, r0 + rows_per_block);
  const double denom = n_fg - 1.0;
  for (int i = r0; i < r1; ++i) {
    const double nf = new_fg[i], nb = new_bg[i];
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
      const size_t e = moff + (size_t)i * lda + j;
      const double rf = add_rn(raw_fg[e], mul_rn(nf, old_fg[j]));
      raw_fg[e] = rf;
      raw_bg[e] = add_rn(raw_bg[e], mul_rn(nb, old_bg[j]));
      if (cov) cov[e] = rf / denom;  // np.divide(raw_cov, n_samples - 1) (:620); NULL: no refresh due (update_window)
    }
  }
  // the means and counters are inputs of every block of this chain: they are advanced by a second, tiny launch
}
__global__ void dense_cov_finish_kernel(const int32_t* __restrict__ idx, int D, long long ld,
                                        const double* __restrict__ x, double* mean_fg, double* mean_bg, double* nsamp) {
  const int chain = idx ? idx[blockIdx.x] : (int)blockIdx.x;
  const double n_fg = nsamp[2 * chain] + 1.0, n_bg = nsamp[2 * chain + 1] + 1.0;
  const size_t voff = (size_t)chain * ld;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const double xv = x[voff + j];
    const double m1 = mean_fg[voff + j], m2 = mean_bg[voff + j];
    mean_fg[voff + j] = add_rn(m1, add_rn(xv, -m1) / n_fg);
    mean_bg[voff + j] = add_rn(m2, add_rn(xv, -m2) / n_bg);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    nsamp[2 * chain] = n_fg;
    nsamp[2 * chain + 1] = n_bg;
  }
}

}  // namespace lmc

extern "C" int64_t lmc_dense_state_bytes(int32_t kind, int32_t n_chains, int32_t ndim, int32_t max_treedepth) {
  int G, NP;
  if (n_chains < 0 || (kind != lmc::DK_NUTS && kind != lmc::DK_HMC)) return LMC_ERR_BADARG;
  if (kind == lmc::DK_NUTS && (max_treedepth < 1 || max_treedepth > lmc::kMaxDepth)) return LMC_ERR_UNSUPPORTED;
  if (!lmc::dn_pick_shape(ndim, &G, &NP)) return LMC_ERR_UNSUPPORTED;
  return (int64_t)((size_t)n_chains * lmc::kDnMachineBytes +
                   (size_t)n_chains * lmc::dn_vecs(kind, max_treedepth) * (size_t)(G * NP) * sizeof(double2));
}

extern "C" int lmc_dense_begin(int32_t kind, const lmc_dense_args* c) {
  int G, NP, n_vecs;
  size_t vec_off;
  const int rc = lmc::dn_check(kind, c, &G, &NP, &vec_off, &n_vecs);
  if (rc != LMC_OK) return rc;
  if (c->base.n_chains == 0) return LMC_OK;
  lmc::dn_begin_kernel<<<c->base.n_chains, 128, 0, (cudaStream_t)c->base.stream>>>(*c);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

extern "C" int lmc_dense_advance(int32_t kind, const lmc_dense_args* c) {
  int G, NP, n_vecs;
  size_t vec_off;
  const int rc = lmc::dn_check(kind, c, &G, &NP, &vec_off, &n_vecs);
  if (rc != LMC_OK) return rc;
  if (c->base.n_chains == 0) return LMC_OK;
  return kind == lmc::DK_NUTS ? lmc::dn_launch<lmc::DK_NUTS>(*c, vec_off, n_vecs, G, NP)
                              : lmc::dn_launch<lmc::DK_HMC>(*c, vec_off, n_vecs, G, NP);
}

extern "C" int lmc_dense_matvec(const int32_t* idx, int32_t n_idx, const double* A, int64_t chain_stride, int64_t lda,
                                int32_t ndim, int64_t ld, const double* x, double* y, int32_t nrhs, void* stream) {
  if (!A || !x || !y || ndim < 1 || n_idx < 0 || lda < ndim || (lda & 1) || ld < ndim || (ld & 1)) return LMC_ERR_BADARG;
  if (nrhs != 1 && nrhs != 2) return LMC_ERR_BADARG;
  if (((uintptr_t)A | (uintptr_t)x) & 15) return LMC_ERR_BADARG;
  if (n_idx == 0) return LMC_OK;
  const int rows_per_cta = lmc::kMvWarps * lmc::kMvRowsPerWarp;
  dim3 grid((ndim + rows_per_cta - 1) / rows_per_cta, n_idx);
  const size_t smem = (size_t)nrhs * ld * sizeof(double);
  if (smem > 48 * 1024) return LMC_ERR_UNSUPPORTED;
  if (nrhs == 1)
    lmc::dense_matvec_kernel<1><<<grid, lmc::kMvWarps * 32, smem, (cudaStream_t)stream>>>(idx, A, chain_stride, lda, ndim,
                                                                                          ld, x, y);
  else
    lmc::dense_matvec_kernel<2><<<grid, lmc::kMvWarps * 32, smem, (cudaStream_t)stream>>>(idx, A, chain_stride, lda, ndim,
                                                                                          ld, x, y);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

extern "C" int lmc_dense_cov_update(const int32_t* idx, int32_t n_idx, int32_t ndim, int64_t ld, int64_t lda,
                                    const double* x, double* mean_fg, double* raw_fg, double* mean_bg, double* raw_bg,
                                    double* nsamp, double* cov, void* stream) {
  if (!x || !mean_fg || !raw_fg || !mean_bg || !raw_bg || !nsamp || ndim < 1 || n_idx < 0 || ld < ndim ||
      lda < ndim)
    return LMC_ERR_BADARG;
  if (n_idx == 0) return LMC_OK;
  const size_t smem = (size_t)4 * ndim * sizeof(double);
  if (smem > 48 * 1024) return LMC_ERR_UNSUPPORTED;
  int bx = (ndim + 15) / 16;  // ~16 rows per block
  if (bx > 64) bx = 64;
  lmc::dense_cov_update_kernel<<<dim3(bx, n_idx), 256, smem, (cudaStream_t)stream>>>(idx, ndim, ld, lda, x, mean_fg, raw_fg,
                                                                                    mean_bg, raw_bg, nsamp, cov);
  LMC_CUDA(cudaGetLastError());
  lmc::dense_cov_finish_kernel<<<n_idx, 256, 0, (cudaStream_t)stream>>>(idx, ndim, ld, x, mean_fg, mean_bg, nsamp);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}
