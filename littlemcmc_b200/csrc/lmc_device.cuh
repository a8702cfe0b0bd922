// Device-side building blocks shared by the sampler and leapfrog kernels (sm_100a).
//
// Layout / mapping used everywhere:
//   * one chain is owned by a thread GROUP of G threads (G = 32: one warp, no block barriers at all;
//     G >= 64: one whole CTA, cross-warp reductions through shared memory + __syncthreads);
//   * a length-D vector is held as NP double2 "pairs" per thread: pair index j = lane + k*G (k < NP) covers
//     elements (2j, 2j+1), so a warp touches 32 consecutive 16-byte words = 512 contiguous bytes per access
//     (coalesced 128-bit LDG/STG);  elements >= D are kept at exactly 0 in registers;
//   * each thread only ever reads back the scratch words it wrote itself, so scratch vectors (tree stack,
//     trajectory edges) need no fences or barriers;
//   * every scalar of the tree state is replicated in all threads of the group; control flow is uniform
//     across a group because every decision is taken on all-reduced values (bitwise identical on all lanes).
//
// Arithmetic contract: element-wise state updates use explicitly un-fused IEEE operations (__dmul_rn /
// __dadd_rn), i.e. the same two roundings NumPy's separate multiply/add ufuncs make in the reference
// (integration.py:108-116), so positions/momenta are bit-identical to the reference given the same step
// size.  Dot products use FMA accumulation + a butterfly; like BLAS ddot in the reference their summation
// order is unspecified, which is where the (<= few ulp) differences come from.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#else  // NVRTC (user targets compiled at run time): built-in declarations only
#define CUDART_INF __longlong_as_double(0x7ff0000000000000LL)
#define CUDART_NAN __longlong_as_double(0xfff8000000000000LL)
#endif

#include "lmc_b200.h"

namespace lmc {

constexpr int kMaxDepth = 16;  // stack scalars are sized for max_treedepth <= 16
constexpr int kRedSlots = 8;   // doubles per warp in the cross-warp reduction buffer

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
// a + b*c with two roundings (NumPy: `a + b * c`)
__device__ __forceinline__ double axpy_rn(double a, double b, double c) { return __dadd_rn(a, __dmul_rn(b, c)); }

__device__ __forceinline__ double2 mul2(double2 a, double2 b) { return make_double2(mul_rn(a.x, b.x), mul_rn(a.y, b.y)); }
__device__ __forceinline__ double2 add2(double2 a, double2 b) { return make_double2(add_rn(a.x, b.x), add_rn(a.y, b.y)); }
__device__ __forceinline__ double2 axpy2(double2 a, double s, double2 c) {
  return make_double2(axpy_rn(a.x, s, c.x), axpy_rn(a.y, s, c.y));
}
__device__ __forceinline__ double dot2(double acc, double2 a, double2 b) { return fma(a.y, b.y, fma(a.x, b.x, acc)); }

// ---- scalar helpers (reference math.py:21-40, numpy's logaddexp) ------------------------------------------
__device__ __forceinline__ double logaddexp(double x, double y) {
  if (x == y) return x + 0.693147180559945309417232121458176568;  // npy_logaddexp: also handles equal infinities
  const double d = x - y;
  if (d > 0) return x + log1p(exp(-d));
  if (d <= 0) return y + log1p(exp(d));
  return d;  // NaN
}
__device__ __forceinline__ double log1mexp(double x) {  // math.py:28-35
  return x < 0.683 ? log(-expm1(-x)) : log1p(-exp(-x));
}

// ---- extended-range positive reals for the tree weights -------------------------------------------------------------
// The reference keeps subtree sizes in the log domain (log_size, log_weighted_accept_sum) and pays two logaddexp and
// one log(u) per merge (nuts.py:400-404).  The same real numbers are kept here as value = m * 2^e with a double
// mantissa and an int exponent: the range is unbounded for any Emax, a merge is two scaled additions and one
// multiply-compare, and the only transcendental left is one exp per leaf.  Results agree with the log-domain ones to
// a few ulp; a decision can only differ when u lies within ~1e-15 (relative) of its threshold.
struct XF {
  double m;  // >= 0
  int e;
};
// Fast path / wide path.  A leaf weight is exp(-dE) with |dE| < Emax (1000 by default): almost always |dE| is small, the
// weight is an ordinary double and every XF in the tree has e == 0, so add / compare are one DADD / one DMUL + compare.
// Only when |x| > kXfFast does a value carry a nonzero exponent; operations that meet one take the (out-of-line) wide
// path.  Invariant: m == 0, or 2^-600 < m < 2^300 -- sums of <= 2^16 fast-path weights (each within 2^+-254) and
// their squares stay normal doubles.
constexpr double kXfFast = 176.0;  // exp(+-176) = 2^+-253.9
__device__ __forceinline__ XF xf_zero() { return XF{0.0, 0}; }
__device__ __forceinline__ XF xf_one() { return XF{1.0, 0}; }
__device__ __forceinline__ double xf_scale(double m, int d) {  // m * 2^d, d <= 0, flushing to 0 far below
  return d < -1000 ? 0.0 : m * __longlong_as_double((long long)(1023 + d) << 52);
}
// exp(x) for finite x of any magnitude: x = n ln2 + r, |r| <= ln2/2 (Cody-Waite, fdlibm's split of ln2)
static __device__ __noinline__ XF xf_exp_wide(double x) {
  const double n = rint(x * 1.44269504088896338700e+00);
  double r = fma(-n, 6.93147180369123816490e-01, x);
  r = fma(-n, 1.90821492927058770002e-10, r);
  return XF{exp(r), (int)n};
}
__device__ __forceinline__ XF xf_exp(double x) {
  if (fabs(x) <= kXfFast) return XF{exp(x), 0};
  return xf_exp_wide(x);
}
static __device__ __noinline__ XF xf_add_wide(XF a, XF b) {
  if (a.m == 0.0) return b;
  if (b.m == 0.0) return a;
  const int e = a.e > b.e ? a.e : b.e;
  return XF{xf_scale(a.m, a.e - e) + xf_scale(b.m, b.e - e), e};
}
__device__ __forceinline__ XF xf_add(XF a, XF b) {
  if (a.e == b.e) return XF{a.m + b.m, a.e};
  return xf_add_wide(a, b);
}
__device__ __forceinline__ XF xf_sqr(XF a) { return XF{a.m * a.m, 2 * a.e}; }
// u * a < b   (u in [0,1), a, b >= 0)
static __device__ __noinline__ bool xf_u_less_wide(double u, XF a, XF b) {
  if (b.m == 0.0) return false;
  if (a.m == 0.0) return true;
  const int e = a.e > b.e ? a.e : b.e;
  return u * xf_scale(a.m, a.e - e) < xf_scale(b.m, b.e - e);
}
__device__ __forceinline__ bool xf_u_less(double u, XF a, XF b) {
  if (a.e == b.e) return u * a.m < b.m;
  return xf_u_less_wide(u, a, b);
}
__device__ __forceinline__ double xf_value(XF a) { return a.m == 0.0 ? 0.0 : scalbn(a.m, a.e); }
__device__ __forceinline__ double xf_ratio(XF a, XF b) {  // a / b, b > 0
  return a.m == 0.0 ? 0.0 : scalbn(a.m / b.m, a.e - b.e);
}

// ---- Philox4x32-10 (Salmon et al. 2011), written out so the stream is a documented function of
//      (key = per-chain seed, transition index, counter) and can be dumped by lmc_rng_fill -------------------
struct u32x4 { uint32_t x, y, z, w; };

__device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
constexpr uint32_t kTagUniform = 0x554E4946u;  // "UNIF"
constexpr uint32_t kTagNormal = 0x4E4F524Du;   // "NORM"

// 52 random bits -> (0,1), never 0 or 1: (m + 0.5) * 2^-52
__device__ __forceinline__ double u52(uint32_t lo, uint32_t hi) {
  const uint64_t m = (((uint64_t)hi << 32) | lo) >> 12;
  return ((double)m + 0.5) * 2.220446049250313080847263336181640625e-16;
}
// k-th uniform of transition `it` of the chain keyed by `seed`
__device__ __forceinline__ double philox_uniform(uint64_t seed, int64_t it, uint32_t k) {
  const u32x4 r = philox4x32_10(u32x4{k, (uint32_t)it, (uint32_t)((uint64_t)it >> 32), kTagUniform},
                                (uint32_t)seed, (uint32_t)(seed >> 32));
  return u52(r.x, r.y);
}
// Code that runs once per transition (not once per leapfrog) is kept OUT OF LINE: the sampler kernel's hot loop has to
// share a 32 KB L1.5 instruction cache with it (profiles/r01d_ncu_full.md: 18% of warp samples wait for instructions),
// and an inlined double-precision log / sincospi / pow per register pair is several KB each.
#ifdef LMC_INLINE_COLD
#define LMC_COLD __device__ __forceinline__
#else
#define LMC_COLD static __device__ __noinline__
#endif
// one Philox block of the uniform stream, out of line (the chunked kernel refills 32 uniforms at a time)
LMC_COLD double philox_uniform_cold(uint64_t seed, int64_t it, uint32_t k) { return philox_uniform(seed, it, k); }
// IEEE 1/sqrt(x) and a/b (two correctly rounded operations each, as NumPy computes them): ~45 instructions apiece
LMC_COLD double inv_sqrt_cold(double x) { return 1.0 / sqrt(x); }
LMC_COLD double div_cold(double a, double b) { return a / b; }
// standard normals for elements (2j, 2j+1) of the momentum draw of transition `it` (Box-Muller)
LMC_COLD double2 philox_normal_pair(uint64_t seed, int64_t it, uint32_t j) {
  const u32x4 r = philox4x32_10(u32x4{j, (uint32_t)it, (uint32_t)((uint64_t)it >> 32), kTagNormal},
                                (uint32_t)seed, (uint32_t)(seed >> 32));
  const double rad = sqrt(-2.0 * log(u52(r.x, r.y)));
  double s, c;
  sincospi(2.0 * u52(r.z, r.w), &s, &c);
  return make_double2(rad * c, rad * s);
}

// 1/sqrt(x) of NP pairs, IEEE sqrt and divide as in inv_sqrt_cold: ONE out-of-line copy of the code, a rolled loop with
// the two chains of a pair interleaved (code size matters more than the last bit of instruction-level parallelism here)
template <int NP>
static __device__ __noinline__ void inv_sqrt_pairs(const double2 (&v)[NP], double2 (&out)[NP]) {
#pragma unroll 1
  for (int k = 0; k < NP; ++k) out[k] = make_double2(1.0 / sqrt(v[k].x), 1.0 / sqrt(v[k].y));
}
// r / w for NP pairs (IEEE divide), same shape
template <int NP>
static __device__ __noinline__ void div_pairs(const double2 (&r)[NP], double w, double2 (&out)[NP]) {
#pragma unroll 1
  for (int k = 0; k < NP; ++k) out[k] = make_double2(r[k].x / w, r[k].y / w);
}

// ---- thread group owning one chain ---------------------------------------------------------------------------
template <int G>
struct Group {
  static constexpr int kWarps = G / 32;
  static_assert(G % 32 == 0, "a group is a whole number of warps");
  int lane;      // 0 .. G-1
  double* red;   // shared: [2][kWarps][kRedSlots] (unused when kWarps == 1)
  int phase;

  __device__ __forceinline__ Group(int lane_, double* red_) : lane(lane_), red(red_), phase(0) {}

  // Sum v[0..N) over the group; every thread receives bitwise-identical totals.
  // (xor butterfly: both partners compute a+b and b+a, which are equal, at every stage.)
  template <int N>
  __device__ __forceinline__ void allreduce(double (&v)[N]) {
    static_assert(N <= kRedSlots, "grow kRedSlots");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int n = 0; n < N; ++n) v[n] += __shfl_xor_sync(0xffffffffu, v[n], o);
    }
    if constexpr (kWarps > 1) {
      // two alternating buffers: a thread can be at most one reduction ahead of the slowest one because of
      // the barrier, so buffer (phase) is never overwritten while still being read.
      double* buf = red + phase * (kWarps * kRedSlots);
      phase ^= 1;
      if ((lane & 31) == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) buf[(lane >> 5) * kRedSlots + n] = v[n];
      }
      __syncthreads();
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double s = buf[n];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) s += buf[w * kRedSlots + n];
        v[n] = s;
      }
    }
  }
};

// barrier over the threads of one group (G == 32: the warp; G >= 64: the whole CTA)
template <int G>
__device__ __forceinline__ void group_barrier() {
  if constexpr (G == 32) __syncwarp(); else __syncthreads();
}

// ---- built-in target densities -----------------------------------------------------------------------------------
// Protocol (all methods are called by every thread of the group with its own NP pairs):
//   kPre            number of sums needed BEFORE the gradient can be formed (0 or 2)
//   pre(...)        per-thread partials of those sums
//   grad(...)       g from q (and the reduced pre-sums); returns this thread's partial of the logp sum
//   finish(...)     logp from the reduced logp sum and the pre-sums
// Optional: a specialisation of StageTraits (below) names per-dimension parameter vectors the chunked warp kernel may
// copy to shared memory once per launch, and the type that evaluates the target from that copy.
struct DiagGaussian {
  static constexpr int kPre = 0;
  const double2* tau;  // [ldh] pairs, padding = 0

  template <int G, int NP>
  __device__ __forceinline__ void pre(int, int, const double2 (&)[NP], double (&)[2]) const {}

  template <int G, int NP>
  __device__ __forceinline__ double grad(int lane, int D, int ldh, const double2 (&q)[NP], double2 (&g)[NP],
                                         const double (&)[2]) const {
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      const double2 t = (j < ldh) ? __ldg(tau + j) : make_double2(0.0, 0.0);
      // g = -(tau * q): one rounding, the negation is exact.
      // Elements >= D: tau = 0 (zero-padded buffer / no load) and q = 0 (kept at exactly 0 in registers), so g = -0.0
      // there, which leaves p and the sums unchanged -- no masking needed.
      g[k] = make_double2(-mul_rn(t.x, q[k].x), -mul_rn(t.y, q[k].y));
      part = dot2(part, q[k], g[k]);
    }
    return part;
  }
  __device__ __forceinline__ double finish(double sum, const double (&)[2]) const { return 0.5 * sum; }
};

// DiagGaussian with tau read from a zero-padded copy in shared memory ([NP][32] pairs, pointer already offset by the
// lane): what the chunked warp kernel evaluates when a resident chain's shared memory has room for it (StageTraits)
struct DiagGaussianStaged {
  static constexpr int kPre = 0;
  const double2* tau_s;

  template <int G, int NP>
  __device__ __forceinline__ void pre(int, int, const double2 (&)[NP], double (&)[2]) const {}

  template <int G, int NP>
  __device__ __forceinline__ double grad(int, int, int, const double2 (&q)[NP], double2 (&g)[NP], const double (&)[2]) const {
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const double2 t = tau_s[k * G];
      g[k] = make_double2(-mul_rn(t.x, q[k].x), -mul_rn(t.y, q[k].y));  // as DiagGaussian::grad
      part = dot2(part, q[k], g[k]);
    }
    return part;
  }
  __device__ __forceinline__ double finish(double sum, const double (&)[2]) const { return 0.5 * sum; }
};

template <bool C, class A, class B>
struct Cond { using type = A; };
template <class A, class B>
struct Cond<false, A, B> { using type = B; };

// Per-dimension parameter vectors of a target that a kernel may stage in shared memory once per launch: kVecs of them,
// `make` fills the copy (every lane its own NP pairs) and returns the target that reads it.  Default: nothing staged.
template <class T>
struct StageTraits {
  static constexpr int kVecs = 0;
  using Staged = T;
  template <int NP>
  static __device__ __forceinline__ Staged make(const T& t, double2*, int, int) { return t; }
};
template <>
struct StageTraits<DiagGaussian> {
  static constexpr int kVecs = 1;
  using Staged = DiagGaussianStaged;
  template <int NP>
  static __device__ __forceinline__ Staged make(const DiagGaussian& t, double2* s_lane, int lane, int ldh) {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * 32;
      s_lane[k * 32] = (j < ldh) ? __ldg(t.tau + j) : make_double2(0.0, 0.0);
    }
    return DiagGaussianStaged{s_lane};
  }
};

struct Funnel {
  static constexpr int kPre = 2;  // [0] = S = sum_{i>=1} q_i^2, [1] = v = q_0 (only its owner contributes)
  double inv_s2;                  // 1 / v_scale^2
  double half_nm1;                // (D - 1) / 2

  template <int G, int NP>
  __device__ __forceinline__ void pre(int lane, int D, const double2 (&q)[NP], double (&out)[2]) const {
    double S = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      if (j > 0) S = fma(q[k].x, q[k].x, S);  // element 2j >= 2; padding elements are 0
      S = fma(q[k].y, q[k].y, S);
    }
    out[0] = S;
    out[1] = (lane == 0) ? q[0].x : 0.0;
  }

  template <int G, int NP>
  __device__ __forceinline__ double grad(int lane, int D, int, const double2 (&q)[NP], double2 (&g)[NP],
                                         const double (&pre)[2]) const {
    const double ev = exp(-pre[1]);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      g[k] = make_double2(-mul_rn(ev, q[k].x), -mul_rn(ev, q[k].y));
      if (2 * j >= D) g[k].x = 0.0;
      if (2 * j + 1 >= D) g[k].y = 0.0;
    }
    if (lane == 0) {
      const double hs = mul_rn(mul_rn(0.5, ev), pre[0]);
      g[0].x = add_rn(add_rn(-mul_rn(pre[1], inv_s2), hs), -half_nm1);
    }
    return 0.0;
  }
  __device__ __forceinline__ double finish(double, const double (&pre)[2]) const {
    const double v = pre[1];
    const double ev = exp(-v);
    const double hs = mul_rn(mul_rn(0.5, ev), pre[0]);
    const double a = -mul_rn(mul_rn(mul_rn(0.5, v), v), inv_s2);
    return add_rn(add_rn(a, -hs), -mul_rn(half_nm1, v));
  }
};

// Evaluate the target at q, optionally finish the leapfrog's second half-kick, and return the energy:
//   g = dlogp(q);  if KICK: p += dt*g;  v = var*p;  K = 0.5 p.v;  E = K - logp      (integration.py:62-65,115-119)
template <bool KICK, class Target, int G, int NP>
__device__ __forceinline__ void eval_energy(const Target& tgt, Group<G>& grp, int D, int ldh, const double2 (&q)[NP],
                                            double2 (&p)[NP], double2 (&g)[NP], const double2 (&var)[NP], double dt,
                                            double& energy, double& logp) {
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
    if constexpr (KICK) p[k] = axpy2(p[k], dt, g[k]);
    acc[0] = dot2(acc[0], p[k], mul2(var[k], p[k]));
  }
  grp.allreduce(acc);
  logp = tgt.finish(acc[1], pre);
  energy = 0.5 * acc[0] - logp;
}

// One full leapfrog step in registers (integration.py:100-121); eps may be negative.
template <class Target, int G, int NP>
__device__ __forceinline__ void leapfrog(const Target& tgt, Group<G>& grp, int D, int ldh, double eps, double2 (&q)[NP],
                                         double2 (&p)[NP], double2 (&g)[NP], const double2 (&var)[NP], double& energy,
                                         double& logp) {
  const double dt = 0.5 * eps;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    p[k] = axpy2(p[k], dt, g[k]);                 // p_half = p + dt * q_grad
    q[k] = axpy2(q[k], eps, mul2(var[k], p[k]));  // q_new  = q + eps * (var * p_half)
  }
  eval_energy<true>(tgt, grp, D, ldh, q, p, g, var, dt, energy, logp);
}

// Load / store this thread's NP pairs of a user-facing row (row stride ld, pairs j >= ldh do not exist).
template <int G, int NP>
__device__ __forceinline__ void load_row(const double* row, int lane, int ldh, double2 (&x)[NP]) {
  const double2* r = reinterpret_cast<const double2*>(row);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    x[k] = (j < ldh) ? r[j] : make_double2(0.0, 0.0);
  }
}
// 128-bit load that bypasses L1 (ld.global.cg): for rows another SM may have written since this SM last read them
__device__ __forceinline__ double2 ldcg2(const double2* p) {
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
template <int G, int NP>
__device__ __forceinline__ void load_row_cg(const double* row, int lane, int ldh, double2 (&x)[NP]) {
  const double2* r = reinterpret_cast<const double2*>(row);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    x[k] = (j < ldh) ? ldcg2(r + j) : make_double2(0.0, 0.0);
  }
}
template <int G, int NP>
__device__ __forceinline__ void store_row(double* row, int lane, int ldh, const double2 (&x)[NP]) {
  double2* r = reinterpret_cast<double2*>(row);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    if (j < ldh) r[j] = x[k];
  }
}
// zero elements >= D (rows handed in by the caller may carry anything in their padding)
template <int G, int NP>
__device__ __forceinline__ void mask_tail(int lane, int D, double2 (&x)[NP]) {
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int j = lane + k * G;
    if (2 * j >= D) x[k].x = 0.0;
    if (2 * j + 1 >= D) x[k].y = 0.0;
  }
}

}  // namespace lmc
