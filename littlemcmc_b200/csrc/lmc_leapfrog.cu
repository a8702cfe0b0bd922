// Batched single-step entry points: compute_state / leapfrog step with a built-in target, the two leapfrog halves
// around an external (torch) gradient, and the Philox tape dump.  These mirror CpuLeapfrogIntegrator one call at
// a time (reference integration.py:52-121) and are what GpuLeapfrogIntegrator binds; the throughput path is
// lmc_sampler.cu, which keeps the same device functions in registers across a whole transition.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <initializer_list>

#include "lmc_common.h"
#include "lmc_device.cuh"

namespace lmc {

static thread_local char g_last_error[256] = "";
void set_last_error(const char* what, cudaError_t err) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", cudaGetErrorString(err), what);
}

template <int G>
__host__ __device__ constexpr int step_block() { return G >= 64 ? G : 128; }

// MODE 0: compute_state, MODE 1: full leapfrog step
template <class Target, int G, int NP, int MODE>
__global__ void __launch_bounds__(step_block<G>()) step_kernel(const Target tgt, int n_chains, int D, long long ld,
                                                               const double* __restrict__ eps, const double* q_in,
                                                               const double* p_in, const double* g_in,
                                                               const double* __restrict__ var_in, long long var_stride,
                                                               double* q_out, double* p_out, double* v_out,
                                                               double* g_out, double* energy, double* logp) {
  constexpr int CPB = step_block<G>() / G;
  __shared__ double red_s[CPB * 2 * Group<G>::kWarps * kRedSlots];
  const int gib = threadIdx.x / G;
  const int lane = threadIdx.x - gib * G;
  const int chain = blockIdx.x * CPB + gib;
  if (chain >= n_chains) return;  // G == 32: per-warp exit; G >= 64: CPB == 1, whole block exits together
  Group<G> grp(lane, red_s + gib * (2 * Group<G>::kWarps * kRedSlots));
  const int ldh = (int)(ld >> 1);
  const size_t off = (size_t)chain * ld;
  double2 q[NP], p[NP], g[NP], var[NP];
  load_row<G, NP>(q_in + off, lane, ldh, q);
  load_row<G, NP>(p_in + off, lane, ldh, p);
  load_row<G, NP>(var_in + (size_t)chain * var_stride, lane, ldh, var);
  mask_tail<G, NP>(lane, D, q);
  mask_tail<G, NP>(lane, D, p);
  mask_tail<G, NP>(lane, D, var);
  double E, lp;
  if constexpr (MODE == 0) {
    eval_energy<false>(tgt, grp, D, ldh, q, p, g, var, 0.0, E, lp);
  } else {
    load_row<G, NP>(g_in + off, lane, ldh, g);
    mask_tail<G, NP>(lane, D, g);
    leapfrog(tgt, grp, D, ldh, eps[chain], q, p, g, var, E, lp);
    store_row<G, NP>(q_out + off, lane, ldh, q);
    store_row<G, NP>(p_out + off, lane, ldh, p);
  }
  double2 v[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) v[k] = mul2(var[k], p[k]);
  store_row<G, NP>(v_out + off, lane, ldh, v);
  store_row<G, NP>(g_out + off, lane, ldh, g);
  if (lane == 0) {
    energy[chain] = E;
    logp[chain] = lp;
  }
}

static bool pick_step_shape(int ndim, int* G, int* NP) {
  const int pairs = (ndim + 1) / 2;
  static const int table[][2] = {{32, 1}, {32, 2}, {32, 4}, {64, 4}, {128, 4}, {256, 4}, {512, 4}, {1024, 4}};
  for (auto& s : table)
    if (s[0] * s[1] >= pairs) { *G = s[0]; *NP = s[1]; return true; }
  return false;
}

template <class Target, int MODE>
static int launch_step(const Target& tgt, int n_chains, int ndim, long long ld, const double* eps, const double* q,
                       const double* p, const double* g, const double* var, long long var_stride, double* q_out,
                       double* p_out, double* v_out, double* g_out, double* energy, double* logp, cudaStream_t st) {
  int G, NP;
  if (!pick_step_shape(ndim, &G, &NP)) return LMC_ERR_UNSUPPORTED;
#define LMC_CASE(gg, np)                                                                                          \
  if (G == gg && NP == np) {                                                                                      \
    constexpr int CPB = step_block<gg>() / gg;                                                                    \
    step_kernel<Target, gg, np, MODE><<<(n_chains + CPB - 1) / CPB, step_block<gg>(), 0, st>>>(                  \
        tgt, n_chains, ndim, ld, eps, q, p, g, var, var_stride, q_out, p_out, v_out, g_out, energy, logp);        \
    LMC_CUDA(cudaGetLastError());                                                                                 \
    return LMC_OK;                                                                                                \
  }
  LMC_CASE(32, 1) LMC_CASE(32, 2) LMC_CASE(32, 4) LMC_CASE(64, 4) LMC_CASE(128, 4) LMC_CASE(256, 4)
  LMC_CASE(512, 4) LMC_CASE(1024, 4)
#undef LMC_CASE
  return LMC_ERR_UNSUPPORTED;
}

static int check_rows(int n_chains, int ndim, long long ld, std::initializer_list<const void*> ptrs) {
  if (n_chains < 0 || ndim < 1 || ld < ndim || (ld & 1)) return LMC_ERR_BADARG;
  for (const void* p : ptrs)
    if (!p || ((uintptr_t)p & 15)) return LMC_ERR_BADARG;
  return LMC_OK;
}

// ---- the two halves around an external gradient: elementwise, HBM-bound, one thread per pair ---------------------
__global__ void half1_kernel(int n_chains, int D, int ldh, const double* __restrict__ eps,
                             const int* __restrict__ active, double2* q, double2* p, const double2* __restrict__ g,
                             const double2* __restrict__ var, long long var_stride_h) {
  const int chain = blockIdx.x;
  if (active && !active[chain]) return;
  const double e = eps[chain], dt = 0.5 * e;
  const size_t off = (size_t)chain * ldh;
  for (int j = threadIdx.x; j < ldh; j += blockDim.x) {
    double2 pj = axpy2(p[off + j], dt, g[off + j]);                      // integration.py:108
    double2 qj = axpy2(q[off + j], e, mul2(var[chain * var_stride_h + j], pj));  // :111-112
    if (2 * j >= D) { pj.x = 0.0; qj.x = 0.0; }
    if (2 * j + 1 >= D) { pj.y = 0.0; qj.y = 0.0; }
    p[off + j] = pj;
    q[off + j] = qj;
  }
}

// one warp-multiple block per chain; kinetic energy reduced in the block
__global__ void __launch_bounds__(256) half2_kernel(int n_chains, int D, int ldh, const double* __restrict__ eps,
                                                    const int* __restrict__ active, double2* p, double2* v,
                                                    const double2* __restrict__ g, const double* __restrict__ logp,
                                                    const double2* __restrict__ var, long long var_stride_h,
                                                    double* energy) {
  __shared__ double part[8];
  const int chain = blockIdx.x;
  if (active && !active[chain]) return;
  const double dt = 0.5 * eps[chain];
  const size_t off = (size_t)chain * ldh;
  double acc = 0.0;
  for (int j = threadIdx.x; j < ldh; j += blockDim.x) {
    double2 pj = axpy2(p[off + j], dt, g[off + j]);  // integration.py:116
    if (2 * j >= D) pj.x = 0.0;
    if (2 * j + 1 >= D) pj.y = 0.0;
    const double2 vj = mul2(var[chain * var_stride_h + j], pj);  // :118 (velocity_energy)
    acc = dot2(acc, pj, vj);
    p[off + j] = pj;
    v[off + j] = vj;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    energy[chain] = 0.5 * s - logp[chain];  // :119
  }
}

__global__ void rng_fill_kernel(const uint64_t* __restrict__ seeds, int n_chains, int D, long long iter0, int n_trans,
                                long long u_stride, double* normals, double* uniforms) {
  const int chain = blockIdx.x;
  const uint64_t seed = seeds[chain];
  const int pairs = (D + 1) / 2;
  const long long per_t = pairs + u_stride;
  for (long long w = (long long)blockIdx.y * blockDim.x + threadIdx.x; w < per_t * n_trans;
       w += (long long)gridDim.y * blockDim.x) {
    const int t = (int)(w / per_t);
    const long long r = w - (long long)t * per_t;
    const size_t row = (size_t)chain * n_trans + t;
    if (r < pairs) {
      const double2 n = philox_normal_pair(seed, iter0 + t, (uint32_t)r);
      normals[row * D + 2 * r] = n.x;
      if (2 * r + 1 < D) normals[row * D + 2 * r + 1] = n.y;
    } else {
      uniforms[row * u_stride + (r - pairs)] = philox_uniform(seed, iter0 + t, (uint32_t)(r - pairs));
    }
  }
}

template <int MODE>
static int step_entry(const lmc_target* target, int32_t n_chains, int32_t ndim, int64_t ld, const double* eps,
                      const double* q, const double* p, const double* g, const double* var, int64_t var_stride,
                      double* q_out, double* p_out, double* v_out, double* g_out, double* energy, double* logp,
                      void* stream) {
  if (!target || !energy || !logp) return LMC_ERR_BADARG;
  int rc = check_rows(n_chains, ndim, ld, {q, p, var, v_out, g_out});
  if (rc != LMC_OK) return rc;
  if (MODE == 1) {
    rc = check_rows(n_chains, ndim, ld, {g, q_out, p_out});
    if (rc != LMC_OK || !eps) return LMC_ERR_BADARG;
  }
  if (var_stride != 0 && var_stride != ld) return LMC_ERR_BADARG;
  if (n_chains == 0) return LMC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (target->kind == LMC_TARGET_DIAG_GAUSSIAN) {
    if (!target->tau || ((uintptr_t)target->tau & 15)) return LMC_ERR_BADARG;
    DiagGaussian t{reinterpret_cast<const double2*>(target->tau)};
    return launch_step<DiagGaussian, MODE>(t, n_chains, ndim, ld, eps, q, p, g, var, var_stride, q_out, p_out, v_out,
                                           g_out, energy, logp, st);
  }
  if (target->kind == LMC_TARGET_FUNNEL) {
    Funnel t{1.0 / (target->v_scale * target->v_scale), 0.5 * (double)(ndim - 1)};
    return launch_step<Funnel, MODE>(t, n_chains, ndim, ld, eps, q, p, g, var, var_stride, q_out, p_out, v_out, g_out,
                                     energy, logp, st);
  }
  return LMC_ERR_UNSUPPORTED;
}

}  // namespace lmc

extern "C" int lmc_abi_version(void) { return LMC_ABI_VERSION; }
extern "C" const char* lmc_last_error(void) { return lmc::g_last_error; }

extern "C" int lmc_compute_state(const lmc_target* target, int32_t n_chains, int32_t ndim, int64_t ld, const double* q,
                                 const double* p, const double* var, int64_t var_stride, double* v, double* g,
                                 double* energy, double* logp, void* stream) {
  return lmc::step_entry<0>(target, n_chains, ndim, ld, nullptr, q, p, nullptr, var, var_stride, nullptr, nullptr, v, g,
                            energy, logp, stream);
}

extern "C" int lmc_leapfrog_step(const lmc_target* target, int32_t n_chains, int32_t ndim, int64_t ld, const double* eps,
                                 const double* q, const double* p, const double* g, const double* var,
                                 int64_t var_stride, double* q_out, double* p_out, double* v_out, double* g_out,
                                 double* energy, double* logp, void* stream) {
  return lmc::step_entry<1>(target, n_chains, ndim, ld, eps, q, p, g, var, var_stride, q_out, p_out, v_out, g_out,
                            energy, logp, stream);
}

extern "C" int lmc_leapfrog_half1(int32_t n_chains, int32_t ndim, int64_t ld, const double* eps, const int32_t* active,
                                  double* q, double* p, const double* g, const double* var, int64_t var_stride,
                                  void* stream) {
  int rc = lmc::check_rows(n_chains, ndim, ld, {q, p, g, var});
  if (rc != LMC_OK || !eps) return LMC_ERR_BADARG;
  if (var_stride != 0 && var_stride != ld) return LMC_ERR_BADARG;
  if (n_chains == 0) return LMC_OK;
  const int ldh = (int)(ld / 2);
  const int threads1 = ldh >= 256 ? 256 : (ldh >= 128 ? 128 : (ldh >= 64 ? 64 : 32));
  lmc::half1_kernel<<<n_chains, threads1, 0, (cudaStream_t)stream>>>(
      n_chains, ndim, ldh, eps, active, reinterpret_cast<double2*>(q), reinterpret_cast<double2*>(p),
      reinterpret_cast<const double2*>(g), reinterpret_cast<const double2*>(var), var_stride / 2);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

extern "C" int lmc_leapfrog_half2(int32_t n_chains, int32_t ndim, int64_t ld, const double* eps, const int32_t* active,
                                  double* p, double* v, const double* g_new, const double* logp, const double* var,
                                  int64_t var_stride, double* energy, void* stream) {
  int rc = lmc::check_rows(n_chains, ndim, ld, {p, v, g_new, var});
  if (rc != LMC_OK || !eps || !logp || !energy) return LMC_ERR_BADARG;
  if (var_stride != 0 && var_stride != ld) return LMC_ERR_BADARG;
  if (n_chains == 0) return LMC_OK;
  const int ldh = (int)(ld / 2);
  const int threads = ldh >= 256 ? 256 : (ldh >= 128 ? 128 : (ldh >= 64 ? 64 : 32));
  lmc::half2_kernel<<<n_chains, threads, 0, (cudaStream_t)stream>>>(
      n_chains, ndim, ldh, eps, active, reinterpret_cast<double2*>(p), reinterpret_cast<double2*>(v),
      reinterpret_cast<const double2*>(g_new), logp, reinterpret_cast<const double2*>(var), var_stride / 2, energy);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

extern "C" int lmc_rng_fill(const uint64_t* seeds, int32_t n_chains, int32_t ndim, int64_t iter0, int32_t n_trans,
                            int64_t u_stride, double* normals, double* uniforms, void* stream) {
  if (!seeds || !normals || !uniforms || n_chains < 0 || ndim < 1 || n_trans < 0 || u_stride < 0) return LMC_ERR_BADARG;
  if (n_chains == 0 || n_trans == 0) return LMC_OK;
  const long long work = ((long long)(ndim + 1) / 2 + u_stride) * n_trans;
  dim3 grid(n_chains, (unsigned)((work + 255) / 256 > 1024 ? 1024 : (work + 255) / 256));
  lmc::rng_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seeds, n_chains, ndim, iter0, n_trans, u_stride, normals,
                                                               uniforms);
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}

extern "C" int lmc_memcpy2d_d2h(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes,
                                int64_t height, void* stream) {
  if (!dst || !src || width_bytes < 0 || height < 0 || dpitch < width_bytes || spitch < width_bytes)
    return LMC_ERR_BADARG;
  if (width_bytes == 0 || height == 0) return LMC_OK;
  LMC_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes, (size_t)height,
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return LMC_OK;
}
