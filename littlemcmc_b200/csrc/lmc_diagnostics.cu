// lmc_chain_moments: per-(chain, segment, dimension) mean and centred sum of squares of the draws -- the one pass over
// the [chains, draws, ndim] trace that cross-chain diagnostics (split R-hat, between/within variances) are built on.
// Streaming and HBM-bound: every draw element is read exactly once (8 bytes), consecutive threads read consecutive
// dimensions of one draw (coalesced), each thread walks down the draws of its segment with a shifted-data accumulation
// (pivot = the segment's first draw, so the subtraction sum(d^2) - sum(d)^2 / n loses no more than a few bits).
#include <cuda_runtime.h>
#include <stdint.h>

#include "lmc_common.h"

namespace lmc {

constexpr int kMomThreads = 128;
constexpr int kMomUnroll = 4;  // independent loads in flight per thread

// W = 1: one dimension per thread (8-byte loads); W = 2: two adjacent dimensions per thread (16-byte loads; needs even
// strides, an even first column and a 16-byte aligned base: the common case of a contiguous trace with even ndim)
template <int W>
__global__ void __launch_bounds__(kMomThreads) chain_moments_kernel(const double* __restrict__ trace, int n_draws, int D,
                                                                    long long chain_stride, long long draw_stride,
                                                                    int n_seg, double* __restrict__ mean,
                                                                    double* __restrict__ m2) {
  const int i = (blockIdx.x * kMomThreads + threadIdx.x) * W;
  const int seg = blockIdx.y, chain = blockIdx.z;
  if (i >= D) return;
  const int len = n_draws / n_seg;
  const int t0 = seg * len;
  const int t1 = (seg == n_seg - 1) ? n_draws : t0 + len;
  const double* col = trace + (size_t)chain * chain_stride + i;
  const size_t o = ((size_t)chain * n_seg + seg) * D + i;
  if (t1 <= t0) {
#pragma unroll
    for (int w = 0; w < W; ++w) mean[o + w] = m2[o + w] = 0.0;
    return;
  }
  auto load = [&](int t, double (&x)[W]) {
    if constexpr (W == 2) {
      const double2 v = __ldcs(reinterpret_cast<const double2*>(col + (size_t)t * draw_stride));
      x[0] = v.x;
      x[1] = v.y;
    } else {
      x[0] = __ldcs(col + (size_t)t * draw_stride);
    }
  };
  double pivot[W];
  load(t0, pivot);
  double s1[kMomUnroll][W], s2[kMomUnroll][W];
#pragma unroll
  for (int u = 0; u < kMomUnroll; ++u)
#pragma unroll
    for (int w = 0; w < W; ++w) s1[u][w] = s2[u][w] = 0.0;
  int t = t0;
  for (; t + kMomUnroll <= t1; t += kMomUnroll) {
    double x[kMomUnroll][W];
#pragma unroll
    for (int u = 0; u < kMomUnroll; ++u) load(t + u, x[u]);
#pragma unroll
    for (int u = 0; u < kMomUnroll; ++u)
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const double d = x[u][w] - pivot[w];
        s1[u][w] += d;
        s2[u][w] = fma(d, d, s2[u][w]);
      }
  }
  for (; t < t1; ++t) {
    double x[W];
    load(t, x);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const double d = x[w] - pivot[w];
      s1[0][w] += d;
      s2[0][w] = fma(d, d, s2[0][w]);
    }
  }
  const double n = (double)(t1 - t0);
#pragma unroll
  for (int w = 0; w < W; ++w) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int u = 0; u < kMomUnroll; ++u) {
      a += s1[u][w];
      b += s2[u][w];
    }
    mean[o + w] = pivot[w] + a / n;
    m2[o + w] = fmax(b - a * a / n, 0.0);
  }
}

}  // namespace lmc

extern "C" int lmc_chain_moments(const double* trace, int32_t n_chains, int32_t n_draws, int32_t ndim,
                                 int64_t chain_stride, int64_t draw_stride, int32_t n_seg, double* mean, double* m2,
                                 void* stream) {
  if (!trace || !mean || !m2 || n_chains < 0 || n_draws < 0 || ndim < 1 || n_seg < 1 || draw_stride < ndim)
    return LMC_ERR_BADARG;
  if (n_chains == 0) return LMC_OK;
  if (n_seg > 65535 || n_chains > 65535) return LMC_ERR_UNSUPPORTED;
  const bool vec = (ndim % 2 == 0) && (chain_stride % 2 == 0) && (draw_stride % 2 == 0) && (((uintptr_t)trace & 15) == 0);
  if (vec) {
    dim3 grid((ndim / 2 + lmc::kMomThreads - 1) / lmc::kMomThreads, n_seg, n_chains);
    lmc::chain_moments_kernel<2><<<grid, lmc::kMomThreads, 0, (cudaStream_t)stream>>>(trace, n_draws, ndim, chain_stride,
                                                                                     draw_stride, n_seg, mean, m2);
  } else {
    dim3 grid((ndim + lmc::kMomThreads - 1) / lmc::kMomThreads, n_seg, n_chains);
    lmc::chain_moments_kernel<1><<<grid, lmc::kMomThreads, 0, (cudaStream_t)stream>>>(trace, n_draws, ndim, chain_stride,
                                                                                     draw_stride, n_seg, mean, m2);
  }
  LMC_CUDA(cudaGetLastError());
  return LMC_OK;
}
