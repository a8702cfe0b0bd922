// Shapes of the chunked CTA-per-chain NUTS kernel (lmc_sampler_cta.cuh): 128 threads per chain,
// (pairs per thread, leaves per chunk, resident CTAs per SM -> register cap 65536 / (128 * MINB)).
#pragma once
#include "lmc_sampler_cta.cuh"

namespace lmc {

#define LMC_CTA_SHAPES(X) X(4, 4, 3) X(4, 2, 4) X(4, 2, 3) X(2, 4, 4) X(2, 2, 4)

// tune_group 2: the library's shape for this ndim / chunk; 203 / 204: force 3 / 4 resident CTAs per SM
inline bool pick_cta_shape(int ndim, int chunk, int group, int* NP, int* B, int* MINB) {
  const int pairs = (ndim + 1) / 2;
  const int np = pairs <= 256 ? 2 : pairs <= 512 ? 4 : 0;
  if (!np) return false;
  const int b = chunk ? chunk : 4;
  int minb = group >= 200 ? group - 200 : 0;
#define LMC_X(n, bb, mb) if (np == n && b == bb && (minb == 0 || minb == mb)) { *NP = n; *B = bb; *MINB = mb; return true; }
  LMC_CTA_SHAPES(LMC_X)
#undef LMC_X
  return false;
}

template <class Target>
int dispatch_cta(const lmc_sampler_args& a, const Target& t) {
  int NP = 0, B = 0, MINB = 0;
  if (!pick_cta_shape(a.ndim, a.tune_chunk, a.tune_group, &NP, &B, &MINB)) return LMC_ERR_UNSUPPORTED;
#define LMC_X(n, bb, mb) if (NP == n && B == bb && MINB == mb) return launch_cta<Target, 128, n, bb, mb>(a, t);
  LMC_CTA_SHAPES(LMC_X)
#undef LMC_X
  return LMC_ERR_UNSUPPORTED;
}

}  // namespace lmc
