// Shapes of the chunked warp-per-chain NUTS kernel (lmc_sampler_warp.cuh): (pairs per lane, leaves per chunk).
#pragma once
#include "lmc_sampler_warp.cuh"

namespace lmc {

// (NP, B, min resident CTAs per SM -> register cap 65536 / (32 * MINB))
// (the cap follows what shared memory lets be resident anyway: 10 chains per SM at (1, 8), 7 at (2, 8) -- with 128 / 168
// registers ptxas rematerialised and re-loaded; at 168 / 216: cfg4's bulk +6%, cfg2 +1%, a chain alone +3%)
#define LMC_WARP_SHAPES(X) X(1, 4, 16) X(1, 8, 10) X(1, 16, 12) X(2, 4, 12) X(2, 8, 7) X(2, 16, 8) X(4, 4, 8) X(4, 8, 8)

// default chunk: as long as the memory of a resident chain allows a useful number of chains per SM
inline int default_chunk(int np) { return np >= 4 ? 4 : 8; }

inline bool pick_warp_shape(int ndim, int chunk, int* NP, int* B) {
  const int pairs = (ndim + 1) / 2;
  const int np = pairs <= 32 ? 1 : pairs <= 64 ? 2 : pairs <= 128 ? 4 : 0;
  if (!np) return false;
  const int b = chunk ? chunk : default_chunk(np);
#define LMC_X(n, bb, mb) if (np == n && b == bb) { *NP = n; *B = bb; return true; }
  LMC_WARP_SHAPES(LMC_X)
#undef LMC_X
  return false;
}

template <class Target>
int dispatch_warp(const lmc_sampler_args& a, const Target& t) {
  int NP = 0, B = 0;
  if (!pick_warp_shape(a.ndim, a.tune_chunk, &NP, &B)) return LMC_ERR_UNSUPPORTED;
#define LMC_X(n, bb, mb) if (NP == n && B == bb) return launch_warp<Target, n, bb, 1, mb>(a, t);
  LMC_WARP_SHAPES(LMC_X)
#undef LMC_X
  return LMC_ERR_UNSUPPORTED;
}

}  // namespace lmc
