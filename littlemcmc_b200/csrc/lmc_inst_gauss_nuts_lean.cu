// Instantiates the lean NUTS sampler kernel (lmc_sampler_lean.cuh) for the built-in targets.
#include "lmc_sampler_lean.cuh"

namespace lmc {

// (threads per chain, pairs per thread, min resident CTAs per SM)
#ifndef LMC_LEAN_MC_128_4
#define LMC_LEAN_MC_128_4 4
#endif
#define LMC_LEAN_SHAPES(X) X(64, 8, 5) X(64, 4, 8) X(128, 4, LMC_LEAN_MC_128_4) X(128, 2, 6) X(256, 2, 3)

bool pick_lean_shape(int ndim, int group, int* G, int* NP) {
  const int pairs = (ndim + 1) / 2;
#define LMC_X(g, np, mc) \
  if (group == g && g * np >= pairs) { *G = g; *NP = np; return true; }
  LMC_X(64, 4, 8) LMC_X(64, 8, 5) LMC_X(128, 2, 6) LMC_X(128, 4, 4) LMC_X(256, 2, 3)
#undef LMC_X
  return false;
}

template <class Target>
static int dispatch_lean(const lmc_sampler_args& a, const Target& t, int group) {
  int G = 0, NP = 0;
  if (!pick_lean_shape(a.ndim, group, &G, &NP)) return LMC_ERR_UNSUPPORTED;
#define LMC_X(g, np, mc) \
  if (G == g && NP == np) return launch_lean<Target, g, np, mc>(a, t);
  LMC_LEAN_SHAPES(LMC_X)
#undef LMC_X
  return LMC_ERR_UNSUPPORTED;
}

int run_gauss_nuts_lean(const lmc_sampler_args& a, const DiagGaussian& t, int group) { return dispatch_lean(a, t, group); }
int run_funnel_nuts_lean(const lmc_sampler_args& a, const Funnel& t, int group) { return dispatch_lean(a, t, group); }

}  // namespace lmc
