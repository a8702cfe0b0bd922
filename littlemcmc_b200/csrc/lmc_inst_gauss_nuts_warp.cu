// Instantiates the chunked warp-per-chain NUTS kernel (lmc_sampler_warp.cuh) for the diagonal Gaussian target.
#include "lmc_inst_warp.cuh"

namespace lmc {
int run_gauss_nuts_warp(const lmc_sampler_args& a, const DiagGaussian& t) { return dispatch_warp(a, t); }
}  // namespace lmc
