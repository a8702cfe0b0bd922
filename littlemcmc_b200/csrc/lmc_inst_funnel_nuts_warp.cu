// Instantiates the chunked warp-per-chain NUTS kernel (lmc_sampler_warp.cuh) for Neal's funnel.
#include "lmc_inst_warp.cuh"

namespace lmc {
int run_funnel_nuts_warp(const lmc_sampler_args& a, const Funnel& t) { return dispatch_warp(a, t); }
}  // namespace lmc
