// Instantiates the chunked CTA-per-chain NUTS kernel (lmc_sampler_cta.cuh) for the diagonal Gaussian target.
#include "lmc_inst_cta.cuh"

namespace lmc {
int run_gauss_nuts_cta(const lmc_sampler_args& a, const DiagGaussian& t) { return dispatch_cta(a, t); }
}  // namespace lmc
