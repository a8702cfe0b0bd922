// Instantiates the chunked CTA-per-chain NUTS kernel (lmc_sampler_cta.cuh) for Neal's funnel.
#include "lmc_inst_cta.cuh"

namespace lmc {
int run_funnel_nuts_cta(const lmc_sampler_args& a, const Funnel& t) { return dispatch_cta(a, t); }
}  // namespace lmc
