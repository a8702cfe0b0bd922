// Instantiates the sampler kernel for (Funnel, KIND_HMC) over every shape in LMC_SHAPES.
#include "lmc_sampler.cuh"

namespace lmc {
int run_funnel_hmc(const lmc_sampler_args& a, const Funnel& t) { return dispatch_shape<Funnel, KIND_HMC>(a, t); }
}  // namespace lmc
