// Instantiates the sampler kernel for (Funnel, KIND_NUTS) over every shape in LMC_SHAPES.
#include "lmc_sampler.cuh"

namespace lmc {
int run_funnel_nuts(const lmc_sampler_args& a, const Funnel& t) { return dispatch_shape<Funnel, KIND_NUTS>(a, t); }
}  // namespace lmc
