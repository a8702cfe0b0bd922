// Instantiates the sampler kernel for (DiagGaussian, KIND_NUTS) over every shape in LMC_SHAPES.
#include "lmc_sampler.cuh"

namespace lmc {
int run_gauss_nuts(const lmc_sampler_args& a, const DiagGaussian& t) { return dispatch_shape<DiagGaussian, KIND_NUTS>(a, t); }
}  // namespace lmc
