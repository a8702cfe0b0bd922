// Instantiates the sampler kernel for (DiagGaussian, KIND_HMC) over every shape in LMC_SHAPES.
#include "lmc_sampler.cuh"

namespace lmc {
int run_gauss_hmc(const lmc_sampler_args& a, const DiagGaussian& t) { return dispatch_shape<DiagGaussian, KIND_HMC>(a, t); }
}  // namespace lmc
