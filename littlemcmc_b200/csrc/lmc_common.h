// Host-side error plumbing shared by the translation units of liblmc_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "lmc_b200.h"

namespace lmc {
void set_last_error(const char* what, cudaError_t err);
}

// Evaluate a CUDA runtime call; on failure remember the message and return LMC_ERR_LAUNCH from the caller.
#define LMC_CUDA(expr)                                  \
  do {                                                  \
    cudaError_t lmc_err__ = (expr);                     \
    if (lmc_err__ != cudaSuccess) {                     \
      ::lmc::set_last_error(#expr, lmc_err__);          \
      return LMC_ERR_LAUNCH;                            \
    }                                                   \
  } while (0)
