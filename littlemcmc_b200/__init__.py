"""littlemcmc_b200: B200-native HMC / NUTS hot path behind littlemcmc's sampler API.

Import surface of the reference package (littlemcmc/__init__.py:19-29) plus `targets` (densities with fused
kernels) and `distributed` (chain sharding over the GPUs of a node).
"""
__version__ = "0.1.0"

from . import diagnostics, distributed, interop, targets  # noqa: F401
from .hmc import HamiltonianMC  # noqa: F401
from .nuts import NUTS  # noqa: F401
from .quadpotential import (  # noqa: F401
    QuadPotentialDiag,
    QuadPotentialDiagAdapt,
    QuadPotentialFull,
    QuadPotentialFullAdapt,
    QuadPotentialFullInv,
    quad_potential,
)
from .sampling import init_nuts, sample  # noqa: F401
