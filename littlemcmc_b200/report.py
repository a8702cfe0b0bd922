"""Sampler warning records (mirror of reference littlemcmc/report.py:20-37)."""
import enum
from collections import namedtuple

SamplerWarning = namedtuple("SamplerWarning", "kind, message, level, step, exec_info, extra")


class WarningType(enum.Enum):
    DIVERGENCE = 1          # a divergence after tuning
    TUNING_DIVERGENCE = 2   # a divergence during tuning
    DIVERGENCES = 3         # summary of all post-tuning divergences
    TREEDEPTH = 4           # max_treedepth was reached too often
    BAD_PARAMS = 5          # problematic sampler parameters
    CONVERGENCE = 6         # convergence diagnostics are bad
    BAD_ACCEPTANCE = 7      # mean acceptance far from target_accept
    BAD_ENERGY = 8          # energy diagnostics are bad
