"""Hamiltonian Monte Carlo with a random fixed path length: mirror of reference hmc.py:27-182 (trajectory and Metropolis
accept run in the CUDA kernel)."""
import numpy as np

from . import _lib as L
from .base_hmc import BaseHMC

__all__ = ["HamiltonianMC"]


class HamiltonianMC(BaseHMC):
    name = "hmc"
    generates_stats = True
    stats_dtypes = [{
        "step_size": np.float64, "n_steps": np.int64, "tune": np.bool_, "step_size_bar": np.float64,
        "accept": np.float64, "diverging": np.bool_, "energy_error": np.float64, "energy": np.float64,
        "path_length": np.float64, "accepted": np.bool_, "model_logp": np.float64,
    }]  # hmc.py:36-50
    _kind = L.KIND_HMC
    _stat_columns = {"step_size": L.STAT_STEP_SIZE, "n_steps": L.STAT_DEPTH, "tune": L.STAT_TUNE,
                     "step_size_bar": L.STAT_STEP_SIZE_BAR, "accept": L.STAT_ACCEPT, "diverging": L.STAT_DIVERGING,
                     "energy_error": L.STAT_ENERGY_ERROR, "energy": L.STAT_ENERGY, "path_length": L.STAT_TREE_SIZE,
                     "accepted": L.STAT_MAX_ENERGY_ERROR, "model_logp": L.STAT_MODEL_LOGP}

    def __init__(self, logp_dlogp_func, model_ndim, scaling=None, is_cov=False, potential=None, target_accept=0.8,
                 Emax=1000, adapt_step_size=True, step_scale=0.25, gamma=0.05, k=0.75, t0=10, step_rand=None,
                 path_length=2.0, max_steps=1024):
        """Arguments and defaults of reference hmc.py:52-69."""
        super().__init__(logp_dlogp_func=logp_dlogp_func, model_ndim=model_ndim, scaling=scaling, is_cov=is_cov,
                         potential=potential, target_accept=target_accept, Emax=Emax,
                         adapt_step_size=adapt_step_size, step_scale=step_scale, gamma=gamma, k=k, t0=t0,
                         step_rand=step_rand)
        self.path_length = path_length
        self.max_steps = max_steps

    def _params(self):
        p = super()._params()
        p.update(path_length=self.path_length, max_steps=self.max_steps)
        return p
