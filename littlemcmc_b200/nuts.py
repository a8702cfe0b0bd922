"""No-U-Turn sampler: mirror of reference nuts.py:31-239.  The tree builder (_Tree, nuts.py:251-435) lives in the CUDA
kernel (csrc/lmc_sampler.cuh)."""
import numpy as np

from . import _lib as L
from .base_hmc import BaseHMC
from .report import SamplerWarning, WarningType

__all__ = ["NUTS"]


class NUTS(BaseHMC):
    name = "nuts"
    default_blocked = True
    generates_stats = True
    stats_dtypes = [{
        "depth": np.int64, "step_size": np.float64, "tune": np.bool_, "mean_tree_accept": np.float64,
        "step_size_bar": np.float64, "tree_size": np.float64, "diverging": np.bool_, "energy_error": np.float64,
        "energy": np.float64, "max_energy_error": np.float64, "model_logp": np.float64,
    }]  # nuts.py:87-101
    _kind = L.KIND_NUTS
    _stat_columns = {"depth": L.STAT_DEPTH, "step_size": L.STAT_STEP_SIZE, "tune": L.STAT_TUNE,
                     "mean_tree_accept": L.STAT_ACCEPT, "step_size_bar": L.STAT_STEP_SIZE_BAR,
                     "tree_size": L.STAT_TREE_SIZE, "diverging": L.STAT_DIVERGING,
                     "energy_error": L.STAT_ENERGY_ERROR, "energy": L.STAT_ENERGY,
                     "max_energy_error": L.STAT_MAX_ENERGY_ERROR, "model_logp": L.STAT_MODEL_LOGP}

    def __init__(self, logp_dlogp_func, model_ndim, scaling=None, is_cov=False, potential=None, target_accept=0.8,
                 Emax=1000, adapt_step_size=True, step_scale=0.25, gamma=0.05, k=0.75, t0=10, step_rand=None,
                 path_length=2.0, max_treedepth=10, early_max_treedepth=8):
        """Arguments and defaults of reference nuts.py:103-121."""
        super().__init__(logp_dlogp_func=logp_dlogp_func, model_ndim=model_ndim, scaling=scaling, is_cov=is_cov,
                         potential=potential, target_accept=target_accept, Emax=Emax,
                         adapt_step_size=adapt_step_size, step_scale=step_scale, gamma=gamma, k=k, t0=t0,
                         step_rand=step_rand)
        self.max_treedepth = max_treedepth
        self.early_max_treedepth = early_max_treedepth
        self.path_length = path_length
        self._reached_max_treedepth = 0

    def _params(self):
        p = super()._params()
        p.update(max_treedepth=self.max_treedepth, early_max_treedepth=self.early_max_treedepth)
        return p

    def _account(self, stats_dev, n_tune_in_block):
        super()._account(stats_dev, n_tune_in_block)
        post = stats_dev[:, n_tune_in_block:, :]
        if post.shape[1]:
            # nuts.py:218-220: the doubling loop ran out without a divergence or a U-turn (flagged by the kernel)
            self._reached_max_treedepth += int(post[:, :, L.STAT_REACHED_MAX_TREEDEPTH].sum().item())

    def warnings(self):
        """reference nuts.py:226-239."""
        warnings = super().warnings()
        n = self._samples_after_tune
        if n > 0 and self._reached_max_treedepth / float(n) > 0.05:
            msg = ("The chain reached the maximum tree depth. Increase max_treedepth, increase target_accept or "
                   "reparameterize.")
            warnings.append(SamplerWarning(WarningType.TREEDEPTH, msg, "warn", None, None, None))
        return warnings
