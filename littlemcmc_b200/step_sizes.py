"""Dual-averaging step-size adaptation: host descriptor of the per-chain device state.

The arithmetic of reference step_sizes.py:71-92 runs inside the sampler kernels (epilogue of every tuning
transition); this class holds the parameters and exposes the reference's attribute names (`_log_step`, `_log_bar`,
`_hbar`, `_count`, `_mu`) as views of the LAST chain's device state, which is what the reference object holds after
sequential sampling.
"""
import numpy as np

from . import _lib as L
from .report import SamplerWarning, WarningType


class DualAverageAdaptation:
    def __init__(self, initial_step, target, gamma, k, t0):
        self._initial_step, self._target, self._k, self._t0, self._gamma = initial_step, target, k, t0, gamma
        self._chains = None
        self._tuned_stats = []
        self.reset()

    def _bind(self, chains):
        self._chains = chains
        self.reset()

    def reset(self):
        """reference step_sizes.py:49-56, applied to every chain."""
        self._tuned_stats = []
        if self._chains is not None:
            self._chains.reset_step_adapt(self._initial_step)

    def _scalar(self, idx):
        if self._chains is None:
            init = {L.ADAPT_LOG_STEP: np.log(self._initial_step), L.ADAPT_LOG_BAR: np.log(self._initial_step),
                    L.ADAPT_HBAR: 0.0, L.ADAPT_COUNT: 1.0, L.ADAPT_MU: np.log(10 * self._initial_step)}
            return float(init[idx])
        return float(self._chains.adapt[-1, idx].item())

    _log_step = property(lambda self: self._scalar(L.ADAPT_LOG_STEP))
    _log_bar = property(lambda self: self._scalar(L.ADAPT_LOG_BAR))
    _hbar = property(lambda self: self._scalar(L.ADAPT_HBAR))
    _mu = property(lambda self: self._scalar(L.ADAPT_MU))
    _count = property(lambda self: int(self._scalar(L.ADAPT_COUNT)))

    def current(self, tune):
        """reference step_sizes.py:58-69 (last chain)."""
        return float(np.exp(self._log_step if tune else self._log_bar))

    def current_all(self, tune):
        """Step size of every chain (device tensor [C])."""
        col = L.ADAPT_LOG_STEP if tune else L.ADAPT_LOG_BAR
        return self._chains.adapt[:, col].exp()

    def stats(self):
        return {"step_size": float(np.exp(self._log_step)), "step_size_bar": float(np.exp(self._log_bar))}

    def warnings(self):
        """reference step_sizes.py:101-121 (post-hoc host diagnostic on the recorded acceptance statistics)."""
        from scipy import stats as sps
        accept = np.asarray(self._tuned_stats, dtype="d")
        if accept.size == 0:
            return []
        mean_accept = float(np.mean(accept))
        n_bound = min(100, accept.size)
        lower, upper = sps.beta(mean_accept * n_bound + 1, (1 - mean_accept) * n_bound + 1).interval(0.95)
        if self._target < lower or self._target > upper:
            msg = ("The acceptance probability does not match the target. It is %s, but should be close to %s. "
                   "Try to increase the number of tuning steps." % (mean_accept, self._target))
            return [SamplerWarning(WarningType.BAD_ACCEPTANCE, msg, "warn", None, None,
                                   {"target": self._target, "actual": mean_accept})]
        return []
