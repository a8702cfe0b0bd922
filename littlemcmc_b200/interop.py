"""Interop helpers for the `(trace, stats)` pair `sample()` returns.

The reference ships no exporter; its cookbook shows a snippet `arviz_from_littlemcmc(trace, stats)`
(docs/tutorials/framework_cookbook.rst:199-205).  `to_arviz_dict` produces the two dictionaries that snippet feeds to
ArviZ -- `posterior = {"x": [chains, draws, ndim]}` and `sample_stats = {name: [chains, draws]}` -- from host arrays or
device tensors; `arviz_from_littlemcmc` builds the `InferenceData` when ArviZ is installed.
"""
import numpy as np


def _host(x):
    try:
        import torch
        if torch.is_tensor(x):
            return x.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    return np.asarray(x)


def to_arviz_dict(trace, stats, var_name="x"):
    """-> (posterior, sample_stats): dicts of host arrays in ArviZ's (chain, draw, ...) layout.  `stats` entries are
    squeezed from `[chains, draws, 1]` to `[chains, draws]`, as the reference's snippet does."""
    tr = _host(trace)
    if tr.ndim != 3:
        raise ValueError("trace must be [chains, draws, ndim]")
    sample_stats = {}
    for k, v in stats.items():
        v = _host(v)
        if v.shape[:2] != tr.shape[:2]:
            raise ValueError("statistic %r has shape %s, expected (%d, %d, 1)" % (k, v.shape, tr.shape[0], tr.shape[1]))
        sample_stats[k] = v.reshape(tr.shape[0], tr.shape[1])
    return {var_name: tr}, sample_stats


def arviz_from_littlemcmc(trace, stats, var_name="x"):
    """The reference cookbook's helper (framework_cookbook.rst:199-205).  Needs ArviZ."""
    try:
        import arviz as az
    except ImportError as e:
        raise ImportError("arviz_from_littlemcmc needs the `arviz` package; `to_arviz_dict` works without it") from e
    posterior, sample_stats = to_arviz_dict(trace, stats, var_name)
    return az.InferenceData(posterior=az.dict_to_dataset(posterior), sample_stats=az.dict_to_dataset(sample_stats))
