"""Cross-chain convergence diagnostics on the device.

The reference stops at per-run warnings (divergences, tree depth, acceptance rate: base_hmc.py:202-230,
nuts.py:226-239, step_sizes.py:101-121 -- mirrored by `step.warnings()`); with thousands of chains the questions "did
they converge / how many effective draws" need reductions over the whole `[chains, draws, ndim]` trace, which lives on
the GPU (`sample(..., return_device=True)`, `distributed.sample`).  Everything here takes a torch CUDA tensor (or a
NumPy array, which is uploaded) and returns per-dimension device tensors:

* `chain_moments` -- per (chain, segment) mean and centred sum of squares in ONE pass over the trace: the hand-written
  HBM-bound kernel `lmc_chain_moments` (8 bytes per draw element, read once);
* `rhat`          -- split R-hat (Gelman et al., BDA3 / Stan `split_rhat`) from those moments;
* `ess`           -- Stan's effective sample size (Geyer's initial positive + monotone sequence on the multi-chain
  autocorrelation estimate), autocovariances by FFT (cuFFT through torch: a library call, not a kernel of ours).

For an odd number of draws the LAST draw is dropped before splitting.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _as_device_trace(trace, device=None):
    if not torch.is_tensor(trace):
        if device is None:
            if not torch.cuda.is_available():
                raise L.LmcError("littlemcmc_b200.diagnostics runs on CUDA devices only; there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        trace = torch.as_tensor(np.asarray(trace, dtype="d"), device=device)
    if trace.device.type != "cuda":
        raise L.LmcError("littlemcmc_b200.diagnostics needs a CUDA tensor (got %s)" % trace.device)
    if trace.ndim != 3:
        raise ValueError("trace must be [chains, draws, ndim]")
    trace = trace.to(torch.float64)
    if trace.stride(2) != 1:
        trace = trace.contiguous()
    return trace


def chain_moments(trace, n_seg=2):
    """-> (mean [chains, n_seg, ndim], m2 [chains, n_seg, ndim], counts [n_seg]): per-segment mean and sum of squared
    deviations of every chain's draws (segments are contiguous, the last one takes the remainder)."""
    trace = _as_device_trace(trace)
    Cn, T, D = trace.shape
    lib = L.load()
    mean = torch.empty(Cn, n_seg, D, dtype=torch.float64, device=trace.device)
    m2 = torch.empty_like(mean)
    with torch.cuda.device(trace.device):
        L.check(lib.lmc_chain_moments(C.c_void_p(trace.data_ptr()), Cn, T, D, trace.stride(0), trace.stride(1), n_seg,
                                      C.c_void_p(mean.data_ptr()), C.c_void_p(m2.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream(trace.device).cuda_stream)),
                "lmc_chain_moments")
    from . import engine
    engine.LAUNCH_COUNT["kernels"] += 1
    ln = T // n_seg
    counts = torch.full((n_seg,), ln, dtype=torch.float64, device=trace.device)
    counts[-1] = T - ln * (n_seg - 1)
    return mean, m2, counts


def _merge_pairs(mean, m2, counts):
    """Chan's pairwise merge of adjacent segments: [C, S, D] -> [C, S/2, D]."""
    na, nb = counts[0::2], counts[1::2]
    n = na + nb
    ma, mb = mean[:, 0::2], mean[:, 1::2]
    delta = mb - ma
    w = (nb / n)[None, :, None]
    merged_mean = ma + delta * w
    merged_m2 = m2[:, 0::2] + m2[:, 1::2] + delta * delta * (na * nb / n)[None, :, None]
    return merged_mean, merged_m2, n


def split_chain_moments(trace):
    """Moments of the 2 * chains half chains: (mean [2C, D], var [2C, D] with ddof 1, n).  Enough segments are used to
    fill the GPU when there are few chains; they are merged pairwise back to two halves."""
    trace = _as_device_trace(trace)
    Cn, T, D = trace.shape
    half = T // 2
    if half < 2:
        raise ValueError("need at least 4 draws per chain")
    trace = trace[:, :2 * half]
    blocks = Cn * ((D + 127) // 128)
    n_seg = 2
    while blocks * n_seg < 2048 and half // n_seg >= 64:     # keep segments long, but give every SM a few blocks
        n_seg *= 2
    mean, m2, counts = chain_moments(trace, n_seg)
    if (2 * half) % n_seg:                                   # ragged last segment: fall back to exact halves
        mean, m2, counts = chain_moments(trace, 2)
    while mean.shape[1] > 2:
        mean, m2, counts = _merge_pairs(mean, m2, counts)
    return mean.reshape(2 * Cn, D), (m2 / (half - 1)).reshape(2 * Cn, D), half


def rhat(trace):
    """Split R-hat per dimension ([ndim] device tensor)."""
    mean, var, n = split_chain_moments(trace)
    W = var.mean(0)
    B_over_n = mean.var(0, unbiased=True)
    var_plus = W * (n - 1) / n + B_over_n
    return torch.sqrt(var_plus / W)


def ess(trace, split=True):
    """Effective sample size per dimension ([ndim] device tensor), Stan's estimator on (split) chains."""
    trace = _as_device_trace(trace)
    Cn, T, D = trace.shape
    if split:
        half = T // 2
        trace = torch.cat([trace[:, :half], trace[:, half:2 * half]], 0)
    m, n, _ = trace.shape
    if n < 4:
        raise ValueError("need at least 4 draws per (split) chain")
    x = trace - trace.mean(1, keepdim=True)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = torch.fft.rfft(x, n=nfft, dim=1)
    acov = torch.fft.irfft(f.real * f.real + f.imag * f.imag, n=nfft, dim=1)[:, :n] / n      # biased autocovariance
    chain_mean = trace.mean(1)
    mean_var = acov[:, 0].mean(0) * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus = var_plus + chain_mean.var(0, unbiased=True)
    rho = 1.0 - (mean_var[None] - acov.mean(0)) / var_plus[None]                         # [n, D]
    rho[0] = 1.0
    n_pairs = (n - 2) // 2 if n > 4 else 1          # Stan's loop computes pair (s+1, s+2) while s < n - 4, s odd
    n_pairs = max(1, min(n_pairs, n // 2))
    even, odd = rho[0:2 * n_pairs:2], rho[1:2 * n_pairs:2]
    P = even + odd                                                                        # [n_pairs, D]
    ok = P >= 0
    ok[0] = True
    included = torch.cumprod(ok.to(torch.int64), 0).bool()                                # pairs before the first negative
    K = included.sum(0)                                                                   # >= 1
    Pm = torch.cummin(torch.where(included, P, torch.full_like(P, float("inf"))), 0).values
    Pm = torch.where(included, Pm, torch.zeros_like(P))
    # the first excluded pair's even term, if positive (Stan's "improved estimate")
    idx = K.clamp(max=n_pairs - 1)
    fail_even = even.gather(0, idx[None])[0]
    extra = torch.where((K < n_pairs) & (fail_even > 0), fail_even, torch.zeros_like(fail_even))
    tau = -1.0 + 2.0 * Pm.sum(0) + extra
    total = float(m * n)
    return torch.minimum(total / tau, torch.full_like(tau, total * np.log10(total)))


def summary(trace):
    """{mean, sd, rhat, ess}: per-dimension device tensors over all chains and draws."""
    trace = _as_device_trace(trace)
    mean, var, n = split_chain_moments(trace)
    grand = mean.mean(0)
    # pooled variance over all draws of all (half) chains
    m = mean.shape[0]
    ss = (var * (n - 1)).sum(0) + n * ((mean - grand) ** 2).sum(0)
    return {"mean": grand, "sd": torch.sqrt(ss / (m * n - 1)), "rhat": rhat(trace), "ess": ess(trace)}
