"""Sampling driver: mirror of reference sampling.py (`sample` :35-222, `init_nuts` :524-605).

The reference runs chains one after another (or one OS process per chain) and, inside each, a Python loop over draws
(sampling.py:370, :507).  Here all chains advance together on the GPU and the draw loop is inside the kernel: the
driver issues a handful of launches (blocks of transitions) and copies finished blocks of the trace to pinned host
memory on a side stream while the next block runs.

Seeding follows the reference exactly on the host (np.random.seed(random_seed); one randint(2**30) per chain,
sampling.py:131-134; init_nuts reseeds with the first one and draws the single jittered start, :574-584).  The
per-chain seeds then key per-chain Philox streams on the device, so results do not depend on how chains are
distributed over GPUs.
"""
import logging
import os
from collections import namedtuple
from collections.abc import Iterable

import numpy as np
import torch

from . import _lib as L
from .hmc import HamiltonianMC  # noqa: F401
from .nuts import NUTS
from .quadpotential import QuadPotentialDiagAdapt

_log = logging.getLogger("littlemcmc_b200")

# what the reference's per-draw callback receives (parallel_sampling.py:374; sampling.py:303-308)
Draw = namedtuple("Draw", ["chain", "is_last", "draw_idx", "tuning", "stats", "point", "warnings"])


def _resolve_seeds(random_seed, chains):
    """reference sampling.py:131-138."""
    if random_seed is None or isinstance(random_seed, (int, np.integer)):
        if random_seed is not None:
            np.random.seed(int(random_seed))
        return [int(np.random.randint(2 ** 30)) for _ in range(chains)]
    if isinstance(random_seed, Iterable):
        seeds = [int(s) for s in random_seed]
        if len(seeds) < chains:
            raise ValueError("need one random seed per chain")
        return seeds[:chains]
    raise TypeError("Invalid value for `random_seed`. Must be tuple, list or int")


def sample(logp_dlogp_func, model_ndim, draws=1000, tune=1000, step=None, init="auto", chains=None, cores=None,
           start=None, progressbar=True, random_seed=None, discard_tuned_samples=True, chain_idx=0, callback=None,
           mp_ctx=None, pickle_backend="pickle", device=None, block=None, return_device=False, host_write="copy",
           stats_as="dict", single_launch=True, _timing=None, **kwargs):
    """Draw samples with the given step method; signature and return value of reference `sample` (sampling.py:35-222).

    Returns ``(trace, stats)``: ``trace`` float64 ``[chains, draws, model_ndim]``; ``stats`` a dict of arrays
    ``[chains, draws, 1]`` with the dtypes of ``step.stats_dtypes[0]``.

    Differences a caller can observe: ``cores``, ``mp_ctx``, ``pickle_backend`` and ``progressbar`` are accepted and
    ignored (chains are a tensor dimension, there are no worker processes); ``start`` may also be ``[chains, ndim]``.
    ``callback(trace=..., draw=Draw(...))`` is the reference's per-draw hook (sampling.py:303-308): it is called for
    every (draw, chain) after the block of transitions containing the draw has finished, with ``trace`` the host array
    being filled (``[chains, draws kept so far.., ndim]``) and ``draw.stats`` the one-element list of that draw's
    statistics dict; a callback makes the driver wait for every block (no copy/compute overlap).  ``KeyboardInterrupt``
    stops the run and returns the transitions finished so far, like the reference's sequential path (:470-478).
    Extra keywords: ``device`` (CUDA device, default current), ``block`` (transitions per launch, default: sized so a
    trace block is <= 128 MiB), ``return_device`` (keep results as torch tensors on the GPU and skip the host copy),
    ``host_write``: how kept draws reach the pinned host trace -- ``"copy"`` (default): double-buffered device blocks
    copied out by the copy engine on a side stream while the next block samples (measured: 66 ms for 3.3 GB of draws
    next to 54 ms of sampling); ``"direct"``: the sampler kernel stores each draw straight into the mapped pinned
    buffer (no staging; measured slower, 74 ms, because the stores back-pressure the sampling groups; fused targets
    only); ``single_launch`` (default True): with a fused or user-source density and a host trace the whole run is ONE
    kernel launch -- the kernel skips discarded tuning draws itself and reports finished blocks of kept draws through
    device counters, which the copy engine follows (False: one launch per block of transitions, as for callbacks);
    ``stats_as="tensor"`` (with ``return_device``): the statistics as the ONE ``[chains, draws, 13]`` device
    tensor the kernels write (columns = ``step._stat_columns``) instead of a dict of views of it.
    """
    import time as _time
    _t = [_time.perf_counter()]

    def _mark(name):
        if _timing is not None:
            now = _time.perf_counter()
            _timing[name] = _timing.get(name, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now
    if cores is None:
        cores = min(4, os.cpu_count() or 1)
    if chains is None:
        chains = max(2, cores)                                                      # sampling.py:124-128
    seeds = _resolve_seeds(random_seed, chains)
    if draws == 0:
        _log.warning("Tuning was enabled throughout the whole trace.")
    elif draws < 500:
        _log.warning("Only %d samples in chain.", draws)

    if step is None or start is None:                                               # sampling.py:148-159
        start_, step_ = init_nuts(logp_dlogp_func=logp_dlogp_func, model_ndim=model_ndim, init=init,
                                  random_seed=seeds, **kwargs)
        step = step_ if step is None else step
        start = start_ if start is None else start
    start = np.asarray(start, dtype="d")
    if start.ndim == 1:
        start = np.broadcast_to(start, (chains, model_ndim))                        # one start for all chains (:163-164)

    T = int(tune) + int(draws)
    ch = step._bind(chains, device=device, seeds=seeds)
    step.tune = bool(tune)                                                          # sampling.py:503
    step.reset_tuning()                                                             # :504-505, for every chain
    step.iter_count = 0                                                             # :508-509
    ch.status.zero_()
    ch.set_position(start)
    dev, D = ch.device, int(model_ndim)

    keep_from = int(tune) if discard_tuned_samples else 0
    n_keep = T - keep_from
    block_kept_on_device = block
    if block is None:
        block = max(1, min(T, (1 << 27) // max(1, chains * D * 8)))
        # kept draws that stay on the device need no staging: one launch for all of them (every launch boundary costs
        # a drain -- the last transitions of a block run on a partly idle GPU); 2^31 scheduler tickets per launch
        block_kept_on_device = max(1, min(T, ((1 << 31) - 1) // max(1, chains))) if callback is None else block
    if return_device:
        trace_out = torch.empty(chains, n_keep, D, dtype=torch.float64, device=dev)
        host_trace = None
    else:
        trace_out = None
        host_trace = torch.empty(chains, n_keep, D, dtype=torch.float64, pin_memory=True)
    if host_write not in ("direct", "copy"):
        raise ValueError("host_write must be 'direct' or 'copy'")
    direct = host_write == "direct" and not return_device and step._fused_target() is not None
    _mark("setup+alloc")
    compute = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(device=dev)
    bufs, copy_done = [None, None], [None, None]      # double-buffered device blocks of the trace
    stats_blocks = []
    done, blk = 0, 0
    interrupted = False
    # ---- host trace, fused (or user-source) density, no per-draw hook: ONE launch for the whole run ----------------------
    # The kernel keeps the kept draws in a device trace, skips the discarded tuning draws itself (trace_skip) and counts
    # finished chains per block of kept draws (progress); the copy engine ships every block to the pinned host trace as
    # soon as its counter is full, while the same launch keeps sampling.  No launch boundaries: no drains (the last
    # transitions of a block otherwise run on a partly idle GPU, ~25 times per run at the headline size).
    single = (not return_device and not direct and callback is None and n_keep > 0 and T > 0
              and step._fused_target() is not None and step._step_rand is None
              and not getattr(step.potential, "_dense", False) and single_launch
              and chains * n_keep * D * 8 <= 0.5 * torch.cuda.mem_get_info(dev)[0]
              and chains * T < (1 << 31))
    if single:
        pb = max(1, min(block, n_keep))
        n_blocks = (n_keep + pb - 1) // pb
        trace_dev = torch.empty(chains, n_keep, D, dtype=torch.float64, device=dev)
        progress = torch.zeros(n_blocks, dtype=torch.int32, device=dev)
        progress_host = torch.zeros(n_blocks, dtype=torch.int32).pin_memory()
        _, st = step._run(T, int(tune), trace=trace_dev, trace_skip=keep_from, progress=progress, progress_block=pb)
        stats_blocks.append(st)
        _mark("enqueue")
        lib = L.load()
        shipped, idle = 0, 0
        with torch.cuda.stream(copy_stream):
            while shipped < n_blocks:
                progress_host.copy_(progress, non_blocking=True)
                copy_stream.synchronize()
                counts = progress_host.numpy()
                ready = shipped
                while ready < n_blocks and int(counts[ready]) == chains:
                    ready += 1
                if ready == shipped:
                    idle += 1
                    if idle > 4:
                        _time.sleep(5e-5)                  # nothing new: do not hammer the driver while the kernel works
                    continue
                idle = 0
                lo, hi = shipped * pb, min(n_keep, ready * pb)   # consecutive finished blocks go out as one copy
                src, dst = trace_dev[:, lo:hi], host_trace[:, lo:hi]
                L.check(lib.lmc_memcpy2d_d2h(dst.data_ptr(), dst.stride(0) * 8, src.data_ptr(), src.stride(0) * 8,
                                             (hi - lo) * D * 8, chains, copy_stream.cuda_stream), "lmc_memcpy2d_d2h")
                shipped = ready
        done = T
    try:
        while done < T:
            kept = done >= keep_from
            n = min(block_kept_on_device if (kept and return_device) else block, T - done)
            if done < keep_from < done + n:
                n = keep_from - done                       # a block never straddles the discard boundary
            if kept and return_device:
                tr_view = trace_out[:, done - keep_from:done - keep_from + n]
                _, st = step._run(n, int(tune), trace=tr_view)
            elif kept and direct:
                # the kernel writes trace[:, i] = q (sampling.py:513) into the caller-visible pinned array itself
                tr_view = host_trace[:, done - keep_from:done - keep_from + n]
                _, st = step._run(n, int(tune), trace=tr_view)
            else:
                i = blk & 1
                if bufs[i] is None:
                    bufs[i] = torch.empty(chains, min(block, T), D, dtype=torch.float64, device=dev)
                if copy_done[i] is not None:
                    compute.wait_event(copy_done[i])       # the previous copy out of this buffer must have finished
                tr_view = bufs[i][:, :n]
                _, st = step._run(n, int(tune), trace=tr_view)
                if kept:
                    ready = torch.cuda.Event()
                    ready.record(compute)
                    copy_stream.wait_event(ready)
                    dst = host_trace[:, done - keep_from:done - keep_from + n]
                    # [chains] rows of n*D contiguous doubles each, different pitches on the two sides
                    L.check(L.load().lmc_memcpy2d_d2h(dst.data_ptr(), dst.stride(0) * 8, tr_view.data_ptr(),
                                                      tr_view.stride(0) * 8, n * D * 8, chains,
                                                      copy_stream.cuda_stream), "lmc_memcpy2d_d2h")
                    copy_done[i] = torch.cuda.Event()
                    copy_done[i].record(copy_stream)
            stats_blocks.append(st)
            if callback is not None:
                compute.synchronize()
                copy_stream.synchronize()
                _per_draw_callbacks(callback, step, st, tr_view, host_trace, done, n, T, int(tune), keep_from,
                                    int(chain_idx))
            done += n
            blk += 1
    except KeyboardInterrupt:                              # sampling.py:470-478: return what has been sampled so far
        interrupted = True
        _log.warning("Interrupted after %d of %d transitions.", done, T)
    _mark("enqueue")
    compute.synchronize()
    _mark("compute_sync")
    copy_stream.synchronize()
    _mark("copy_sync")
    step._check_status()
    if tune < done:
        step.stop_tuning()                                                          # sampling.py:510-511
    if interrupted:                                                                 # keep the finished transitions only
        n_keep = max(0, done - keep_from)
        if return_device:
            trace_out = trace_out[:, :n_keep]
        else:
            host_trace = host_trace[:, :n_keep]
    if not stats_blocks:
        stats_blocks = [torch.empty(chains, 0, L.NSTATS, dtype=torch.float64, device=dev)]
    stats_dev = torch.cat(stats_blocks, 1) if len(stats_blocks) > 1 else stats_blocks[0]
    stats_dev = stats_dev[:, :done]                     # an interrupted block's statistics are dropped with its draws
    step._account(stats_dev, min(int(tune), done))
    # leapfrogs of the WHOLE run (tuning included, whatever is shipped): BASELINE.json's metric counts them all
    count_col = step._stat_columns.get("tree_size", step._stat_columns.get("n_steps"))
    step._last_run_leapfrogs = float(stats_dev[:, :, count_col].sum().item()) if stats_dev.numel() else 0.0

    _mark("account")
    stats_kept = stats_dev[:, keep_from:]
    if return_device:
        if stats_as == "tensor":
            return trace_out, stats_kept.contiguous()
        stats = {name: stats_kept[:, :, col].unsqueeze(-1) for name, col in step._stat_columns.items()}
        return trace_out, stats
    # statistics: transpose on the device to one contiguous [chains, draws] plane per statistic, one pinned copy, and
    # only the integer / bool planes are converted on the host (the float64 ones are returned as views)
    planes = stats_kept.permute(2, 0, 1).contiguous()
    planes_h = torch.empty(planes.shape, dtype=torch.float64, pin_memory=True)
    planes_h.copy_(planes, non_blocking=True)
    compute.synchronize()
    sh = planes_h.numpy()
    stats = {}
    for name, dtype in step.stats_dtypes[0].items():                                # sampling.py:212-220
        plane = sh[step._stat_columns[name]][:, :, None]
        stats[name] = plane if np.dtype(dtype) == np.float64 else plane.astype(dtype)
    _mark("stats_to_host")
    return host_trace.numpy(), stats


def _per_draw_callbacks(callback, step, st_dev, tr_dev_or_host, host_trace, first, n, T, tune, keep_from, chain_idx=0):
    """Call the reference-style hook once per (draw, chain) of a finished block (sampling.py:303-308).  `Draw.chain`
    counts from `chain_idx`, like the reference's `chain=i + chain_idx` (sampling.py:173)."""
    st = st_dev.cpu().numpy()
    pts = tr_dev_or_host.cpu().numpy() if tr_dev_or_host.is_cuda else tr_dev_or_host.numpy()
    trace_np = None if host_trace is None else host_trace.numpy()
    dtypes = step.stats_dtypes[0]
    for j in range(n):
        idx = first + j
        for c in range(st.shape[0]):
            sd = {name: np.asarray(st[c, j, step._stat_columns[name]]).astype(dt)[()] for name, dt in dtypes.items()}
            callback(trace=trace_np, draw=Draw(c + chain_idx, idx == T - 1, idx, idx < tune, [sd], pts[c, j], None))


def init_nuts(logp_dlogp_func, model_ndim, init="auto", random_seed=None, **kwargs):
    """Mass-matrix initialisation for NUTS: reference sampling.py:524-605 (diagonal initialisers)."""
    if not isinstance(init, str):
        raise TypeError("init must be a string.")
    init = init.lower()
    if init == "auto":
        init = "jitter+adapt_diag"
    _log.info("Initializing NUTS using %s...", init)
    if random_seed is not None:
        np.random.seed(int(np.atleast_1d(random_seed)[0]))                          # :574-576
    if init == "adapt_diag":
        start = np.zeros(model_ndim)
    elif init == "jitter+adapt_diag":
        start = 2 * np.random.rand(model_ndim) - 1                                  # :584
    elif init == "adapt_full":
        start = np.zeros(model_ndim)                                                # :588-592
    elif init == "jitter+adapt_full":
        start = 2 * np.random.rand(model_ndim) - 1                                  # :593-597
    else:
        raise ValueError("Unknown initializer: {}.".format(init))
    if init.endswith("adapt_full"):
        from .quadpotential import QuadPotentialFullAdapt
        potential = QuadPotentialFullAdapt(model_ndim, start, np.eye(model_ndim), 10)
    else:
        potential = QuadPotentialDiagAdapt(model_ndim, start, np.ones(model_ndim), 10)  # :582,587
    step = NUTS(logp_dlogp_func=logp_dlogp_func, model_ndim=model_ndim, potential=potential, **kwargs)
    return start, step
