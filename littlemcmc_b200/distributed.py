"""Chains sharded over the GPUs of one node: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Chains share nothing -- separate seed (reference sampling.py:134,497), separate step-size state and mass matrix
(re-initialised per chain, sampling.py:504-505) -- so the path shards without any data-path collective: rank r runs the
single-GPU sampler on a contiguous block of chains, and ONE exchange follows at the end: an all-gather of the draws
and an all-gather of the statistics tensor ([chains, draws, 13], all statistics packed, as the kernels write it).
The reference's analogue of that gather is the host-side ``np.array([... for chain_trace in traces])``
(sampling.py:208); its analogue of the sharding is one OS process per chain (parallel_sampling.py), which this
replaces.

Per-chain seeds and the single jittered start are resolved ONCE (rank 0) and broadcast, so chain c produces the same
draws whatever the number of ranks, also with ``random_seed=None``.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_chains, rank, world):
    """Contiguous block [lo, hi) of chains owned by `rank`: the first n_chains % world ranks get one extra chain."""
    n_chains, rank, world = int(n_chains), int(rank), int(world)
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_chains, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_chains, world):
    return [shard_range(n_chains, r, world)[1] - shard_range(n_chains, r, world)[0] for r in range(world)]


def gather_chains(local, n_chains, group=None, out=None):
    """All-gather along dim 0 (the chain dimension): [c_local, ...] on every rank -> [n_chains, ...] on every rank.

    Equal shards go through a single all_gather_into_tensor (one NCCL all-gather writing straight into the result);
    ragged shards (n_chains % world != 0) are padded to the largest shard for the collective and trimmed after.
    `out`: preallocated result (reused across calls: no allocation between the caller's timing events)."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_chains, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("local shard has %d chains, expected %d" % (local.shape[0], sizes[dist.get_rank(group)]))
    local = local.contiguous()
    shape = (n_chains,) + tuple(local.shape[1:])
    if out is None:
        out = torch.empty(shape, dtype=local.dtype, device=local.device)
    elif tuple(out.shape) != shape or out.dtype != local.dtype or not out.is_contiguous():
        raise ValueError("`out` must be a contiguous %s tensor of shape %s" % (local.dtype, shape))
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    big = max(sizes)
    padded = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    lo = 0
    for r, n in enumerate(sizes):
        out[lo:lo + n] = buf[r * big:r * big + n]
        lo += n
    return out


def gather_draw_chunks(local_trace, n_chains, draws_per_chunk, consume, group=None):
    """The final exchange in pieces, for traces too large to hold gathered: all-gather `draws_per_chunk` draws of every
    chain at a time into ONE reused buffer and hand each gathered block [n_chains, <=draws_per_chunk, ndim] to
    `consume(block, first_draw)` (thin it, reduce it to moments, write it out ...) before the next chunk overwrites it.
    Returns the number of collectives issued."""
    c_local, draws, D = local_trace.shape
    k = max(1, int(draws_per_chunk))
    buf = torch.empty((n_chains, min(k, draws), D), dtype=local_trace.dtype, device=local_trace.device)
    n = 0
    for lo in range(0, draws, k):
        hi = min(draws, lo + k)
        block = gather_chains(local_trace[:, lo:hi], n_chains, group, out=buf if hi - lo == buf.shape[1] else None)
        consume(block, lo)
        n += 1
    return n


def _resolve_run(logp_dlogp_func, model_ndim, chains, random_seed, step, start, kwargs, group):
    """Seeds, start and step method, identical on every rank (sampling.py:131-134, 148-164)."""
    from . import sampling
    rank = dist.get_rank(group)
    # keywords of the drivers (sample / distributed.sample); everything else configures the step method (init_nuts)
    driver_keys = ("init", "cores", "progressbar", "chain_idx", "callback", "mp_ctx", "pickle_backend", "device",
                   "block", "return_device", "host_write", "stats_as", "single_launch", "_timing")
    nuts_kwargs = {k: kwargs.pop(k) for k in list(kwargs) if k not in driver_keys}
    # the seed list depends on the process-global NumPy stream when random_seed is None: resolve it once, broadcast
    box = [sampling._resolve_seeds(random_seed, chains) if rank == 0 else None]
    if dist.get_world_size(group) > 1:
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
    seeds = box[0]
    if step is None or start is None:
        # the reference draws ONE jittered start for all chains after reseeding with the first GLOBAL seed, and seeds
        # the potential's running mean with it (sampling.py:148-164, 574-587): a pure function of the seed list, so every
        # rank computes the same start and the same step method whatever the number of ranks
        start_, step_ = sampling.init_nuts(logp_dlogp_func, model_ndim, init=kwargs.get("init", "auto"),
                                           random_seed=seeds, **nuts_kwargs)
        step = step_ if step is None else step
        start = start_ if start is None else start
    return seeds, np.asarray(start, dtype="d"), step


def sample(logp_dlogp_func, model_ndim, draws=1000, tune=1000, step=None, chains=None, start=None, random_seed=None,
           discard_tuned_samples=True, group=None, gather=True, out=None, _local_sample=None, **kwargs):
    """`littlemcmc_b200.sample` over all ranks of `group`: `chains` is the GLOBAL number of chains.

    Every rank must call this with the same arguments.  Returns ``(trace, stats)`` as device tensors: the gathered
    ``[chains, draws, ndim]`` trace and ``{name: [chains, draws, 1]}`` statistics on every rank (``gather=True``: two
    collectives, the draws and the packed statistics tensor), or the rank's own shard (``gather=False``, e.g. to thin
    or reduce before exchanging, or to exchange in pieces with `gather_draw_chunks`).  ``out``: optional preallocated
    ``[chains, draws kept, ndim]`` float64 device tensor for the gathered trace."""
    from . import sampling
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if chains is None:
        raise ValueError("distributed.sample needs the global number of chains")
    seeds, start, step = _resolve_run(logp_dlogp_func, model_ndim, chains, random_seed, step, start, kwargs, group)
    lo, hi = shard_range(chains, rank, world)
    if start.ndim == 2:
        start = start[lo:hi]
    kwargs.pop("return_device", None)
    if _local_sample is not None:          # CPU tests: a stand-in for the single-GPU driver (statistics as a dict)
        trace, stats = _local_sample(logp_dlogp_func, model_ndim, draws=draws, tune=tune, step=step, chains=hi - lo,
                                     start=start, random_seed=seeds[lo:hi], discard_tuned_samples=discard_tuned_samples,
                                     chain_idx=lo, return_device=True, **kwargs)
        if not gather:
            return trace, stats
        names = sorted(stats)
        packed = torch.cat([stats[k] for k in names], -1)
        trace = gather_chains(trace, chains, group, out=out)
        packed = gather_chains(packed, chains, group)
        return trace, {k: packed[..., i:i + 1] for i, k in enumerate(names)}
    trace, packed = sampling.sample(logp_dlogp_func, model_ndim, draws=draws, tune=tune, step=step, chains=hi - lo,
                                    start=start, random_seed=seeds[lo:hi], discard_tuned_samples=discard_tuned_samples,
                                    chain_idx=lo, return_device=True, stats_as="tensor", **kwargs)
    if gather:
        trace = gather_chains(trace, chains, group, out=out)
        packed = gather_chains(packed, chains, group)      # all 13 statistics of a draw in one row: ONE collective
    return trace, {name: packed[:, :, col].unsqueeze(-1) for name, col in step._stat_columns.items()}
