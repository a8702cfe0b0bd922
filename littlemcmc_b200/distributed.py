"""Chains sharded over the GPUs of one node: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Chains share nothing -- separate seed (reference sampling.py:134,497), separate step-size state and mass matrix
(re-initialised per chain, sampling.py:504-505) -- so the path shards without any data-path collective: rank r runs the
single-GPU sampler on a contiguous block of chains, and ONE all-gather of the draws (and of the small statistics
tensors) follows at the end.  The reference's analogue of that gather is the host-side
``np.array([... for chain_trace in traces])`` (sampling.py:208); its analogue of the sharding is one OS process per
chain (parallel_sampling.py), which this replaces.

Per-chain seeds are derived from the GLOBAL seed list exactly as in the single-GPU driver, so chain c produces the same
draws whatever the number of ranks.
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


def gather_chains(local, n_chains, group=None):
    """All-gather along dim 0 (the chain dimension): [c_local, ...] on every rank -> [n_chains, ...] on every rank.

    Equal shards go through a single all_gather_into_tensor (one NCCL all-gather writing straight into the result);
    ragged shards (n_chains % world != 0) are padded to the largest shard for the collective and trimmed after."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_chains, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("local shard has %d chains, expected %d" % (local.shape[0], sizes[dist.get_rank(group)]))
    local = local.contiguous()
    out = torch.empty((n_chains,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
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


def sample(logp_dlogp_func, model_ndim, draws=1000, tune=1000, step=None, chains=None, start=None, random_seed=None,
           discard_tuned_samples=True, group=None, gather=True, _local_sample=None, **kwargs):
    """`littlemcmc_b200.sample` over all ranks of `group`: `chains` is the GLOBAL number of chains.

    Every rank must call this with the same arguments.  Returns ``(trace, stats)`` as device tensors: the gathered
    ``[chains, draws, ndim]`` trace and ``{name: [chains, draws, 1]}`` statistics on every rank (``gather=True``), or
    the rank's own shard (``gather=False``, e.g. to thin or reduce before exchanging)."""
    from . import sampling
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if chains is None:
        raise ValueError("distributed.sample needs the global number of chains")
    seeds = sampling._resolve_seeds(random_seed, chains)          # identical on every rank (same random_seed)
    lo, hi = shard_range(chains, rank, world)
    if step is None or start is None:
        # the reference draws ONE jittered start for all chains after reseeding with the first GLOBAL seed
        # (sampling.py:148-164, 574-584): do it here so the start does not depend on the number of ranks
        driver_keys = ("init", "cores", "progressbar", "chain_idx", "callback", "mp_ctx", "pickle_backend", "device",
                       "block")
        nuts_kwargs = {k: kwargs.pop(k) for k in list(kwargs) if k not in driver_keys}
        start_, step_ = sampling.init_nuts(logp_dlogp_func, model_ndim, init=kwargs.get("init", "auto"),
                                           random_seed=seeds, **nuts_kwargs)
        step = step_ if step is None else step
        start = start_ if start is None else start
    if start is not None:
        start = np.asarray(start, dtype="d")
        if start.ndim == 2:
            start = start[lo:hi]
    local = _local_sample or sampling.sample
    trace, stats = local(logp_dlogp_func, model_ndim, draws=draws, tune=tune, step=step, chains=hi - lo, start=start,
                         random_seed=seeds[lo:hi], discard_tuned_samples=discard_tuned_samples, chain_idx=lo,
                         return_device=True, **kwargs)
    if not gather:
        return trace, stats
    trace = gather_chains(trace, chains, group)
    stats = {k: gather_chains(v, chains, group) for k, v in stats.items()}
    return trace, stats
