"""Host logic of the multi-GPU path on CPU: world_size-2 (and 3, ragged) gloo process groups.  The sampler itself has
no CPU path, so the local sampler is replaced by a deterministic stand-in that produces what the kernels would hand
back (device-style tensors keyed by the per-chain seed); what is tested is the sharding, the seed / start plumbing and
the final all-gather -- i.e. everything that differs between N = 1 and N > 1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_local_sample(logp_dlogp_func, model_ndim, draws, tune, step, chains, start, random_seed,
                       discard_tuned_samples, chain_idx, return_device, **kw):
    """Stand-in for sampling.sample(return_device=True): draw t of the chain with seed s is s + t + start (so every
    value identifies its chain, draw and start)."""
    assert return_device and len(random_seed) == chains
    n = draws if discard_tuned_samples else draws + tune
    seeds = torch.tensor(random_seed, dtype=torch.float64)
    st = torch.as_tensor(np.broadcast_to(np.asarray(start, dtype="d"), (chains, model_ndim)).copy())
    trace = seeds[:, None, None] + torch.arange(n, dtype=torch.float64)[None, :, None] + st[:, None, :]
    stats = {"tree_size": (seeds[:, None, None] % 7 + 1).expand(chains, n, 1).contiguous(),
             "chain": (torch.arange(chains, dtype=torch.float64) + chain_idx)[:, None, None].expand(chains, n, 1).contiguous()}
    return trace, stats


def _worker(rank, world, port, n_chains, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from littlemcmc_b200 import distributed as D
        from littlemcmc_b200 import targets
        tgt = targets.StdNormal(3)
        trace, stats = D.sample(tgt, 3, draws=4, tune=2, chains=n_chains, random_seed=123, discard_tuned_samples=False,
                                _local_sample=_fake_local_sample)
        np.save(os.path.join(out_dir, "trace_%d.npy" % rank), trace.numpy())
        np.save(os.path.join(out_dir, "chain_%d.npy" % rank), stats["chain"].numpy())
        lo, hi = D.shard_range(n_chains, rank, world)
        t_loc, _ = D.sample(tgt, 3, draws=4, tune=2, chains=n_chains, random_seed=123, discard_tuned_samples=False,
                            gather=False, _local_sample=_fake_local_sample)
        assert t_loc.shape[0] == hi - lo
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_chains", [(2, 8), (3, 8), (2, 5)])
def test_sharded_sample_equals_single_process(tmp_path, world, n_chains):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_chains, str(tmp_path)), nprocs=world, join=True)
    # single-process expectation: the same global seeds and the same (single) jittered start for all chains
    from littlemcmc_b200 import sampling, targets
    seeds = sampling._resolve_seeds(123, n_chains)
    start, _ = sampling.init_nuts(targets.StdNormal(3), 3, random_seed=seeds)
    expect, _ = _fake_local_sample(None, 3, 4, 2, None, n_chains, start, seeds, False, 0, True)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "trace_%d.npy" % r))
        assert got.shape == (n_chains, 6, 3)
        assert np.array_equal(got, expect.numpy()), "rank %d" % r
        chain = np.load(os.path.join(str(tmp_path), "chain_%d.npy" % r))
        assert np.array_equal(chain[:, 0, 0], np.arange(n_chains))      # chain_idx offsets line the shards up


def test_shard_ranges_partition_the_chains():
    from littlemcmc_b200.distributed import shard_range, shard_sizes
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
