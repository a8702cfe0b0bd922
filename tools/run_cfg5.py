"""BASELINE config 5 through the product entry point: NUTS, 65536 chains (8192 per GPU), 1000-dim diagonal Gaussian,
chains sharded over the GPUs of one box by littlemcmc_b200.distributed.sample, final NCCL all-gather of the draws.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/run_cfg5.py [--chains-per-gpu 8192] [--tune 200] [--draws 50] [--gather full|chunks|none] [--check 4]

What it reports (one JSON line from rank 0, also written to gpurun_out/cfg5.json):
  * sampling: leapfrog-steps/s of the whole job (sum over ranks / slowest rank's device time), tree depth, acceptance;
  * the exchange: `full` = one all-gather of the kept draws [C, draws, D] + one of the packed statistics, onto every GPU
    (its time and bus bandwidth); `chunks` = the same bytes in pieces of --chunk-draws draws through one reused buffer,
    each gathered block reduced to per-chain moments on the spot, so the full gathered tensor never exists;
  * a per-rank oracle check: --check chains of every rank's shard are replayed on the CPU (oracle/lmc_oracle.py) from the
    device state they had when sampling started, with their own Philox streams (lmc_rng_fill) -- the first transitions of
    the production run, bit-for-bit decisions and 1e-9 floats, exactly the bar of tests/test_gpu_parity.py.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains-per-gpu", type=int, default=8192)
    ap.add_argument("--ndim", type=int, default=1000)
    ap.add_argument("--tune", type=int, default=200)
    ap.add_argument("--draws", type=int, default=50)
    ap.add_argument("--gather", default="full", choices=["full", "chunks", "none"])
    ap.add_argument("--chunk-draws", type=int, default=10)
    ap.add_argument("--check", type=int, default=4, help="chains per rank replayed in the CPU oracle")
    ap.add_argument("--check-trans", type=int, default=3)
    args = ap.parse_args()

    import littlemcmc_b200 as lmc
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200 import distributed as lmcd
    from littlemcmc_b200 import engine

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D, Cg = args.ndim, args.chains_per_gpu
    C = Cg * world
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    target = lmc.targets.DiagGaussian(sigma=sigma)
    seed = 5                                                   # cfg5, seed 5

    # ---- per-rank oracle check on the first transitions of the production run -----------------------------------------
    check = None
    if args.check > 0:
        from oracle import lmc_oracle as orc
        from tests import parity_utils as pu
        seeds_all = lmc.sampling._resolve_seeds(seed, C)
        lo, hi = lmcd.shard_range(C, rank, world)
        start, _ = lmc.sampling.init_nuts(target, D, random_seed=seeds_all)
        f = orc.diag_gaussian(1 / sigma**2)
        ch = engine.DeviceChains(hi - lo, D, dev)
        ch.reset_potential(np.ones(D), start, 10.0, 101)        # init_nuts: QuadPotentialDiagAdapt(D, start, ones, 10)
        ch.reset_step_adapt(0.25 / D ** 0.25)
        ch.set_position(start)
        seeds_t = engine.seeds_tensor(np.asarray(seeds_all[lo:hi]), dev)
        params = dict(adapt_mass=1, adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10, Emax=1000.0,
                      max_treedepth=10, early_max_treedepth=8)
        sel = torch.as_tensor(np.sort(np.random.RandomState(100 + rank).choice(hi - lo, args.check, replace=False)), device=dev)
        n_ok = 0
        for t in range(args.check_trans):
            wel = torch.stack([ch.mean_fg[sel, :D], ch.rawvar_fg[sel, :D], ch.mean_bg[sel, :D], ch.rawvar_bg[sel, :D]], 1)
            pre = (ch.q[sel, :D].cpu().numpy(), ch.var[sel, :D].cpu().numpy(), wel.cpu().numpy(), ch.adapt[sel, :9].cpu().numpy())
            _, st = engine.run_transitions(L.KIND_NUTS, ch, target.fused, n_trans=1, iter0=t, n_tune=10**9, params=params,
                                           seeds=seeds_t)
            torch.cuda.synchronize()
            post_q = ch.q[sel, :D].cpu().numpy()
            st = st[sel, 0].cpu().numpy()
            normals, uniforms = engine.rng_fill(seeds_t[sel], D, t, 1, 1100)
            normals, uniforms = normals.cpu().numpy(), uniforms.cpu().numpy()
            for j in range(args.check):
                smp = pu.oracle_sampler_from_state(f, D, pre[0][j], pre[1][j], pre[2][j], pre[3][j], iter_count=t, tune=True,
                                                   max_treedepth=10, early_max_treedepth=8)
                q, sd = smp.astep(pre[0][j], orc.TapeRNG(normals[j], uniforms[j]))
                assert int(sd["tree_size"]) == int(st[j, L.STAT_TREE_SIZE]) and int(sd["depth"]) == int(st[j, L.STAT_DEPTH])
                np.testing.assert_allclose(post_q[j], q, rtol=1e-9, atol=1e-12)
                np.testing.assert_allclose(st[j, L.STAT_ENERGY], sd["energy"], rtol=1e-9)
                n_ok += 1
        check = n_ok
        del ch

    # ---- the production run ------------------------------------------------------------------------------------------------------
    kw = dict(draws=args.draws, tune=args.tune, chains=C, random_seed=seed, progressbar=False, device=dev)
    lmcd.sample(target, D, **dict(kw, draws=2, tune=4), gather=(args.gather == "full"))     # warm-up: communicator, kernels
    torch.cuda.synchronize()
    dist.barrier()
    step = lmc.NUTS(target, D, potential=lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10))
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trace, stats = lmcd.sample(target, D, step=step, start=np.zeros(D), gather=False, **kw)
    e1.record()
    torch.cuda.synchronize()
    t_sample = time.perf_counter() - t0
    leap = float(step._last_run_leapfrogs)
    depth = float(stats["depth"].double().mean())
    accept = float(stats["mean_tree_accept"].mean())

    gather = {"mode": args.gather}
    if args.gather != "none":
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        if args.gather == "full":
            out = torch.empty(C, args.draws, D, dtype=torch.float64, device=dev)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            full = lmcd.gather_chains(trace, C, out=out)
            g1.record()
            torch.cuda.synchronize()
            gather.update(ms=g0.elapsed_time(g1), bytes=full.numel() * 8, collectives=1,
                          checksum=float(full[::1024, -1, 0].sum()))
        else:
            mom = torch.zeros(C, D, dtype=torch.float64, device=dev)

            def consume(block, first):
                mom.add_(block.sum(1))                           # per-chain running sums: the block is dropped afterwards
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            n = lmcd.gather_draw_chunks(trace, C, args.chunk_draws, consume)
            g1.record()
            torch.cuda.synchronize()
            gather.update(ms=g0.elapsed_time(g1), bytes=C * args.draws * D * 8, collectives=n,
                          checksum=float(mom[::1024, 0].sum()), chunk_draws=args.chunk_draws)
        gather["wall_ms"] = (time.perf_counter() - t1) * 1e3

    agg = torch.tensor([leap, t_sample, gather.get("ms", 0.0)], dtype=torch.float64, device=dev)
    tot, mx = agg.clone(), agg.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    chk = torch.tensor([float(check or 0)], dtype=torch.float64, device=dev)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    if rank == 0:
        line = {"workload": "cfg5: NUTS, %d chains (%d per GPU) x %d-dim diagonal Gaussian on %d GPUs, tune %d, draws %d"
                            % (C, Cg, D, world, args.tune, args.draws),
                "entry": "littlemcmc_b200.distributed.sample", "n_gpus": world,
                "leapfrog_steps_per_s": float(tot[0]) / float(mx[1]), "leapfrogs": float(tot[0]),
                "sample_seconds_max_over_ranks": float(mx[1]), "mean_tree_depth_rank0": depth, "mean_tree_accept_rank0": accept,
                "oracle_checked_transitions": int(chk[0]), "gather": gather}
        if "ms" in gather:
            gather["ms_max_over_ranks"] = float(mx[2])
            gather["bus_bandwidth_GBs"] = gather["bytes"] * (world - 1) / world / (float(mx[2]) * 1e-3) / 1e9
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "cfg5_%s_n%d.json" % (args.gather, world)), "w") as fh:
            fh.write(json.dumps(line) + "\n")
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
