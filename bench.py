"""bench.py -- leapfrog-steps/sec of the NUTS hot path (BASELINE.json's metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload headline|cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (default "headline", the north star's target workload): NUTS, 1024 chains per GPU,
1000-dim diagonal Gaussian (sigma_i = 10^linspace(-.5,.5)), QuadPotentialDiagAdapt + dual averaging, max_treedepth 10,
in-kernel Philox randomness.  One STEP = one launch of the sampler kernel = `--trans-per-step` consecutive NUTS
transitions of every chain (default 16: the block size littlemcmc_b200.sample() uses at this problem size; the default
25 steps after 5 warm-up steps cover iterations 80..480 of one run, 120 of them tuning) (momentum draw, tree building, both adaptations, trace + statistics written to HBM),
continuing one run: the first `--tune` transitions tune.  Only useful leapfrogs (sum of the `tree_size` statistic over
the timed steps) are counted.  Weak scaling: every GPU runs its own block of chains, no collective while sampling,
one NCCL all-gather of the last step's draws afterwards (timed separately, reported as `allgather_ms`).

Timing: CUDA events on the launching stream around every timed step; a 512 MiB buffer is rewritten between steps to
flush L2 (outside the events); barrier + synchronize before and after the timed region; max over ranks.
`e2e` is one `littlemcmc_b200.sample()` call (the public API) for the same number of transitions with HOST buffers in
and out: pinned start positions H2D, every draw of the trace and all statistics D2H, tuning included, wall clock.
`cpu_baseline` / `--impl reference`: the NumPy oracle port of the reference sampler (oracle/lmc_oracle.py, bit-identical
to eigenfoo/littlemcmc on the golden fixtures) on the box's host cores, one process per core, bounded sample.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (chains per GPU, ndim, target, max_treedepth, tune, description)
    "headline": (1024, 1000, "gauss", 10, 200, "NUTS, 1024 chains/GPU x 1000-dim diagonal Gaussian, max_treedepth=10"),
    "cfg2": (1024, 100, "gauss", 10, 200, "NUTS, 1024 chains x 100-dim diagonal Gaussian, max_treedepth=10"),
    "cfg3": (4096, 1000, "illcond", 10, 500, "NUTS, 4096 chains x 1000-dim ill-conditioned Gaussian (kappa=1e4)"),
    "cfg4": (8192, 50, "funnel", 12, 300, "NUTS, 8192 chains x 50-dim Neal's funnel, max_treedepth=12"),
    "cfg5": (8192, 1000, "gauss", 10, 200, "NUTS, 8192 chains/GPU x 1000-dim diagonal Gaussian (65536 chains on 8 GPUs)"),
}


def target_params(kind, D):
    """Parameters of the synthetic target densities."""
    if kind == "gauss":
        sigma = 10 ** np.linspace(-0.5, 0.5, D)
        return dict(tau=1 / sigma**2)
    if kind == "illcond":
        return dict(tau=1 / 10 ** np.linspace(0, 4, D))
    return dict()


def make_target_np(kind, D):
    """CPU arm only: the oracle's NumPy callable for the same density (the GPU arm never imports oracle/)."""
    from oracle import lmc_oracle as orc
    if kind in ("gauss", "illcond"):
        return orc.diag_gaussian(target_params(kind, D)["tau"])
    return orc.neal_funnel(D)


# ---- CPU arm: the oracle port of the reference on the host cores -------------------------------------------------------
def _cpu_worker(args):
    kind, D, max_depth, n_trans, n_tune, seed = args
    from oracle import lmc_oracle as orc
    f = make_target_np(kind, D)
    smp = orc.Sampler(f, D, orc.DiagPotential(D, var=np.ones(D), initial_mean=np.zeros(D), initial_weight=10.0),
                      kind="nuts", max_treedepth=max_depth)
    rng = np.random.RandomState(seed)
    t0 = time.perf_counter()
    _, st = orc.sample_chain(smp, np.zeros(D), n_trans - n_tune, n_tune, rng)
    dt = time.perf_counter() - t0
    return float(st["tree_size"].sum()), dt


def cpu_reference_step(kind, D, max_depth, n_trans, cores, seed0):
    """One bounded sample: every core runs one chain for n_trans transitions.
    -> (leapfrogs, slowest worker's seconds, sum of per-process leapfrog rates)."""
    jobs = [(kind, D, max_depth, n_trans, n_trans, seed0 + i) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    # aggregate = sum of per-process rates, interpreter / pool start-up excluded
    return sum(r[0] for r in res), max(r[1] for r in res), sum(r[0] / r[1] for r in res)


def run_reference_arm(args, wl):
    chains, D, kind, max_depth, tune, desc = wl
    cores = os.cpu_count() or 1
    # bounded: about 60 s of host work for the whole run whatever K is (a chain does ~150-400 transitions/s)
    n_trans = args.cpu_trans or max(40, min(2000, int(60.0 * 250 / max(1, args.steps))))
    for w in range(min(args.warmup, 2)):
        cpu_reference_step(kind, D, max_depth, max(8, n_trans // 10), cores, 1000 + w)
    tot_wall, rates = 0.0, []
    for k in range(args.steps):
        _, wall, rate = cpu_reference_step(kind, D, max_depth, n_trans, cores, 5000 + 97 * k)
        rates.append(rate)
        tot_wall += wall
    value = float(np.mean(rates))
    sample = "%d processes x 1 chain x %d tuning transitions per step (same target density, D=%d)" % (cores, n_trans, D)
    line = {
        "impl": "reference", "metric": "leapfrog-steps/sec (all chains)", "value": value, "unit": "leapfrog-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, desc), "ndim": D},
        "cpu_baseline": {"value": value, "unit": "leapfrog-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "leapfrog-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        # 10 ms: dense enough for a ~50 ms timed region, sparse enough that NVML queries (which take driver locks)
        # do not compete with the kernel launches of the loop being timed
        self.period = 0.01
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Multi-GPU runs: pin this rank to the CPU cores NVML reports as local to its GPU, so that the pinned host buffers
    of the end-to-end path are allocated (first touch) on the GPU's own NUMA node and its D2H traffic does not cross the
    socket interconnect.  Best effort: silently keeps the inherited affinity when NVML or the OS call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


# ---- GPU arm -----------------------------------------------------------------------------------------------------------
def run_gpu_arm(args, wl):
    import torch
    import torch.distributed as dist

    import littlemcmc_b200 as lmc
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200 import engine

    chains, D, kind, max_depth, tune, desc = wl
    if args.chains:
        chains = args.chains
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1 and not args.no_affinity:
        bind_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tparams = target_params(kind, D)
    target = (lmc.targets.NealFunnel(D) if kind == "funnel" else lmc.targets.DiagGaussian(tau=tparams["tau"]))
    if args.logp != "fused":   # the density as a batched torch op between launches (callback mode)
        target = target.torch_batched(dev, cuda_graph=(args.logp == "torch-graph"))
    tps = args.trans_per_step

    def make_step():
        pot = lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10)
        step = lmc.NUTS(target, D, potential=pot, max_treedepth=max_depth)
        step._knobs = dict(group=args.group, smem_vecs=args.smem_vecs, max_slots=args.max_slots)
        return step

    # -- kernel-level measurement (inputs resident in HBM) ---------------------------------------------------------------
    step = make_step()
    seeds = 1_000_003 * (rank + 1) + np.arange(chains)          # distinct streams on every rank
    ch = step._bind(chains, device=dev, seeds=seeds)
    step.reset_tuning()
    step.iter_count = 0
    ch.set_position(np.zeros(D))
    trace = torch.empty(chains, tps, D, dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    # every buffer the timed loop touches exists before it starts (no allocator calls between the events)
    stats_all = torch.empty(args.steps, chains, tps, L.NSTATS, dtype=torch.float64, device=dev)
    for _ in range(args.warmup):
        step._run(tps, tune, trace=trace, stats=stats_all[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()

    def timed_loop():
        evs = []
        l0 = engine.LAUNCH_COUNT["kernels"]
        t0_ = time.perf_counter()
        for k in range(args.steps):
            flush.zero_()                                       # L2 flush, outside the timed events
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            # the events are recorded on the launching stream immediately around the library call (engine.py)
            step._run(tps, tune, trace=trace, stats=stats_all[k], events=ev)
            evs.append(ev)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], time.perf_counter() - t0_, engine.LAUNCH_COUNT["kernels"] - l0

    step_ms, t_wall, n_launches = timed_loop()
    remeasured = False
    if sum(step_ms) > 2.0 * float(np.median(step_ms)) * len(step_ms):
        # a host stall landed inside an event bracket (the GPU idled waiting for the launch): take the K steps again
        # from the same chain state position in the run (the run simply continues; tuning is over by then)
        remeasured = True
        step_ms, t_wall, n_launches = timed_loop()
    stats_keep = [stats_all[k] for k in range(args.steps)]
    clocks.stop_flag = True
    clocks.join()
    if world > 1:
        dist.barrier()
    step._check_status()
    dev_ms = float(sum(step_ms))
    leap_per_step = [float(s[:, :, L.STAT_TREE_SIZE].sum().item()) for s in stats_keep]
    leapfrogs = float(sum(leap_per_step))
    depth_mean = float(torch.stack([s[:, :, L.STAT_DEPTH].mean() for s in stats_keep]).mean().item())
    accept_mean = float(torch.stack([s[:, :, L.STAT_ACCEPT].mean() for s in stats_keep]).mean().item())
    n_div = float(sum(s[:, :, L.STAT_DIVERGING].sum().item() for s in stats_keep))

    # -- the single collective of the design: all-gather the last step's draws ---------------------------------------------
    allgather_ms = None
    if world > 1:
        from littlemcmc_b200 import distributed as lmcd
        gathered = lmcd.gather_chains(trace, world * chains)    # warm-up (communicator setup)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gathered = lmcd.gather_chains(trace, world * chains)
        g1.record()
        torch.cuda.synchronize()
        allgather_ms = g0.elapsed_time(g1)
        agg = torch.tensor([leapfrogs, dev_ms, allgather_ms], dtype=torch.float64, device=dev)
        tot = agg.clone()
        dist.all_reduce(tot[0:1], op=dist.ReduceOp.SUM)
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        leapfrogs_all, dev_ms_max, allgather_ms = float(tot[0]), float(mx[1]), float(mx[2])
    else:
        leapfrogs_all, dev_ms_max = leapfrogs, dev_ms
    value = leapfrogs_all / (dev_ms_max * 1e-3)

    # -- end to end through the public API with host buffers -----------------------------------------------------------------
    # One warm-up call with the SAME shapes first: sample() returns its trace in pinned host memory, which torch's
    # caching host allocator hands back without a new cudaHostAlloc once a block of that size has been freed.
    # the same transitions as the timed steps, capped so that the pinned host trace stays below ~5 GB whatever --steps
    # and the workload are (640 transitions at the headline size); bytes are reported per step of `tps` transitions
    n_e2e = min(args.steps * tps, max(2 * tps, int(5.3e9 // (chains * D * 8))))
    e2e_steps = n_e2e / float(tps)
    e2e_tune = min(tune, n_e2e // 2)
    start_host = torch.zeros(chains, D, dtype=torch.float64).pin_memory()
    step2 = make_step()

    def e2e_call(seed_shift, discard=False):
        return lmc.sample(target, D, draws=n_e2e - e2e_tune, tune=e2e_tune, step=step2, chains=chains,
                          start=start_host.numpy(), random_seed=list(seeds + seed_shift), discard_tuned_samples=discard,
                          device=dev, progressbar=False)
    warm = e2e_call(5)
    del warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tr_h, st_h = e2e_call(17)
    e2e_s = time.perf_counter() - t0
    e2e_leap = float(st_h["tree_size"].sum())
    if world > 1:
        agg = torch.tensor([e2e_leap, e2e_s], dtype=torch.float64, device=dev)
        tot, mx = agg.clone(), agg.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        e2e_leap, e2e_s = float(tot[0]), float(mx[1])
    e2e_value = e2e_leap / e2e_s
    # for information: the same call with the API default discard_tuned_samples=True (only the post-tuning draws are
    # shipped to the host; every transition is still sampled).  Rank-local, not the reported e2e.
    warm = e2e_call(5, discard=True)                           # warm-up: pinned buffers of this (smaller) size
    del warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_call(17, discard=True)
    e2e_default_s = time.perf_counter() - t0
    h2d = (chains * D * 8 + chains * 8) / e2e_steps
    d2h = (tr_h.nbytes + chains * n_e2e * L.NSTATS * 8) / e2e_steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # -- roofline of the dominant (only) kernel -------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    bytes_per_leapfrog = 48 * D                                 # read q,p,g + write q',p',g' in fp64
    achieved = (leapfrogs / args.steps) * bytes_per_leapfrog / (dev_ms / args.steps * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)

    # -- CPU baseline on this box's host cores (bounded sample) ----------------------------------------------------------------
    cores = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu and world == 1:                          # the CPU baseline is an N=1 line
        n_cpu = args.cpu_trans or 3000
        _, _, cpu_rate = cpu_reference_step(kind, D, max_depth, n_cpu, cores, 4242)
        cpu = {"value": cpu_rate, "unit": "leapfrog-steps/s", "cores": cores, "kind": "port",
               "sample": "%d processes x 1 chain x %d tuning transitions of the same target (D=%d), sum of per-process "
                         "rates, oracle/lmc_oracle.py" % (cores, n_cpu, D)}

    line = {
        "metric": "leapfrog-steps/sec (all chains)", "value": value, "unit": "leapfrog-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, desc), "chains_per_gpu": chains, "ndim": D,
                   "transitions_per_step": tps, "tune": tune, "rng": "in-kernel Philox4x32-10", "logp": args.logp,
                   "cache": "512 MiB buffer rewritten between timed steps (L2 flush outside the CUDA events)",
                   "mean_tree_depth": depth_mean, "mean_tree_accept": accept_mean, "divergences": n_div,
                   "leapfrogs_timed": leapfrogs_all, "wall_ms_incl_flush": t_wall * 1e3,
                   "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(max(step_ms)),
                   "remeasured_after_host_stall": remeasured,
                   "parallelism": "chains sharded x%d, no collective while sampling" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "note": "achieved = algorithmic 48*D bytes per leapfrog x leapfrogs per launch / launch time; chain "
                             "state is register/shared-memory resident, so measured DRAM traffic is far below it"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "leapfrog-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "one littlemcmc_b200.sample() call, %d transitions (%d tuning), pinned host start in, full host "
                        "trace (tuning draws included: discard_tuned_samples=False) + stats out, wall clock %.1f ms; the "
                        "same run with the API default (tuning draws not shipped) takes %.1f ms on rank 0"
                        % (n_e2e, e2e_tune, e2e_s * 1e3, e2e_default_s * 1e3)},
        "gpu_launches": n_launches,   # fused: sched_init_kernel + sampler_kernel per step; callback mode: one per gradient
        "clocks": clocks.summary(),
    }
    if allgather_ms is not None:
        line["allgather_ms"] = allgather_ms
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--trans-per-step", type=int, default=16,
                    help="transitions per launch; 16 = the block sample() itself uses at 1024 chains x 1000 dimensions")
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--cpu-trans", type=int, default=0, help="transitions per CPU-baseline chain (0 = bounded default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-affinity", action="store_true", help="N>1: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--logp", default="fused", choices=["fused", "torch", "torch-graph"],
                    help="fused: density inside the kernel; torch: batched torch op between launches (callback mode)")
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--smem-vecs", type=int, default=-1)
    ap.add_argument("--max-slots", type=int, default=0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        run_reference_arm(args, wl)
        return
    run_gpu_arm(args, wl)


if __name__ == "__main__":
    main()
