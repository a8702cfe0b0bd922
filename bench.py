"""bench.py -- leapfrog-steps/sec of the NUTS hot path (BASELINE.json's metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload headline|cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (default "headline", the north star's target workload): NUTS, 1024 chains per GPU,
1000-dim diagonal Gaussian (sigma_i = 10^linspace(-.5,.5)), QuadPotentialDiagAdapt + dual averaging, max_treedepth 10,
in-kernel Philox randomness.  One STEP = one launch of the sampler kernel = `transitions_per_step` consecutive NUTS
transitions of every chain (momentum draw, tree building, both adaptations, trace + statistics written to HBM),
continuing one run: the first `tune` transitions tune.  Only useful leapfrogs (sum of the `tree_size` statistic over the
timed steps) are counted.  Weak scaling: every GPU runs its own block of chains, no collective while sampling.

Timing: CUDA events on the launching stream around every timed step; a 512 MiB buffer is rewritten between steps to
flush L2 (outside the events); barrier + synchronize before and after the timed region; max over ranks.

The one JSON line also carries
  e2e          one `littlemcmc_b200.sample()` call with the API defaults (discard_tuned_samples=True) and HOST buffers:
               pinned start positions in, kept draws + statistics out, tuning included, wall clock;
  configs      the other BASELINE configurations measured the same way in the same process (cfg2, cfg3, cfg4 with the
               fused kernels; cfg2 / cfg4 with the density as a torch op, device-driven loop; cfg2 with the density as
               user CUDA source compiled at run time), each with its own roofline fraction;
  cpu_baseline the reference on this box's host cores (N = 1 only), bounded sample;
  dist_e2e     (N > 1) `littlemcmc_b200.distributed.sample(gather=True)`: sharded sampling + the final NCCL all-gather of
               draws and statistics, with the all-gather's bus bandwidth and the box's concurrent D2H ceiling.
`--impl reference` times the UNMODIFIED reference (oracle/_ref, installed from /root/reference by oracle/build_ref.py;
the oracle port when that copy is absent) on all host cores, one process per core, same target density.
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
    # name: (chains per GPU, ndim, target, max_treedepth, tune, transitions per step, description)
    "headline": (1024, 1000, "gauss", 10, 200, 16, "NUTS, 1024 chains/GPU x 1000-dim diagonal Gaussian, max_treedepth=10"),
    "cfg2": (1024, 100, "gauss", 10, 200, 64, "NUTS, 1024 chains x 100-dim diagonal Gaussian, max_treedepth=10"),
    "cfg3": (4096, 1000, "illcond", 10, 500, 16, "NUTS, 4096 chains x 1000-dim ill-conditioned Gaussian (kappa=1e4)"),
    "cfg4": (8192, 50, "funnel", 12, 300, 40, "NUTS, 8192 chains x 50-dim Neal's funnel, max_treedepth=12"),
    "cfg5": (8192, 1000, "gauss", 10, 200, 8, "NUTS, 8192 chains/GPU x 1000-dim diagonal Gaussian (65536 chains on 8 GPUs)"),
}


def static_config(name, gpus, logp="fused"):
    """The workload description both arms print (identical keys and values: it names the problem, not the run)."""
    chains, D, kind, max_depth, tune, tps, desc = WORKLOADS[name]
    return {"workload": "%s: %s" % (name, desc), "chains_per_gpu": chains, "ndim": D, "target": kind,
            "max_treedepth": max_depth, "tune": tune, "transitions_per_step": tps,
            "potential": "QuadPotentialDiagAdapt(mean 0, var 1, weight 10), dual averaging target_accept 0.8",
            "start": "zeros", "logp": logp,
            "cache": "512 MiB buffer rewritten between timed steps (L2 flush outside the CUDA events)",
            "parallelism": "chains sharded x%d, no collective while sampling" % gpus}


def target_params(kind, D):
    """Parameters of the synthetic target densities."""
    if kind == "gauss":
        sigma = 10 ** np.linspace(-0.5, 0.5, D)
        return dict(tau=1 / sigma**2)
    if kind == "illcond":
        return dict(tau=1 / 10 ** np.linspace(0, 4, D))
    return dict()


def numpy_target(kind, D):
    """The density as a reference-style callable q[D] -> (logp, dlogp[D]) (base_hmc.py:34), for the CPU arms."""
    if kind in ("gauss", "illcond"):
        tau = target_params(kind, D)["tau"]

        def f(q):
            g = -(tau * q)
            return 0.5 * np.dot(q, g), g
        return f
    inv_s2, half_nm1 = 1.0 / 9.0, 0.5 * (D - 1)

    def funnel(q):
        v, x = q[0], q[1:]
        S = np.dot(x, x)
        with np.errstate(over="ignore", invalid="ignore"):
            ev = np.exp(-v)
            g = np.empty_like(q)
            g[1:] = -(ev * x)
            hs = 0.5 * ev * S
            g[0] = -(v * inv_s2) + hs - half_nm1
            return -(0.5 * v * v * inv_s2) - hs - half_nm1 * v, g
    return funnel


# ---- CPU arm: the reference itself (oracle/_ref) or its oracle port, on the host cores --------------------------------------
def reference_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "littlemcmc", "__init__.py"))


def _cpu_worker(args):
    kind, D, max_depth, n_trans, n_tune, seed, use_ref = args
    f = numpy_target(kind, D)
    if use_ref:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        ref = build_ref.import_reference()            # the unmodified eigenfoo/littlemcmc package under oracle/_ref
        import logging
        logging.getLogger("littlemcmc").setLevel(logging.ERROR)
        pot = ref.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10, dtype="float64")
        step = ref.NUTS(logp_dlogp_func=f, model_ndim=D, potential=pot, max_treedepth=max_depth)
        t0 = time.perf_counter()
        _, st = ref.sample(f, D, draws=n_trans - n_tune, tune=n_tune, step=step, chains=1, cores=1, start=np.zeros(D),
                           progressbar=False, random_seed=[int(seed)], discard_tuned_samples=False)
        dt = time.perf_counter() - t0
        return float(st["tree_size"].sum()), dt
    from oracle import lmc_oracle as orc
    smp = orc.Sampler(f, D, orc.DiagPotential(D, var=np.ones(D), initial_mean=np.zeros(D), initial_weight=10.0),
                      kind="nuts", max_treedepth=max_depth)
    rng = np.random.RandomState(seed)
    t0 = time.perf_counter()
    _, st = orc.sample_chain(smp, np.zeros(D), n_trans - n_tune, n_tune, rng)
    dt = time.perf_counter() - t0
    return float(st["tree_size"].sum()), dt


def cpu_reference_step(kind, D, max_depth, n_trans, n_tune, cores, seed0, use_ref):
    """One bounded sample: every core runs one chain for n_trans transitions (n_tune of them tuning).
    -> (leapfrogs, slowest worker's seconds, sum of per-process leapfrog rates)."""
    jobs = [(kind, D, max_depth, n_trans, n_tune, seed0 + i, use_ref) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    # aggregate = sum of per-process rates, interpreter / pool start-up excluded
    return sum(r[0] for r in res), max(r[1] for r in res), sum(r[0] / r[1] for r in res)


def cpu_sample_text(use_ref, cores, n_trans, n_tune, D, chains):
    return ("%s; %d processes x 1 chain x %d transitions (%d tuning, start zeros) of the same target (D=%d), sum of "
            "per-process leapfrog rates; chains are independent, so the rate for the workload's %d chains on these cores "
            "is the same number (the job would take chains/cores times longer)"
            % ("unmodified reference (oracle/_ref, littlemcmc 0.2.2) sample(cores=1)" if use_ref
               else "oracle/lmc_oracle.py port of the reference", cores, n_trans, n_tune, D, chains))


def run_reference_arm(args):
    chains, D, kind, max_depth, tune, tps, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    use_ref = reference_available()
    # bounded: about a minute of host work for the whole run whatever K is
    per_s = 60.0 if use_ref else 150.0                     # transitions per second per core, roughly (D = 1000)
    n_trans = args.cpu_trans or max(40, min(2000, int(60.0 * per_s * (1000.0 / max(D, 50)) ** 0.5 / max(1, args.steps))))
    n_tune = min(tune, n_trans // 2)
    for w in range(min(args.warmup, 2)):
        cpu_reference_step(kind, D, max_depth, max(8, n_trans // 10), max(4, n_trans // 20), cores, 1000 + w, use_ref)
    tot_wall, rates = 0.0, []
    for k in range(args.steps):
        _, wall, rate = cpu_reference_step(kind, D, max_depth, n_trans, n_tune, cores, 5000 + 97 * k, use_ref)
        rates.append(rate)
        tot_wall += wall
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "leapfrog-steps/sec (all chains)", "value": value, "unit": "leapfrog-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "leapfrog-steps/s", "cores": cores,
                         "kind": "reference" if use_ref else "port",
                         "sample": cpu_sample_text(use_ref, cores, n_trans, n_tune, D, chains)},
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


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


# ---- GPU arm -----------------------------------------------------------------------------------------------------------
def make_target(lmc, kind, D, dev, logp):
    tparams = target_params(kind, D)
    if logp == "user-source":     # the density as user CUDA source, compiled into the fused kernel at run time (NVRTC)
        assert kind in ("gauss", "illcond")
        return lmc.targets.ElementwiseTarget(D, logp="0.5 * q * g", grad="-(tau * q)",
                                             params={"tau": tparams["tau"]})
    target = (lmc.targets.NealFunnel(D) if kind == "funnel" else lmc.targets.DiagGaussian(tau=tparams["tau"]))
    if logp != "fused":           # the density as a batched torch op around the state-machine kernel (callback mode)
        target = target.torch_batched(dev, cuda_graph=(logp == "torch-graph"))
    return target


def measure_kernel(name, logp, steps, warmup, rank, dev, knobs=None, chains=0, tps=0, clocks=None):
    """Kernel-level measurement of one workload: inputs resident in HBM, CUDA events around every launch, L2 flushed
    between launches.  -> dict (per-rank numbers; the caller reduces over ranks)."""
    import torch

    import littlemcmc_b200 as lmc
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200 import engine
    chains0, D, kind, max_depth, tune, tps0, desc = WORKLOADS[name]
    chains, tps = chains or chains0, tps or tps0
    target = make_target(lmc, kind, D, dev, logp)
    pot = lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10)
    step = lmc.NUTS(target, D, potential=pot, max_treedepth=max_depth)
    step._knobs = knobs or {}
    seeds = 1_000_003 * (rank + 1) + np.arange(chains)          # distinct streams on every rank
    ch = step._bind(chains, device=dev, seeds=seeds)
    step.reset_tuning()
    step.iter_count = 0
    ch.set_position(np.zeros(D))
    trace = torch.empty(chains, tps, D, dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    # every buffer the timed loop touches exists before it starts (no allocator calls between the events)
    stats_all = torch.empty(steps, chains, tps, L.NSTATS, dtype=torch.float64, device=dev)
    for _ in range(warmup):
        step._run(tps, tune, trace=trace, stats=stats_all[0])
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.start()

    def timed_loop():
        evs = []
        l0 = engine.LAUNCH_COUNT["kernels"]
        t0_ = time.perf_counter()
        for k in range(steps):
            flush.zero_()                                       # L2 flush, outside the timed events
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            # the events are recorded on the launching stream immediately around the library call (engine.py)
            step._run(tps, tune, trace=trace, stats=stats_all[k], events=ev)
            evs.append(ev)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], time.perf_counter() - t0_, engine.LAUNCH_COUNT["kernels"] - l0

    step_ms, t_wall, n_launches = timed_loop()
    remeasured = False
    med_ms = float(np.median(step_ms))
    if sum(step_ms) > 2.0 * med_ms * len(step_ms) and sum(1 for x in step_ms if x > 3.0 * med_ms) == 1:
        # a host stall landed inside ONE event bracket (the GPU idled waiting for the launch): take the K steps again
        # from the same chain state position in the run (the run simply continues; tuning is over by then).  Several
        # long steps are the workload, not the host (cfg4: once the early tree-depth cap lifts, a few chains in the
        # funnel's neck build depth-12 trees and the steps get 5-10x longer): those are kept as measured.
        remeasured = True
        step_ms, t_wall, n_launches = timed_loop()
    if clocks is not None:
        clocks.stop_flag = True
        clocks.join()
    step._check_status()
    dev_ms = float(sum(step_ms))
    st = stats_all
    leap_each = [float(x) for x in st[..., L.STAT_TREE_SIZE].sum((1, 2)).tolist()]
    leapfrogs = float(sum(leap_each))
    peak, peak_src = hbm_peak()
    achieved = leapfrogs * 48 * D / (dev_ms * 1e-3) / 1e9      # 48*D algorithmic bytes per leapfrog
    out = dict(value=leapfrogs / (dev_ms * 1e-3), ms_per_step=dev_ms / steps, leapfrogs=leapfrogs, dev_ms=dev_ms,
               roofline_frac=achieved / peak, achieved_gbs=achieved, steps=steps, warmup=warmup, chains=chains, ndim=D,
               transitions_per_step=tps, logp=logp, gpu_launches=n_launches,
               mean_tree_depth=float(st[..., L.STAT_DEPTH].mean().item()),
               mean_tree_accept=float(st[..., L.STAT_ACCEPT].mean().item()),
               divergences=float(st[..., L.STAT_DIVERGING].sum().item()),
               ms_per_step_median=float(np.median(step_ms)), ms_per_step_max=float(max(step_ms)),
               ms_per_step_each=[round(float(x), 3) for x in step_ms], leapfrogs_per_step_each=leap_each,
               max_tree_depth=float(st[..., L.STAT_DEPTH].max().item()),
               wall_ms_incl_flush=t_wall * 1e3, remeasured_after_host_stall=remeasured)
    del flush, stats_all
    return out, step, trace, seeds


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import littlemcmc_b200 as lmc
    from littlemcmc_b200 import _lib as L

    chains, D, kind, max_depth, tune, tps, desc = WORKLOADS[args.workload]
    chains = args.chains or chains
    tps = args.trans_per_step or tps
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1 and not args.no_affinity:
        bind_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    knobs = dict(group=args.group, smem_vecs=args.smem_vecs, max_slots=args.max_slots, chunk=args.chunk)

    # -- kernel-level measurement (inputs resident in HBM) ---------------------------------------------------------------
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    head, step, trace, seeds = measure_kernel(args.workload, args.logp, args.steps, args.warmup, rank, dev, knobs=knobs,
                                              chains=chains, tps=tps, clocks=clocks)
    if world > 1:
        dist.barrier()
    leapfrogs, dev_ms = head["leapfrogs"], head["dev_ms"]

    # -- the exchange of the design: all-gather of one step's draws, into a preallocated result ------------------------------
    allgather = None
    if world > 1:
        from littlemcmc_b200 import distributed as lmcd
        out = torch.empty(world * chains, tps, D, dtype=torch.float64, device=dev)
        lmcd.gather_chains(trace, world * chains, out=out)      # warm-up (communicator setup)
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            dist.barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            lmcd.gather_chains(trace, world * chains, out=out)
            g1.record()
            torch.cuda.synchronize()
            times.append(g0.elapsed_time(g1))
        agg = torch.tensor([leapfrogs, dev_ms, min(times)], dtype=torch.float64, device=dev)
        tot = agg.clone()
        dist.all_reduce(tot[0:1], op=dist.ReduceOp.SUM)
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        leapfrogs_all, dev_ms_max, ag_ms = float(tot[0]), float(mx[1]), float(mx[2])
        nbytes = out.numel() * 8
        allgather = {"ms": ag_ms, "bytes_gathered_per_rank": nbytes,
                     "bus_bandwidth_GBs": nbytes * (world - 1) / world / (ag_ms * 1e-3) / 1e9,
                     "note": "all_gather_into_tensor of one step's draws [%d x %d x %d] f64 into a preallocated result, "
                             "best of 3, max over ranks; bus bandwidth = bytes x (N-1)/N / time" % (world * chains, tps, D)}
        del out
    else:
        leapfrogs_all, dev_ms_max = leapfrogs, dev_ms
    value = leapfrogs_all / (dev_ms_max * 1e-3)

    # -- end to end through the public API with host buffers -----------------------------------------------------------------
    # One warm-up call with the SAME shapes first: sample() returns its trace in pinned host memory, which torch's
    # caching host allocator hands back without a new cudaHostAlloc once a block of that size has been freed.
    # The same transitions as the timed steps, capped so that the pinned host trace stays below ~5 GB whatever --steps
    # and the workload are; bytes are reported per step of `tps` transitions.
    target = make_target(lmc, kind, D, dev, args.logp)
    n_e2e = min(args.steps * tps, max(2 * tps, int(5.3e9 // (chains * D * 8))))
    e2e_steps = n_e2e / float(tps)
    e2e_tune = min(tune, n_e2e // 2)
    start_host = torch.zeros(chains, D, dtype=torch.float64).pin_memory()

    def make_step():
        pot = lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10)
        s = lmc.NUTS(target, D, potential=pot, max_treedepth=max_depth)
        s._knobs = knobs
        return s
    step2 = make_step()

    def e2e_call(seed_shift, discard=True):
        return lmc.sample(target, D, draws=n_e2e - e2e_tune, tune=e2e_tune, step=step2, chains=chains,
                          start=start_host.numpy(), random_seed=list(seeds + seed_shift), discard_tuned_samples=discard,
                          device=dev, progressbar=False)

    def timed_e2e(discard):
        warm = e2e_call(5, discard)
        del warm
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        tr_h, st_h = e2e_call(17, discard)
        secs = time.perf_counter() - t0
        # every transition is sampled whatever is shipped: count the leapfrogs of the whole run from the step's counters
        leap = float(step2._last_run_leapfrogs)
        nbytes = tr_h.nbytes + sum(v.nbytes for v in st_h.values())
        if world > 1:
            a2 = torch.tensor([leap, secs], dtype=torch.float64, device=dev)
            tot2, mx2 = a2.clone(), a2.clone()
            dist.all_reduce(tot2, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx2, op=dist.ReduceOp.MAX)
            leap, secs = float(tot2[0]), float(mx2[1])
        return leap / secs, secs, nbytes
    e2e_value, e2e_s, e2e_bytes = timed_e2e(True)
    e2e_full_value, e2e_full_s, e2e_full_bytes = timed_e2e(False)
    h2d = (chains * D * 8 + chains * 8) / e2e_steps
    d2h = e2e_bytes / e2e_steps

    # -- N > 1: the product entry point, distributed.sample(gather=True), and the box's concurrent D2H ceiling ---------------
    dist_e2e = None
    if world > 1:
        from littlemcmc_b200 import distributed as lmcd
        step3 = make_step()
        kw = dict(draws=n_e2e - e2e_tune, tune=e2e_tune, step=step3, chains=world * chains, start=np.zeros(D),
                  device=dev, progressbar=False)
        lmcd.sample(target, D, random_seed=11, **kw)            # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        tr_g, st_g = lmcd.sample(target, D, random_seed=12, **kw)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        a3 = torch.tensor([float(step3._last_run_leapfrogs), secs], dtype=torch.float64, device=dev)
        tot3, mx3 = a3.clone(), a3.clone()
        dist.all_reduce(tot3, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx3, op=dist.ReduceOp.MAX)
        dist_e2e = {"value": float(tot3[0]) / float(mx3[1]), "unit": "leapfrog-steps/s", "seconds": float(mx3[1]),
                    "gathered_shape": list(tr_g.shape), "collectives": 2,
                    "note": "littlemcmc_b200.distributed.sample(gather=True): %d chains sharded x%d, %d transitions (%d "
                            "tuning), kept draws and the packed statistics all-gathered onto every GPU (device results), "
                            "wall clock, max over ranks" % (world * chains, world, n_e2e, e2e_tune)}
        del tr_g, st_g
        # concurrent device -> pinned-host copies on all ranks: what the box can absorb when every GPU ships its trace
        src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        dst = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        a4 = torch.tensor([secs], dtype=torch.float64, device=dev)
        dist.all_reduce(a4, op=dist.ReduceOp.MAX)
        ceiling = world * 2 * (1 << 30) / float(a4[0]) / 1e9
        dist_e2e["d2h_ceiling_GBs_all_ranks"] = ceiling
        dist_e2e["e2e_d2h_GBs_all_ranks"] = world * e2e_bytes / e2e_s / 1e9
        del src, dst

    # -- the other BASELINE configurations, same measurement, same process (rank 0, N = 1) ------------------------------------
    configs = None
    if world == 1 and not args.no_configs and args.workload == "headline" and args.logp == "fused":
        configs = {}
        extra = [("cfg2", "cfg2", "fused"), ("cfg3", "cfg3", "fused"), ("cfg4", "cfg4", "fused"),
                 ("cfg5_shard", "cfg5", "fused"), ("cfg2_torch_graph", "cfg2", "torch-graph"),
                 ("cfg4_torch_graph", "cfg4", "torch-graph"), ("cfg2_user_source", "cfg2", "user-source")]
        for key, wl, mode in extra:
            # cfg4's run is 400 transitions (tune 300 + draws 100): 3 warm-up + 7 timed steps of 40 cover it exactly
            n_steps = 4 if mode == "torch-graph" else 7 if wl == "cfg4" else 8
            try:
                r, *_ = measure_kernel(wl, mode, n_steps, 3, rank, dev)
                for k in ("leapfrogs", "dev_ms"):
                    r.pop(k)
                r["workload"] = "%s: %s" % (wl, WORKLOADS[wl][6])
                if wl == "cfg4":
                    # steps whose transitions all run under the early depth cap (iteration < 200) vs the rest
                    tps_ = r["transitions_per_step"]
                    capped = [k for k in range(n_steps) if (3 + k + 1) * tps_ <= 200]
                    rest = [k for k in range(n_steps) if k not in capped]
                    for name_, ks in (("value_steps_under_early_depth_cap", capped), ("value_steps_after_cap_lifts", rest)):
                        if ks:
                            r[name_] = (sum(r["leapfrogs_per_step_each"][k] for k in ks) /
                                        (sum(r["ms_per_step_each"][k] for k in ks) * 1e-3))
                    r["note"] = ("a launch lasts as long as its slowest chain: once the early depth cap (8, while tuning and "
                                 "iter_count < 200) lifts, the few chains sitting in the funnel's neck with a tiny adapted "
                                 "step size build depth-12 trees (4095 leapfrogs) transition after transition, strictly "
                                 "one after the other, while the other chains' groups idle -- see ms_per_step_each: the steps "
                                 "before iteration 200 run at the kernel's throughput, the later ones at one chain's latency")
                configs[key] = r
            except Exception as e:            # a failed side measurement must not take the headline line down
                configs[key] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # -- roofline of the dominant (only) kernel -------------------------------------------------------------------------------
    peak, peak_src = hbm_peak()
    bytes_per_leapfrog = 48 * D                                 # read q,p,g + write q',p',g' in fp64
    achieved = (leapfrogs / args.steps) * bytes_per_leapfrog / (dev_ms / args.steps * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get(args.workload)
        traffic_src = "profiles/traffic.json (%s)" % tj.get(args.workload + "_capture", "ncu --set full capture")

    # -- CPU baseline on this box's host cores (bounded sample) ----------------------------------------------------------------
    cores = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu and world == 1:                          # the CPU baseline is an N=1 line
        use_ref = reference_available()
        n_cpu = args.cpu_trans or (1200 if use_ref else 3000)
        n_cpu_tune = min(tune, n_cpu // 2)
        _, _, cpu_rate = cpu_reference_step(kind, D, max_depth, n_cpu, n_cpu_tune, cores, 4242, use_ref)
        cpu = {"value": cpu_rate, "unit": "leapfrog-steps/s", "cores": cores, "kind": "reference" if use_ref else "port",
               "sample": cpu_sample_text(use_ref, cores, n_cpu, n_cpu_tune, D, chains)}

    cfg = static_config(args.workload, world, args.logp)
    cfg["chains_per_gpu"], cfg["transitions_per_step"] = chains, tps
    line = {
        "metric": "leapfrog-steps/sec (all chains)", "value": value, "unit": "leapfrog-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "run": {"rng": "in-kernel Philox4x32-10", "mean_tree_depth": head["mean_tree_depth"],
                "mean_tree_accept": head["mean_tree_accept"], "divergences": head["divergences"],
                "leapfrogs_timed": leapfrogs_all, "wall_ms_incl_flush": head["wall_ms_incl_flush"],
                "ms_per_step_median": head["ms_per_step_median"], "ms_per_step_max": head["ms_per_step_max"],
                "remeasured_after_host_stall": head["remeasured_after_host_stall"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "note": "achieved = algorithmic 48*D bytes per leapfrog x leapfrogs per launch / launch time; chain "
                             "state is register/shared-memory resident, so measured DRAM traffic is far below it"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "leapfrog-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "one littlemcmc_b200.sample() call with the API defaults, %d transitions (%d tuning, not shipped: "
                        "discard_tuned_samples=True), pinned host start in, kept draws + statistics out to host memory "
                        "(%.2f GB), wall clock %.1f ms; with discard_tuned_samples=False (every draw shipped, %.2f GB): "
                        "%.3e leapfrog-steps/s, %.1f ms"
                        % (n_e2e, e2e_tune, e2e_bytes / 1e9, e2e_s * 1e3, e2e_full_bytes / 1e9, e2e_full_value,
                           e2e_full_s * 1e3)},
        "gpu_launches": head["gpu_launches"],   # fused: sched_init_kernel + sampler kernel per step
        "clocks": clocks.summary(),
    }
    if configs is not None:
        line["configs"] = configs
    if allgather is not None:
        line["allgather"] = allgather
        line["allgather_ms"] = allgather["ms"]
    if dist_e2e is not None:
        line["dist_e2e"] = dist_e2e
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
    ap.add_argument("--trans-per-step", type=int, default=0, help="transitions per launch (0 = the workload's default)")
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--cpu-trans", type=int, default=0, help="transitions per CPU-baseline chain (0 = bounded default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the side measurements of the other configurations")
    ap.add_argument("--no-affinity", action="store_true", help="N>1: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--logp", default="fused", choices=["fused", "torch", "torch-graph", "user-source"],
                    help="fused: built-in density inside the kernel; torch / torch-graph: batched torch op around the "
                         "state-machine kernel (host loop / device-driven graph loop); user-source: CUDA source compiled "
                         "into the fused kernel at run time")
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--smem-vecs", type=int, default=-1)
    ap.add_argument("--max-slots", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        run_reference_arm(args)
        return
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
