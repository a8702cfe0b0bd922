"""Scratch throughput probe (not the contract bench): NUTS with in-kernel Philox, sweeping launch knobs.
Usage: python tools/quick_bench.py C D n_trans [groups] [smems] [slots] [chunks] [target=gauss|funnel] [depth] [warm] [early_depth]
(lists are comma separated; group 0 = library default, 1 = chunked warp kernel, >= 32 register kernel, < 0 lean)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from littlemcmc_b200 import _lib as L  # noqa: E402
from littlemcmc_b200 import engine  # noqa: E402

arg = sys.argv[1:]
C = int(arg[0]) if len(arg) > 0 else 1024
D = int(arg[1]) if len(arg) > 1 else 1000
T = int(arg[2]) if len(arg) > 2 else 40
groups = [int(x) for x in arg[3].split(",")] if len(arg) > 3 else [0]
smems = [int(x) for x in arg[4].split(",")] if len(arg) > 4 else [-1]
slots = [int(x) for x in arg[5].split(",")] if len(arg) > 5 else [0]
chunks = [int(x) for x in arg[6].split(",")] if len(arg) > 6 else [0]
target = arg[7] if len(arg) > 7 else "gauss"
depth = int(arg[8]) if len(arg) > 8 else 10
warm = int(arg[9]) if len(arg) > 9 else 100
early = int(arg[10]) if len(arg) > 10 else 8   # early_max_treedepth (the cap while iteration < 200)
# QB_EPS=<step size>: every chain integrates with this fixed step size in the timed launch (step_size_override); a tiny
# one makes every tree reach max_treedepth without a U-turn -- the single-chain latency probe for deep trees
import os  # noqa: E402
fixed_eps = float(os.environ["QB_EPS"]) if os.environ.get("QB_EPS") else None

dev = "cuda:0"
if target == "funnel":
    tgt = engine.FusedTarget(L.TARGET_FUNNEL, D, v_scale=3.0)
else:
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    tgt = engine.FusedTarget(L.TARGET_DIAG_GAUSSIAN, D, tau=1 / sigma**2)
# QB_ADAPT_MASS / QB_ADAPT_STEP = 0 switch one adaptation off in the TIMED launch; QB_TUNE = its n_tune (0: tuning is over)
params = dict(adapt_mass=1, adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10, Emax=1000.0,
              max_treedepth=depth, early_max_treedepth=early)
params_timed = dict(params, adapt_mass=int(os.environ.get("QB_ADAPT_MASS", "1")),
                    adapt_step_size=int(os.environ.get("QB_ADAPT_STEP", "1")))
seeds = engine.seeds_tensor(np.arange(C) + 12345, dev)
for g in groups:
    for sm in smems:
        for sl in slots:
            for ck in chunks:
                ch = engine.DeviceChains(C, D, dev)
                ch.reset_potential(np.ones(D), np.zeros(D), 10.0, 101)
                ch.reset_step_adapt(0.25 / D**0.25)
                ch.set_position(np.zeros(D))
                knobs = dict(group=g, smem_vecs=sm, max_slots=sl, chunk=ck)
                try:
                    # warm-up: tuning transitions so step size / mass matrix are roughly adapted
                    engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=warm, iter0=0, n_tune=10**9, params=params,
                                           seeds=seeds, knobs=knobs)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ov = None if fixed_eps is None else torch.full((C,), fixed_eps, dtype=torch.float64, device=dev)
                    tr, st = engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=T, iter0=warm,
                                                    n_tune=int(os.environ.get("QB_TUNE", 10**9)),
                                                    params=params_timed, seeds=seeds, knobs=knobs, step_size_override=ov)
                    e1.record()
                    torch.cuda.synchronize()
                except Exception as e:  # unsupported shape
                    print("C=%d D=%d group=%d smem=%d slots=%d chunk=%d: %s" % (C, D, g, sm, sl, ck, e), flush=True)
                    continue
                ms = e0.elapsed_time(e1)
                steps = float(st[:, :, L.STAT_TREE_SIZE].sum())
                print("C=%d D=%d T=%d %s group=%d smem=%d slots=%d chunk=%d: %.2f ms  %.3e leapfrog/s  (depth %.2f max %d, "
                      "accept %.3f, div %d, alg %.1f GB/s = %.3f of 6457)"
                      % (C, D, T, target, g, sm, sl, ck, ms, steps / ms * 1e3, float(st[:, :, L.STAT_DEPTH].mean()),
                         int(st[:, :, L.STAT_DEPTH].max()), float(st[:, :, L.STAT_ACCEPT].mean()),
                         int(st[:, :, L.STAT_DIVERGING].sum()), steps / ms * 1e3 * 48 * D / 1e9,
                         steps / ms * 1e3 * 48 * D / 1e9 / 6457.4), flush=True)
