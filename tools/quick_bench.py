"""Scratch throughput probe (not the contract bench): NUTS on a diagonal Gaussian with in-kernel Philox, sweeping
launch knobs.  Usage: python tools/quick_bench.py [C] [D] [n_trans] [group,group,...] [smem,smem,...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from littlemcmc_b200 import _lib as L  # noqa: E402
from littlemcmc_b200 import engine  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 40
groups = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
smems = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [-1]
slots = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [0]

dev = "cuda:0"
sigma = 10 ** np.linspace(-0.5, 0.5, D)
tgt = engine.FusedTarget(L.TARGET_DIAG_GAUSSIAN, D, tau=1 / sigma**2)
params = dict(adapt_mass=1, adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10, Emax=1000.0,
              max_treedepth=10, early_max_treedepth=8)
seeds = engine.seeds_tensor(np.arange(C) + 12345, dev)
for g in groups:
    for sm in smems:
        for sl in slots:
            ch = engine.DeviceChains(C, D, dev)
            ch.reset_potential(np.ones(D), np.zeros(D), 10.0, 101)
            ch.reset_step_adapt(0.25 / D**0.25)
            ch.set_position(np.zeros(D))
            knobs = dict(group=g, smem_vecs=sm, max_slots=sl)
            # warm-up: 100 tuning transitions so step size / mass matrix are roughly adapted
            engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=100, iter0=0, n_tune=10**9, params=params,
                                   seeds=seeds, knobs=knobs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tr, st = engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=T, iter0=100, n_tune=10**9, params=params,
                                            seeds=seeds, knobs=knobs)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            steps = float(st[:, :, L.STAT_TREE_SIZE].sum())
            depth = float(st[:, :, L.STAT_DEPTH].mean())
            acc = float(st[:, :, L.STAT_ACCEPT].mean())
            print("C=%d D=%d T=%d group=%d smem=%d slots=%d: %.2f ms  %.3e leapfrog/s  (mean depth %.2f, accept %.3f, "
                  "alg %.1f GB/s)" % (C, D, T, g, sm, sl, ms, steps / ms * 1e3, depth, acc,
                                     steps / ms * 1e3 * 48 * D / 1e9), flush=True)
