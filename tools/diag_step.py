"""Scratch diagnostic: per-step device and host times of the headline bench loop."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import littlemcmc_b200 as lmc
from littlemcmc_b200 import _lib as L
C_, D, tps = 1024, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 10
flush_on = (sys.argv[2] != "noflush") if len(sys.argv) > 2 else True
dev = torch.device("cuda", 0)
sigma = 10 ** np.linspace(-0.5, 0.5, D)
target = lmc.targets.DiagGaussian(tau=1 / sigma**2)
pot = lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10)
step = lmc.NUTS(target, D, potential=pot, max_treedepth=10)
ch = step._bind(C_, device=dev, seeds=1_000_003 + np.arange(C_))
step.reset_tuning(); step.iter_count = 0
ch.set_position(np.zeros(D))
trace = torch.empty(C_, tps, D, dtype=torch.float64, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    step._run(tps, 200, trace=trace)
torch.cuda.synchronize()
rows = []
for i in range(30):
    if flush_on:
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    _, st = step._run(tps, 200, trace=trace)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    rows.append((e0.elapsed_time(e1), (t1 - t0) * 1e3, (t2 - t0) * 1e3, float(st[:, :, L.STAT_TREE_SIZE].sum()),
                 float(st[:, :, L.STAT_TREE_SIZE].sum(1).max())))
for r in rows:
    print("dev %.3f ms  host-enqueue %.3f ms  host-total %.3f ms  leapfrogs %d  max-per-chain %d" % r)
