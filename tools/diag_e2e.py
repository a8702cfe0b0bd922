"""Scratch diagnostic: where the end-to-end time of sample() goes (kernel only vs host trace, block sizes)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import littlemcmc_b200 as lmc
C_, D, T = 1024, 1000, 400
dev = torch.device("cuda", 0)
sigma = 10 ** np.linspace(-0.5, 0.5, D)
target = lmc.targets.DiagGaussian(tau=1 / sigma**2)
start = torch.zeros(C_, D, dtype=torch.float64).pin_memory()
def mk():
    pot = lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10)
    return lmc.NUTS(target, D, potential=pot, max_treedepth=10)
step = mk()
TM = {}
def call(block, ret_dev, seed, hw="copy"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr, st = lmc.sample(target, D, draws=T // 2, tune=T // 2, step=step, chains=C_, start=start.numpy(),
                        random_seed=list(1000 + seed + np.arange(C_)), discard_tuned_samples=False, device=dev,
                        progressbar=False, block=block, return_device=ret_dev, host_write=hw, _timing=TM)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = float(st["tree_size"].sum())
    return dt, n
call(None, False, 0); call(None, True, 0)
for block in (None,):
    for mode in ("device", "copy", "direct"):
        dt, n = min(call(block, mode == "device", s, mode if mode != "device" else "copy") for s in (1, 2, 3))
        TM.clear(); call(block, mode == "device", 9, mode if mode != "device" else "copy")
        print("   ", {k: round(v, 1) for k, v in TM.items()})
        print("block=%s %s: %.1f ms, %.3e leapfrog/s (%d leapfrogs)" % (block, mode, dt * 1e3, n / dt, n), flush=True)
