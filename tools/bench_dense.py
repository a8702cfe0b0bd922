"""Dense-mass mode measurements (not the contract bench): (1) roofline of lmc_dense_matvec, the HBM-bound operation of
per-chain dense mass matrices (8 n^2 bytes per chain and call, both right-hand sides in one pass); (2) NUTS
leapfrog-steps/s with QuadPotentialFull (one shared matrix) and QuadPotentialFullAdapt (one matrix per chain) on a
correlated Gaussian.  Usage: python tools/bench_dense.py [C] [D]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import littlemcmc_b200 as lmc  # noqa: E402
from littlemcmc_b200 import _lib as L  # noqa: E402
from littlemcmc_b200.targets import TorchBatched  # noqa: E402

Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda", 0)
lib = L.load()
out = {"chains": Cn, "ndim": D}
peak = 6650.0
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])

# ---- (1) the matvec kernel alone ----------------------------------------------------------------------------------
lda = ld = D + (D & 1)
A = torch.randn(Cn, D, lda, dtype=torch.float64, device=dev)
x = torch.randn(Cn, 2, ld, dtype=torch.float64, device=dev)
y = torch.empty_like(x)
p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for nrhs in (2,):
    for _ in range(3):
        L.check(lib.lmc_dense_matvec(None, Cn, p(A), D * lda, lda, D, ld, p(x), p(y), nrhs, stream), "mv")
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.lmc_dense_matvec(None, Cn, p(A), D * lda, lda, D, ld, p(x), p(y), nrhs, stream), "mv")
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    nbytes = Cn * D * D * 8                       # algorithmic: every matrix element once
    out["matvec"] = {"ms": ms, "achieved_gbs": nbytes / ms / 1e6, "peak_gbs": peak, "frac": nbytes / ms / 1e6 / peak,
                     "bytes": nbytes, "nrhs": nrhs}
    print("lmc_dense_matvec  C=%d D=%d nrhs=%d: %.3f ms, %.0f GB/s (%.2f of %.0f GB/s measured HBM peak)"
          % (Cn, D, nrhs, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak), flush=True)
    want = torch.einsum("cij,crj->cri", A[:8, :, :D], x[:8, :, :D])
    assert torch.allclose(y[:8, :, :D], want, rtol=1e-11, atol=1e-9)
del A, x, y
torch.cuda.empty_cache()

# ---- (2) NUTS throughput in dense mode -------------------------------------------------------------------------------
rs = np.random.RandomState(0)
qm, _ = np.linalg.qr(rs.randn(D, D))
ev = 10 ** np.linspace(-0.5, 0.5, D)
prec = (qm * (1 / ev**2)) @ qm.T
prec = 0.5 * (prec + prec.T)
cov = np.linalg.inv(prec)
P = torch.as_tensor(prec, device=dev)


def fn(q):
    g = -(q @ P)
    return 0.5 * (q * g).sum(1), g


target = TorchBatched(fn)
for name, mk, n_trans, tune in (("QuadPotentialFull (shared matrix)", lambda: lmc.QuadPotentialFull(cov), 30, 20),
                                ("QuadPotentialFullAdapt (per-chain matrix)",
                                 lambda: lmc.QuadPotentialFullAdapt(D, np.zeros(D), np.eye(D), 10), 12, 12)):
    chains = Cn if "shared" in name else min(Cn, 256)
    # warm-up call: library handles (cuBLAS, cuSOLVER), allocator pools, first-launch costs
    lmc.sample(target, D, draws=0, tune=2, step=lmc.NUTS(target, D, potential=mk(), max_treedepth=8), chains=chains,
               start=np.zeros(D), random_seed=list(range(chains)), discard_tuned_samples=False, return_device=True)
    step = lmc.NUTS(target, D, potential=mk(), max_treedepth=8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr, st = lmc.sample(target, D, draws=n_trans - tune, tune=tune, step=step, chains=chains, start=np.zeros(D),
                        random_seed=list(range(chains)), discard_tuned_samples=False, return_device=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    leap = float(st["tree_size"].sum())
    out[name] = {"chains": chains, "transitions": n_trans, "leapfrogs": leap, "seconds": dt, "leapfrog_per_s": leap / dt}
    print("%s: %d chains x %d transitions, %d leapfrogs in %.2f s -> %.3e leapfrog/s (mean depth %.2f)"
          % (name, chains, n_trans, leap, dt, leap / dt, float(st["depth"].double().mean())), flush=True)
    del step, tr, st
    torch.cuda.empty_cache()
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dense_bench.json", "w"))
