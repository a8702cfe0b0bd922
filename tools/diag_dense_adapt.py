"""Scratch diagnostic: where a dense-mode tuning transition with per-chain adapted matrices (QuadPotentialFullAdapt)
spends its time, and batched Cholesky variants on that shape.  Usage: python tools/diag_dense_adapt.py [C] [D]"""
import collections
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import littlemcmc_b200 as lmc  # noqa: E402
from littlemcmc_b200 import engine, quadpotential_dense as qd  # noqa: E402
from littlemcmc_b200.targets import TorchBatched  # noqa: E402

Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda", 0)


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(Cn, D, D + 50, dtype=torch.float64, device=dev, generator=g)
A = X @ X.mT / (D + 50) + 0.1 * torch.eye(D, dtype=torch.float64, device=dev)
del X
ms, ref = timeit(lambda: torch.linalg.cholesky_ex(A)[0], 2)
print("torch.linalg.cholesky_ex [%d,%d,%d]: %.1f ms" % (Cn, D, D, ms), flush=True)
if hasattr(qd, "batched_cholesky"):
    for nb in (64, 128, 256):
        ms, (Lb, info) = timeit(lambda: qd.batched_cholesky(A, nb), 2)
        err = float(((Lb - ref).abs().amax() / ref.abs().amax()).item())
        print("blocked nb=%d: %.1f ms, max rel diff vs cuSOLVER %.2e, info max %d" % (nb, ms, err, int(info.max())), flush=True)
del A, ref
torch.cuda.empty_cache()

rs = np.random.RandomState(0)
qm, _ = np.linalg.qr(rs.randn(D, D))
ev = 10 ** np.linspace(-0.5, 0.5, D)
prec = (qm * (1 / ev**2)) @ qm.T
prec = 0.5 * (prec + prec.T)
P = torch.as_tensor(prec, device=dev)
target = TorchBatched(lambda q: ((lambda gg: (0.5 * (q * gg).sum(1), gg))(-(q @ P))))
T = collections.Counter()
N = collections.Counter()


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        T[name] += time.perf_counter() - t0
        N[name] += 1
        return r
    return w


pot = lmc.QuadPotentialFullAdapt(D, np.zeros(D), np.eye(D), 10)
pot._velocity_rows = timed("velocity", pot._velocity_rows)
pot._momentum_rows = timed("momentum", pot._momentum_rows)
pot._update_rows = timed("update", pot._update_rows)
engine.evaluate_callback = timed("gradient", engine.evaluate_callback)
step = lmc.NUTS(target, D, potential=pot, max_treedepth=8)
kw = dict(step=step, chains=Cn, start=np.zeros(D), random_seed=list(range(Cn)), discard_tuned_samples=False,
          return_device=True)
lmc.sample(target, D, draws=0, tune=2, **kw)
T.clear()
N.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
tr, st = lmc.sample(target, D, draws=0, tune=8, **kw)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print("total %.3f s, leapfrogs %d -> %.3e /s (with per-call synchronisation for the breakdown)"
      % (tot, int(st["tree_size"].sum()), float(st["tree_size"].sum()) / tot))
print({k: (round(v, 3), N[k]) for k, v in T.items()}, "other", round(tot - sum(T.values()), 3))
