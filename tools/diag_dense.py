"""Scratch diagnostic: where a dense-mode iteration spends its time (host loop of engine.DenseRun)."""
import sys, time, collections
import numpy as np, torch
sys.path.insert(0, ".")
import littlemcmc_b200 as lmc
from littlemcmc_b200 import engine
from littlemcmc_b200.targets import TorchBatched
Cn, D = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
rs = np.random.RandomState(0)
qm, _ = np.linalg.qr(rs.randn(D, D))
ev = 10 ** np.linspace(-0.5, 0.5, D)
prec = (qm * (1 / ev**2)) @ qm.T; prec = 0.5 * (prec + prec.T)
cov = np.linalg.inv(prec)
P = torch.as_tensor(prec, device=dev)
target = TorchBatched(lambda q: ((lambda g: (0.5 * (q * g).sum(1), g))(-(q @ P))))
T = collections.Counter()
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T[name] += time.perf_counter() - t0
        return r
    return w
pot = lmc.QuadPotentialFull(cov)
pot._velocity_rows = timed("velocity", pot._velocity_rows)
pot._momentum_rows = timed("momentum", pot._momentum_rows)
engine.evaluate_callback = timed("gradient", engine.evaluate_callback)
step = lmc.NUTS(target, D, potential=pot, max_treedepth=8)
lmc.sample(target, D, draws=1, tune=2, step=step, chains=Cn, start=np.zeros(D), random_seed=list(range(Cn)),
           discard_tuned_samples=False, return_device=True)      # warm-up: library page-in, cuBLAS / cuSOLVER handles
T.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
tr, st = lmc.sample(target, D, draws=5, tune=10, step=step, chains=Cn, start=np.zeros(D), random_seed=list(range(Cn)),
                    discard_tuned_samples=False, return_device=True)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print("total %.3f s, leapfrogs %d -> %.3e /s" % (tot, int(st["tree_size"].sum()), float(st["tree_size"].sum()) / tot))
print({k: round(v, 3) for k, v in T.items()}, "other", round(tot - sum(T.values()), 3))
