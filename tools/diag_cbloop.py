"""Where the time of a device-driven callback run goes: capture / graph create (instantiate) / loop execution.
Usage: python tools/diag_cbloop.py [C] [D] [n_trans] [iters_per_body]"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import littlemcmc_b200 as lmc  # noqa: E402
from littlemcmc_b200 import _lib as L, engine  # noqa: E402

Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 100
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
K = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
sigma = 10 ** np.linspace(-0.5, 0.5, D)
cb = lmc.targets.DiagGaussian(sigma=sigma).torch_batched(dev, cuda_graph=True)
params = dict(adapt_mass=1, adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10, Emax=1000.0,
              max_treedepth=10, early_max_treedepth=8)
seeds = engine.seeds_tensor(np.arange(Cn) + 7, dev)
ch = engine.DeviceChains(Cn, D, dev)
ch.reset_potential(np.ones(D), np.zeros(D), 10.0, 101)
ch.reset_step_adapt(0.25 / D ** 0.25)
ch.set_position(np.zeros(D))
engine.run_transitions_callback(L.KIND_NUTS, ch, cb, n_trans=40, iter0=0, n_tune=10**9, params=params, seeds=seeds)
torch.cuda.synchronize()
for rep in range(3):
    run = engine.CallbackRun(L.KIND_NUTS, ch, cb, n_trans=T, iter0=40 + rep * T, n_tune=10**9, params=params, seeds=seeds)
    t0 = time.perf_counter()
    run.begin()
    graph = run._capture_iterations(K, keep_graph=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    iters = torch.zeros(1, dtype=torch.int32, device=dev)
    loop = C.c_void_p()
    L.check(run.lib.lmc_callback_loop_create(C.c_void_p(int(graph.raw_cuda_graph())), engine._ptr(run.n_running),
                                             engine._ptr(iters), run.max_iters, C.byref(loop)), "create")
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(run.lib.lmc_callback_loop_launch(loop, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "launch")
    e1.record()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    n_it = int(iters.item())
    leap = float(run.stats[:, :, L.STAT_TREE_SIZE].sum())
    print("capture %.1f ms  create %.1f ms  loop %.2f ms (events %.2f) bodies %d x %d iters -> %.1f us/iteration, %.3e leapfrog/s "
          "(loop only), left %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, e0.elapsed_time(e1), n_it, K,
                                    e0.elapsed_time(e1) * 1e3 / max(1, n_it * K), leap / (e0.elapsed_time(e1) * 1e-3),
                                    int(run.n_running.item())), flush=True)
    run.lib.lmc_callback_loop_destroy(loop)

# ---- the whole public call, as bench.py times it (CUDA events around run_transitions_callback) --------------------------
import threading  # noqa: E402


def timed_call(tag):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        tr, st = engine.run_transitions_callback(L.KIND_NUTS, ch, cb, n_trans=T, iter0=200 + rep * T, n_tune=10**9,
                                                 params=params, seeds=seeds, cuda_graph=True)
        e1.record()
        torch.cuda.synchronize()
        print("%s: whole call %.1f ms wall, %.1f ms events" % (tag, (time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1)),
              flush=True)


timed_call("plain")
stop = [False]


def poll():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop[0]:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(0.01)


th = threading.Thread(target=poll, daemon=True)
th.start()
timed_call("with NVML polling")
stop[0] = True
