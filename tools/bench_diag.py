"""Roofline of lmc_chain_moments (diagnostics): one pass over a [chains, draws, ndim] float64 trace."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from littlemcmc_b200 import diagnostics as dg  # noqa: E402

Cn, T, D = (int(a) for a in (sys.argv[1:4] + ["1024", "400", "1000"][len(sys.argv) - 1:]))
peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]) if os.path.exists("MEASURED_PEAKS.json") else 6650.0
x = torch.randn(Cn, T, D, dtype=torch.float64, device="cuda")
for n_seg in (2, 4):
    for _ in range(3):
        dg.chain_moments(x, n_seg)
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dg.chain_moments(x, n_seg)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    gbs = x.numel() * 8 / t / 1e6
    print(json.dumps({"kernel": "lmc_chain_moments", "chains": Cn, "draws": T, "ndim": D, "n_seg": n_seg, "ms": t,
                      "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak}))
