#!/bin/bash
# user-source template (branch-free), torch-op densities with fewer kernels, callback advance kernel capped at 128 registers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_user_target_gpu.py tests/test_callback_gpu.py -m gpu -q -x 2>&1 | tail -3
python - <<'P'
import torch, numpy as np, sys
sys.path.insert(0, ".")
import littlemcmc_b200 as lmc
from torch.profiler import profile, ProfilerActivity
for name, t, D in (("gauss", lmc.targets.DiagGaussian(tau=np.linspace(0.5, 2, 100)).torch_batched("cuda:0"), 100),
                   ("funnel", lmc.targets.NealFunnel(50).torch_batched("cuda:0"), 50)):
    q = torch.randn(1024, D, dtype=torch.float64, device="cuda:0")
    for _ in range(20): t(q)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10): lp, g = t(q)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    print(name, "kernels per evaluation:", len(ev) / 10.0)
P
b() { python bench.py --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))"; }
echo "== cfg2 user-source"; b --workload cfg2 --logp user-source --steps 8 --warmup 3; b --workload cfg2 --logp user-source --steps 8 --warmup 3
echo "== cfg2 built-in";  b --workload cfg2 --steps 8 --warmup 3
echo "== cfg2 torch-graph"; b --workload cfg2 --logp torch-graph --steps 4 --warmup 3; b --workload cfg2 --logp torch-graph --steps 4 --warmup 3
echo "== cfg4 torch-graph"; b --workload cfg4 --logp torch-graph --steps 4 --warmup 3
