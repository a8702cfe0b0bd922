#!/bin/bash
# torch-op Gaussian with two instead of three kernels per evaluation
for v in mv bmm; do
  export LMC_TORCH_GAUSS=$v
  echo "== $v"
  python - <<'P'
import os, torch, numpy as np, sys
sys.path.insert(0, ".")
import littlemcmc_b200 as lmc
t = lmc.targets.DiagGaussian(tau=np.linspace(0.5, 2, 100)).torch_batched("cuda:0")
q = torch.randn(1024, 100, dtype=torch.float64, device="cuda:0")
for _ in range(20): t(q)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): lp, g = t(q)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
print("kernels per evaluation:", len(ev) / 10.0, sorted({e.name[:60] for e in ev}))
ref = 0.5 * (q.cpu().numpy() * (-(np.linspace(0.5, 2, 100) * q.cpu().numpy()))).sum(1)
print("max abs err logp", float(np.abs(lp.cpu().numpy() - ref).max()))
P
  for i in 1 2; do python bench.py --no-cpu --no-configs --workload cfg2 --logp torch-graph --steps 4 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))"; done
done
unset LMC_TORCH_GAUSS
echo "== user-source Gaussian, logp in terms of g"
for i in 1 2; do python bench.py --no-cpu --no-configs --workload cfg2 --logp user-source --steps 8 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step' % (d['value'], d['ms_per_step']))"; done
python bench.py --no-cpu --no-configs --workload cfg2 --steps 8 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('built-in %.3e  %.3f ms/step' % (d['value'], d['ms_per_step']))"
timeout 600 python -m pytest tests/test_user_target_gpu.py -m gpu -q 2>&1 | tail -3
