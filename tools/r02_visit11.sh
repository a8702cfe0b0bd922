#!/bin/bash
# product = seed hoist + relaxed register caps + tau staged in shared memory (NP = 2)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_user_target_gpu.py tests/test_full_size_replay_gpu.py tests/test_edge_cases_gpu.py -m gpu -q 2>&1 | tail -3
b() { python bench.py --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
echo -n "cfg2 fused       "; b --workload cfg2 --steps 8 --warmup 3
echo -n "cfg2 fused       "; b --workload cfg2 --steps 8 --warmup 3
echo -n "cfg2 user-source "; b --workload cfg2 --logp user-source --steps 8 --warmup 3
python tools/quick_bench.py 1024 100 64 2>&1 | tail -1
python tools/quick_bench.py 1024 128 64 2>&1 | tail -1
python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
