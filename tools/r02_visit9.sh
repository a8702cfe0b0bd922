#!/bin/bash
bash tools/gpu_tests.sh
b() { python bench.py --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step launches %d  e2e %.3e' % (d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value']))"; }
echo "== cfg2 torch-graph"; b --workload cfg2 --logp torch-graph --steps 4 --warmup 3; b --workload cfg2 --logp torch-graph --steps 4 --warmup 3
echo "== cfg4 torch-graph"; b --workload cfg4 --logp torch-graph --steps 4 --warmup 3
echo "== cfg2 fused"; b --workload cfg2 --steps 8 --warmup 3
echo "== quick_bench tuning on"; python tools/quick_bench.py 1024 100 64 2>&1 | tail -1; python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
