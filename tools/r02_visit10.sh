#!/bin/bash
# seed read once per sticky launch (product) and register caps relaxed to what shared memory allows anyway (variant)
b() { python bench.py --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
for v in "" relaxed; do
  if [ -n "$v" ]; then export LMC_LIB_PATH=$PWD/littlemcmc_b200/liblmc_b200_$v.so; else unset LMC_LIB_PATH; fi
  echo "=== variant '${v:-product}'"
  echo -n "cfg2 fused       "; b --workload cfg2 --steps 8 --warmup 3
  echo -n "cfg2 fused       "; b --workload cfg2 --steps 8 --warmup 3
  echo -n "cfg2 user-source "; b --workload cfg2 --logp user-source --steps 8 --warmup 3
  python tools/quick_bench.py 1024 100 64 2>&1 | tail -1
  python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
  python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
  QB_EPS=0.001 python tools/quick_bench.py 148 50 2 0 -1 0 0 funnel 12 20 12 2>&1 | tail -1
done
