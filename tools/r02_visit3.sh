#!/bin/bash
# warps of a CTA re-aligned at the head of every transition (sticky launches): does sharing instruction fetches pay?
run() { python bench.py --workload cfg2 --no-cpu --no-configs --steps 8 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step  e2e %.3e  depth %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['run']['mean_tree_depth']))"; }
for v in "" wpb7 wpb7a; do
  if [ -n "$v" ]; then export LMC_LIB_PATH=$PWD/littlemcmc_b200/liblmc_b200_$v.so; else unset LMC_LIB_PATH; fi
  echo "=== variant '${v:-product}'"; run; run
  python tools/quick_bench.py 1024 100 64 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "warp" 2>&1 | tail -2
