#!/bin/bash
# cfg2 after tuning: chunk of 4 leaves with the freed shared memory given to tree scratch
b() { python bench.py --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3e  %.3f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
for c in "8 3" "8 4" "4 3" "4 6" "4 10" "4 14" "4 16" "2 16"; do set -- $c
  echo -n "cfg2 chunk $1 smem_vecs $2:  "; b --workload cfg2 --steps 8 --warmup 3 --chunk $1 --smem-vecs $2
done
for c in "8 3" "4 3" "4 8" "4 12"; do set -- $c
  echo -n "tuning-on chunk $1 smem $2: "; python tools/quick_bench.py 1024 100 64 0 $2 0 $1 2>&1 | tail -1 | cut -c1-110
done
for c in "8 3" "8 2" "8 0" "4 3" "4 8"; do set -- $c
  echo -n "cfg4 shape chunk $1 smem $2: "; python tools/quick_bench.py 8192 50 16 0 $2 0 $1 funnel 12 2>&1 | tail -1 | cut -c1-110
done
