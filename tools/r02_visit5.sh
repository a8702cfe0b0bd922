#!/bin/bash
# why is the user-source Gaussian 15% slower than the built-in one at cfg2?  grid sizes / durations of both
mkdir -p gpurun_out
for m in fused user-source; do
  timeout 600 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__shared_mem_per_block_static,smsp__inst_executed.sum --clock-control none -k regex:sampler_warp -s 3 -c 2 --csv --log-file gpurun_out/v5_$m.csv \
    python bench.py --workload cfg2 --logp $m --no-cpu --no-configs --steps 2 --warmup 3 > /dev/null 2>&1
  echo "== $m"; grep -v "^==" gpurun_out/v5_$m.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:60], r['Grid Size'], r['Block Size'], r['Metric Name'], r['Metric Value'])"
done
