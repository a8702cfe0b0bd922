#!/bin/bash
# Round-2 baseline: every BASELINE config through the fused kernels as they stood at the end of round 1, plus one
# full ncu capture of a one-warp-per-chain shape (cfg2: 1024 x 100).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for w in headline cfg2 cfg3 cfg4; do
  timeout 600 python bench.py --no-cpu --workload $w --steps 10 --warmup 3 > gpurun_out/r02a_$w.json 2> gpurun_out/r02a_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02a_$w.json").readline())
    print("$w", "%.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], "depth %.2f" % d["config"]["mean_tree_depth"])
except Exception as e:
    print("$w failed", e)
PY
done
for sh in "1024 100 16" "8192 50 8" "1024 100 16 64"; do
  timeout 300 python tools/quick_bench.py $sh 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_ -s 1 -c 1 -f -o gpurun_out/prof_cfg2 \
  python tools/quick_bench.py 1024 100 16 > gpurun_out/ncu_cfg2.log 2>&1
tail -1 gpurun_out/ncu_cfg2.log
