#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err; cat gpurun_out/bench_exp.json; tail -3 gpurun_out/bench_exp.err
for m in fused torch torch-graph; do
  for w in cfg2 cfg4; do
    echo "== $w $m"; timeout 600 python bench.py --no-cpu --workload $w --logp $m --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['config']['mean_tree_depth'])"
  done
done 2>&1 | tee gpurun_out/callback_bench.log
