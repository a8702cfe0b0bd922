#!/bin/bash
# post-tuning cfg2 launch: per-phase clock probe + full ncu capture; deep-tree single-chain latency probe on the funnel
mkdir -p gpurun_out
export LMC_LIB_PATH=$PWD/littlemcmc_b200/liblmc_b200_wtiming.so
echo "== cfg2 bench, phase probe (one line per launch; last ones are post-tuning)"
python bench.py --workload cfg2 --no-cpu --no-configs --steps 2 --warmup 3 2>&1 | grep "warp-timing" | tail -3
echo "== funnel deep trees, fixed eps 1e-3, one chain per SM"
QB_EPS=0.001 python tools/quick_bench.py 148 50 2 0 -1 0 0 funnel 12 20 12 2>&1 | tail -2
echo "== gauss D=100 deep trees, fixed eps 1e-4"
QB_EPS=0.0001 python tools/quick_bench.py 148 100 2 0 -1 0 0 gauss 12 20 12 2>&1 | tail -2
unset LMC_LIB_PATH
QB_EPS=0.001 python tools/quick_bench.py 148 50 2 0 -1 0 0 funnel 12 20 12 2>&1 | tail -1
QB_EPS=0.0001 python tools/quick_bench.py 148 100 2 0 -1 0 0 gauss 12 20 12 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_warp -s 8 -c 1 -f -o gpurun_out/prof_r02v2_cfg2post \
  python bench.py --workload cfg2 --no-cpu --no-configs --steps 3 --warmup 3 > gpurun_out/ncu_cfg2post.log 2>&1
tail -2 gpurun_out/ncu_cfg2post.log
