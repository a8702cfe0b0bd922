#!/bin/bash
# Round-2 profile visit: the default bench line (both arms), launch list of the default bench, full ncu captures of the
# fused kernels on bench-shaped launches.  Outputs under gpurun_out/ (summarised into profiles/ by tools/ncu_summary.py).
# Usage: TAG=r02z bash tools/r02_profiles.sh
TAG=${TAG:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_lean -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_headline \
  python tools/quick_bench.py 1024 1000 16 > gpurun_out/ncu_headline.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_cta -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_cta \
  python tools/quick_bench.py 1024 1000 16 2 > gpurun_out/ncu_cta.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_warp -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_cfg2 \
  python tools/quick_bench.py 1024 100 64 > gpurun_out/ncu_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_warp -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_cfg4 \
  python tools/quick_bench.py 8192 50 40 0 -1 0 0 funnel 12 > gpurun_out/ncu_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cb_advance -s 50 -c 1 -f -o gpurun_out/prof_${TAG}_cb \
  python bench.py --no-cpu --no-configs --workload cfg2 --logp torch --steps 1 --warmup 1 > gpurun_out/ncu_cb.log 2>&1
timeout 600 python tools/bench_dense.py 1024 1000 2>&1 | grep -v "samples in chain\|Tuning was" | tail -4 | cut -c1-200
cp gpurun_out/dense_bench.json gpurun_out/${TAG}_dense_bench.json
tail -1 gpurun_out/ncu_cfg4.log
