#!/bin/bash
# Round-2 profile visit: launch list of the default bench, full ncu captures of the three fused kernels on bench-shaped
# launches, the bench line itself.  Outputs under gpurun_out/ (summarised into profiles/ by tools/ncu_summary.py).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02z_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_lean -s 1 -c 1 -f -o gpurun_out/prof_r02z_headline \
  python tools/quick_bench.py 1024 1000 16 > gpurun_out/ncu_headline.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_warp -s 1 -c 1 -f -o gpurun_out/prof_r02z_cfg2 \
  python tools/quick_bench.py 1024 100 64 > gpurun_out/ncu_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_warp -s 1 -c 1 -f -o gpurun_out/prof_r02z_cfg4 \
  python tools/quick_bench.py 8192 50 40 0 -1 0 0 funnel 12 > gpurun_out/ncu_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cb_advance -s 50 -c 1 -f -o gpurun_out/prof_r02z_cb \
  python bench.py --no-cpu --no-configs --workload cfg2 --logp torch --steps 1 --warmup 1 > gpurun_out/ncu_cb.log 2>&1
tail -1 gpurun_out/ncu_cfg4.log
