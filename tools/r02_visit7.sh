#!/bin/bash
# full ncu capture of the same post-tuning cfg2 launch: built-in Gaussian vs user-source Gaussian
mkdir -p gpurun_out
for m in fused user-source; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampler_warp -s 4 -c 1 -f -o gpurun_out/prof_r02v7_$m \
    python bench.py --workload cfg2 --logp $m --no-cpu --no-configs --steps 3 --warmup 3 > gpurun_out/ncu_v7_$m.log 2>&1
  tail -1 gpurun_out/ncu_v7_$m.log
done
cp ~/.cache/littlemcmc_b200/*.cubin gpurun_out/ 2>/dev/null; ls ~/.cache/littlemcmc_b200/ | head
