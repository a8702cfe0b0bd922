#!/bin/bash
# kernel experiments: parity first, then throughput of the product library and of any variant libraries
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
echo "== product"; timeout 300 python tools/quick_bench.py 1024 1000 40 0,256
for v in littlemcmc_b200/liblmc_b200_*.so; do
  [ -e "$v" ] || continue
  echo "== $v"; LMC_LIB_PATH=$PWD/$v timeout 300 python tools/quick_bench.py 1024 1000 40 0
done
} > gpurun_out/exp.log 2>&1
cat gpurun_out/exp.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err; cat gpurun_out/bench_exp.json
