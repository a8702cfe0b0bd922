#!/bin/bash
# full GPU suite + smoke on the product library, then pop-ahead variant vs product on FIFO shapes, chunk sweep on the funnel
bash tools/gpu_tests.sh
for v in "" popahead; do
  if [ -n "$v" ]; then export LMC_LIB_PATH=$PWD/littlemcmc_b200/liblmc_b200_$v.so; else unset LMC_LIB_PATH; fi
  echo "=== variant '${v:-product}'"
  python tools/quick_bench.py 8192 50 16 0 -1 0 4,8 funnel 12 2>&1 | tail -2
  python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
  python tools/quick_bench.py 4096 100 32 0 -1 0 0 2>&1 | tail -1
  python tools/quick_bench.py 1024 100 64 0 -1 0 0 2>&1 | tail -1
done
if [ -n "$LMC_LIB_PATH" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "warp or full" 2>&1 | tail -2; fi
