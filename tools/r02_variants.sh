#!/bin/bash
# Compare kernel variants built with LMC_VARIANT=<name> (littlemcmc_b200/liblmc_b200_<name>.so): tools/r02_variants.sh name1 name2 ...
for v in "" "$@"; do
  if [ -n "$v" ]; then export LMC_LIB_PATH=$PWD/littlemcmc_b200/liblmc_b200_$v.so; else unset LMC_LIB_PATH; fi
  echo "=== variant '${v:-product}'"
  python tools/quick_bench.py 1024 1000 16 0 -1 0 0 2>&1 | tail -1
  python tools/quick_bench.py 1024 1000 64 0 -1 0 0 2>&1 | tail -1
  python tools/quick_bench.py 1024 100 64 0 -1 0 0 2>&1 | tail -1
  python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
done
