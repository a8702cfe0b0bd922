#!/bin/bash
# what makes a tuning transition twice as expensive as a sampling transition at 1024 x 100?
q() { python tools/quick_bench.py 1024 100 64 2>&1 | tail -1 | sed 's/group=0 smem=-1 slots=0 chunk=0//'; }
for v in "" wpb7a; do
  if [ -n "$v" ]; then export LMC_LIB_PATH=$PWD/littlemcmc_b200/liblmc_b200_$v.so; else unset LMC_LIB_PATH; fi
  echo "=== variant '${v:-product}'"
  echo -n "tuning, both adaptations : "; q
  echo -n "tuning, no mass adapt    : "; QB_ADAPT_MASS=0 q
  echo -n "tuning, no step adapt    : "; QB_ADAPT_STEP=0 q
  echo -n "tuning, neither          : "; QB_ADAPT_MASS=0 QB_ADAPT_STEP=0 q
  echo -n "tuning over (n_tune=0)   : "; QB_TUNE=0 q
done
