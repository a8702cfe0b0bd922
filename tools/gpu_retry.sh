#!/bin/bash
# tools/gpu_retry.sh <timeout-s> '<command>' [gpurun args]: like tools/gpu.sh, retrying while the pod answers busy (rc 3).
cd "$(dirname "$0")/.."
for i in $(seq 1 20); do
  tools/gpu.sh "$@" > gpurun_out/.retry.out 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" gpurun_out/.retry.out; then sleep 90; continue; fi
  break
done
cat gpurun_out/.retry.out
