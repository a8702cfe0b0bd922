#!/bin/bash
# compute-sanitizer over small parity cases: memcheck (out-of-bounds / misaligned) and racecheck (shared-memory hazards)
# on the register-resident sampler, the lean sampler, callback mode and dense-mass mode.  Outputs under gpurun_out/.
mkdir -p gpurun_out
SEL='test_transition_level_parity and (b1_d10 or static_d100) or test_lean_kernel_parity and (b1_d10 or static_d100) and 128'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" \
    > gpurun_out/sanitize_${tool}_sampler.log 2>&1; echo "$tool sampler rc=$?"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_callback_gpu.py tests/test_dense_gpu.py -m gpu -x -q \
    -k "test_callback_transition_level_parity and b1_d10 or test_dense_transition_level_parity and full_hmc or test_dense_matvec_kernel and 257 or cov_update" \
    > gpurun_out/sanitize_${tool}_modes.log 2>&1; echo "$tool modes rc=$?"
done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitize_*.log
