#!/bin/bash
# compute-sanitizer memcheck (+ racecheck on one case) over the chunked warp kernel (sticky and FIFO, NP = 1 and 2: staged
# tau), the user-source path and the callback advance kernel, small parity cases; every command under its own timeout
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "(test_transition_level_parity or test_warp_kernel_parity) and (static_d100 or b1_d10 or diag_d37)" \
  > gpurun_out/sanitize_memcheck_warp.log 2>&1; echo "memcheck warp rc=$?"
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_callback_gpu.py tests/test_user_target_gpu.py -m gpu -x -q \
  -k "test_callback_transition_level_parity and b1_d10 or test_user_gaussian_logp_in_terms_of_the_gradient" \
  > gpurun_out/sanitize_memcheck_cb_user.log 2>&1; echo "memcheck callback + user rc=$?"
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "test_warp_kernel_parity and static_d100 and 8" > gpurun_out/sanitize_racecheck_warp.log 2>&1; echo "racecheck warp rc=$?"
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitize_memcheck_warp.log gpurun_out/sanitize_memcheck_cb_user.log gpurun_out/sanitize_racecheck_warp.log
