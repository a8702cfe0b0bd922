#!/bin/bash
# GPU test suite (+ smoke) on one box; log under gpurun_out/.  Usage: gpu_tests.sh [pytest args]
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=10 "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
