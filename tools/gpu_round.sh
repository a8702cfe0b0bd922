#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines (both arms), ncu launch list, one full ncu capture of the sampler
# kernel on a bench-shaped launch, dense-mode and diagnostics kernel rooflines.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
cut -c1-200 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sampler_ -s 1 -c 1 -f -o gpurun_out/prof_sampler \
  python tools/quick_bench.py 1024 1000 16 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
timeout 600 python tools/bench_dense.py 1024 1000 2>&1 | grep -v "samples in chain" | tail -4 | cut -c1-160
timeout 300 python tools/bench_diag.py > gpurun_out/diag_bench.jsonl 2>&1; cat gpurun_out/diag_bench.jsonl | cut -c1-200
