#!/bin/bash
# multi-transition units (warp + lean kernels), every command under its own timeout
timeout 60 python tools/quick_bench.py 64 50 16 0 -1 8 0 funnel 8 10 > gpurun_out/v14_a.log 2>&1 || { echo "FIFO warp sanity run failed / hung: stop"; tail -3 gpurun_out/v14_a.log; exit 1; }
tail -1 gpurun_out/v14_a.log
timeout 60 python tools/quick_bench.py 1024 1000 16 > gpurun_out/v14_b.log 2>&1 || { echo "lean sanity run failed / hung: stop"; tail -3 gpurun_out/v14_b.log; exit 1; }
tail -1 gpurun_out/v14_b.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_full_size_replay_gpu.py tests/test_api_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -3
for i in 1 2; do timeout 60 python tools/quick_bench.py 1024 1000 16 2>&1 | tail -1; done
timeout 60 python tools/quick_bench.py 1024 1000 64 2>&1 | tail -1
timeout 60 python tools/quick_bench.py 4096 1000 16 2>&1 | tail -1
timeout 60 python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
timeout 60 python tools/quick_bench.py 8192 50 40 0 -1 0 0 funnel 12 2>&1 | tail -1
timeout 60 python tools/quick_bench.py 4096 100 32 2>&1 | tail -1
timeout 200 python bench.py --no-cpu --no-configs --steps 25 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('headline %.3e  %.3f ms/step  e2e %.3e frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
timeout 200 python bench.py --no-cpu --no-configs --workload cfg4 --steps 7 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('cfg4 %.3e  %.3f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['run']['ms_per_step_median'])"
