#!/bin/bash
# FIFO launches: a popped chain stays on its warp for up to 4 transitions
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_gpu.py tests/test_full_size_replay_gpu.py tests/test_user_target_gpu.py tests/test_api_gpu.py -m gpu -q 2>&1 | tail -3
python tools/quick_bench.py 8192 50 16 0 -1 0 0 funnel 12 2>&1 | tail -1
python tools/quick_bench.py 8192 50 40 0 -1 0 0 funnel 12 2>&1 | tail -1
python tools/quick_bench.py 4096 100 32 2>&1 | tail -1
python tools/quick_bench.py 2048 100 64 2>&1 | tail -1
python bench.py --no-cpu --no-configs --workload cfg4 --steps 7 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('cfg4 %.3e  %.3f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['run'])"
