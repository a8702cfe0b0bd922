#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): bench.py at N, cfg5 through distributed.sample (full and chunked gather).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 25 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -2 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]), d.get("allgather"), d.get("dist_e2e"))
except Exception as e:
    print("bench failed", e)
PY
timeout 900 $TR --master-port 29513 tools/run_cfg5.py --gather full --draws 50 2>&1 | tail -1 | cut -c1-1200
timeout 900 $TR --master-port 29515 tools/run_cfg5.py --gather chunks --draws 50 --check 0 2>&1 | tail -1 | cut -c1-1200
