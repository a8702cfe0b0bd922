#!/bin/bash
mkdir -p gpurun_out
{
python tools/quick_bench.py 1024 100 16 1 -1,0 0 4,8
python tools/quick_bench.py 1024 100 64 1 -1 0 8
python tools/quick_bench.py 1024 100 16 1 -1 600 8
python tools/quick_bench.py 8192 50 16 1 -1,0 0 4,8 funnel 12
python tools/quick_bench.py 8192 50 40 1 -1 0 4,8 funnel 12
python tools/quick_bench.py 8192 50 16 1 -1 0 4,8 gauss 10
} 2>&1 | tee gpurun_out/r02c_warp_bench.log
