#!/bin/bash
# Build in-tree, make sure the library loads, then ship the tree to a B200 box:  tools/gpu.sh <timeout-s> '<command>' [gpurun args]
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py | tail -1
python -c "from littlemcmc_b200 import _lib; _lib.load(); print('library loads')"
T=$1; shift; CMD=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
