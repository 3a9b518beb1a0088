#!/bin/bash
mkdir -p gpurun_out
echo "== pytest cycle_multi"; timeout 900 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu 2>&1 | tail -4
echo "== routed local G=8 GROUP=64 lanes=2"; EXP_CYCLES=6 timeout 900 python tools/exp_routed_local.py 8 2 64 34 2>&1 | tail -2
echo "== routed local G=8 GROUP=64 lanes=3 plain"; EXP_CYCLES=6 GPUHASH_LIB=build/lib2p/libgpuhash.so timeout 900 python tools/exp_routed_local.py 8 2 64 34 2>&1 | tail -1
