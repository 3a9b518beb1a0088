#!/bin/bash
mkdir -p gpurun_out
echo "== pytest cycle_multi"; timeout 900 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu 2>&1 | tail -4
echo "== routed local, LTC64B"; EXP_CYCLES=16 timeout 600 python tools/exp_routed_local.py 8 4 8 34 2>&1 | tail -2
echo "== routed local, LTC64B, 8 lanes"; EXP_CYCLES=16 EXP_SKIP_PARTS=1 timeout 600 python tools/exp_routed_local.py 8 8 8 34 2>&1 | tail -1
echo "== routed local, plain loads"; EXP_CYCLES=16 GPUHASH_LIB=build/lib2p/libgpuhash.so timeout 600 python tools/exp_routed_local.py 8 4 8 34 2>&1 | tail -1
