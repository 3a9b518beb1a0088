#!/bin/bash
mkdir -p gpurun_out
echo "== golden"; timeout 900 python -m tests.golden.make_ref_golden sparse 2>&1 | tail -4
ls gpurun_out/ref_golden/ | tail -12
echo "== legacy_segments fence lib"
GPUHASH_LIB=$PWD/build/fence/libgpuhash.so timeout 300 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu -k "legacy_segments" 2>&1 | grep -E "^E  |passed|failed" | head -6
echo "== race test"
timeout 900 python -m pytest tests/test_gpu_parity2.py -q -m gpu -k races 2>&1 | grep -E "^E  |passed|failed|^>" | head -12
echo "== envelope numbers"
timeout 600 python tools/dbg_env.py 2>&1 | tail -12
