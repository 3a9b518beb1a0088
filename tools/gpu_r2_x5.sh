#!/bin/bash
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu -k "legacy_segments" 2>&1 | grep -E "^E|passed|failed" | head -8; done
echo "== fence lib"
for i in 1 2; do GPUHASH_LIB=$PWD/build/fence/libgpuhash.so timeout 300 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu -k "legacy_segments" 2>&1 | grep -E "^E|passed|failed" | head -8; done
echo "== zipf"
timeout 300 python -m pytest tests/test_zipf.py -x -q -m gpu 2>&1 | tail -3
