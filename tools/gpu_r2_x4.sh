#!/bin/bash
mkdir -p gpurun_out
for sh in 16x4 16x8 24x8; do
GPUHASH_XCHG_SHAPE=$sh timeout 600 python tools/exp_xchg_local.py 1 64 34 12 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['env'], d['mixed_us_per_rank_step'], d['search_us_per_rank_step'], d['alone_us'], d['mismatches'])"
done
GPUHASH_XCHG_SHAPE=24x8 timeout 900 python -m pytest tests/test_gpu_xchg.py -x -q -m gpu 2>&1 | tail -3
