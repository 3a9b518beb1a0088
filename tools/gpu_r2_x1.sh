#!/bin/bash
mkdir -p gpurun_out
echo "== xchg tests"; timeout 900 python -m pytest tests/test_gpu_xchg.py -x -q -m gpu 2>&1 | tail -15
echo "== xchg local G=8"; timeout 600 python tools/exp_xchg_local.py 8 64 34 12 2>&1 | tail -3 | tee gpurun_out/r02_xchg_local.jsonl
echo "== xchg local G=2"; timeout 600 python tools/exp_xchg_local.py 2 64 34 12 2>&1 | tail -3 | tee -a gpurun_out/r02_xchg_local.jsonl
echo "== xchg local G=1"; timeout 600 python tools/exp_xchg_local.py 1 64 34 12 2>&1 | tail -3 | tee -a gpurun_out/r02_xchg_local.jsonl
