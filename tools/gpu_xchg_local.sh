#!/bin/bash
mkdir -p gpurun_out
echo "== xchg tests"; timeout 900 python -m pytest tests/test_gpu_xchg.py -x -q -m gpu 2>&1 | tail -8
echo "== xchg tests, 8 router warps"; GPUHASH_XCHG_ROUTER_WARPS=8 timeout 900 python -m pytest tests/test_gpu_xchg.py -x -q -m gpu 2>&1 | tail -3
: > gpurun_out/r02_xchg_local.jsonl
for rw in 4 8; do for g in 1 8; do
echo "== xchg local G=$g router warps $rw"; GPUHASH_XCHG_ROUTER_WARPS=$rw timeout 600 python tools/exp_xchg_local.py $g 64 34 12 2>&1 | tail -2 | tee -a gpurun_out/r02_xchg_local.jsonl
done; done
