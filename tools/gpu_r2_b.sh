#!/bin/bash
mkdir -p gpurun_out
echo "== pytest cycle_multi"; timeout 900 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu 2>&1 | tail -15
echo "== pytest cycle subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cycle or delete_insert or zero_copy" 2>&1 | tail -5
echo "== exp_cycles (3 CTAs/SM build)"; timeout 600 python tools/exp_cycles.py 34 20 > gpurun_out/r02_exp_cycles.jsonl 2> gpurun_out/r02_exp_cycles.err; cat gpurun_out/r02_exp_cycles.jsonl; tail -3 gpurun_out/r02_exp_cycles.err
echo "== exp_cycles (4 CTAs/SM build)"; EXP_ONLY=1 GPUHASH_LIB=build/lib4/libgpuhash.so timeout 600 python tools/exp_cycles.py 34 20 2>&1 | tail -5
echo "== exp_cycles (2 CTAs/SM build)"; EXP_ONLY=1 GPUHASH_LIB=build/lib2/libgpuhash.so timeout 600 python tools/exp_cycles.py 34 20 2>&1 | tail -5
EXP_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:cycle_multi -s 10 -c 2 -f -o gpurun_out/r02_cycle_multi python tools/exp_cycles.py 34 4 > gpurun_out/ncu_cycle.log 2>&1
ls -la gpurun_out/*.ncu-rep
