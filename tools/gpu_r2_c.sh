#!/bin/bash
mkdir -p gpurun_out
echo "== atomic probe"; timeout 120 ./tools/atomic_probe > gpurun_out/r02_atomic_probe.jsonl 2>&1; grep -E '"addresses": (1|64),' gpurun_out/r02_atomic_probe.jsonl
echo "== pytest cycle_multi"; timeout 900 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu 2>&1 | tail -15
echo "== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -x -q -m gpu 2>&1 | tail -5
echo "== exp_cycles (3 CTAs/SM build)"; timeout 600 python tools/exp_cycles.py 34 20 > gpurun_out/r02_exp_cycles.jsonl 2> gpurun_out/r02_exp_cycles.err; cat gpurun_out/r02_exp_cycles.jsonl; tail -3 gpurun_out/r02_exp_cycles.err
for v in lib4 lib2 lib2p; do echo "== exp_cycles ($v)"; EXP_ONLY=1 GPUHASH_LIB=build/$v/libgpuhash.so timeout 600 python tools/exp_cycles.py 34 20 2>&1 | tail -4; done
