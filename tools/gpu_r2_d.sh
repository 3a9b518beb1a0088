#!/bin/bash
mkdir -p gpurun_out
echo "== pytest cycle_multi"; timeout 900 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu 2>&1 | tail -8
echo "== bench N=1 (driver shape)"; timeout 1500 python bench.py --steps 20 --warmup 5 --no-configs --verbose > gpurun_out/r02_bench_n1_s20.json 2> gpurun_out/r02_bench_n1_s20.err; tail -25 gpurun_out/r02_bench_n1_s20.err; cut -c1-1500 gpurun_out/r02_bench_n1_s20.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 --verbose > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -3 gpurun_out/r02_bench_ref.err; cut -c1-600 gpurun_out/r02_bench_ref.json
