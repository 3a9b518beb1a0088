#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.txt
echo "== bench N=1 (driver's call shape)"; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 --verbose > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; grep -v "config" gpurun_out/r02_bench_n1.err | tail -6; cut -c1-200 gpurun_out/r02_bench_n1.json
