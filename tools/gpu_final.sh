#!/bin/bash
# last check of the tree as committed: CPU-visible build info, full GPU suite, smoke, the two bench arms
mkdir -p gpurun_out
echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/pytest_gpu_final.txt; grep -E "^E  |^FAILED" gpurun_out/pytest_gpu_final.txt | head
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --verbose > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_r01.json)"; cut -c1-200 gpurun_out/bench_r01.json; grep -E "resident|e2e|ring" gpurun_out/bench_r01.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_ref.json)"; cut -c1-200 gpurun_out/bench_ref.json
