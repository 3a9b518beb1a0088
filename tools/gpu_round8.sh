#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu8.txt 2>&1; tail -6 gpurun_out/pytest_gpu8.txt; grep -E "^E  |Error|FAILED" gpurun_out/pytest_gpu8.txt | head -30
echo "== bench N=1"; timeout 900 python bench.py --verbose > gpurun_out/bench8_n1.json 2> gpurun_out/bench8_n1.err; cut -c1-3000 gpurun_out/bench8_n1.json; tail -8 gpurun_out/bench8_n1.err
