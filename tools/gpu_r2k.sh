#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 300 python -m pytest tests/test_gpu_ring.py -m gpu -x -q > gpurun_out/pytest_r2k.txt 2>&1; tail -3 gpurun_out/pytest_r2k.txt; grep -E "^E  " gpurun_out/pytest_r2k.txt | head -20
grep -q "passed" gpurun_out/pytest_r2k.txt && ! grep -q "failed\|error" gpurun_out/pytest_r2k.txt || exit 1
echo "== bench"; timeout 400 python bench.py --verbose --no-cpu --no-ops > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; grep -E "resident|ring|zero_copy" gpurun_out/bench_r2k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2k.json')); print(d['value'], d['e2e']['ring'])"
