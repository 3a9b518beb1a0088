#!/bin/bash
mkdir -p gpurun_out
echo "== pytest ring"; timeout 240 python -m pytest tests/test_gpu_ring.py -m gpu -x -q > gpurun_out/pytest_ring.txt 2>&1; tail -5 gpurun_out/pytest_ring.txt; grep -E "^E  " gpurun_out/pytest_ring.txt | head -20
grep -q "passed" gpurun_out/pytest_ring.txt && ! grep -q "failed\|error" gpurun_out/pytest_ring.txt || exit 1
echo "== bench"; timeout 400 python bench.py --verbose --no-cpu --no-ops > gpurun_out/bench_ring.json 2> gpurun_out/bench_ring.err; tail -8 gpurun_out/bench_ring.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ring.json')); print(d['e2e'])"
