#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ring.py -m gpu -x -q -k "cycle or zero_copy or legacy or ring" > gpurun_out/pytest_r2n.txt 2>&1; tail -3 gpurun_out/pytest_r2n.txt; grep -E "^E  |^FAILED" gpurun_out/pytest_r2n.txt | head -20
grep -q "passed" gpurun_out/pytest_r2n.txt && ! grep -q "failed\|error" gpurun_out/pytest_r2n.txt || exit 1
echo "== bench (graph 0: cycle kernel) "; timeout 300 python bench.py --verbose --no-cpu --no-ops --no-ring --graph 0 > gpurun_out/b0.json 2> gpurun_out/b0.err; grep -E "resident|ring|zero" gpurun_out/b0.err
