#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2m.txt 2>&1; tail -3 gpurun_out/pytest_r2m.txt; grep -E "^E  |^FAILED" gpurun_out/pytest_r2m.txt | head -20
grep -q "passed" gpurun_out/pytest_r2m.txt && ! grep -q "failed\|error" gpurun_out/pytest_r2m.txt || exit 1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench (graph 0: cycle kernel) "; timeout 300 python bench.py --verbose --no-cpu --no-ops --graph 0 > gpurun_out/b0.json 2> gpurun_out/b0.err; grep -E "resident|ring|zero" gpurun_out/b0.err
echo "== routed virtual"; EXP_ARGS="8 8 16" EXP_CYCLES=32 EXP_SKIP_PARTS=1 timeout 100 python tools/exp_routed_local.py 2>gpurun_out/exp_routed.err || tail -5 gpurun_out/exp_routed.err
