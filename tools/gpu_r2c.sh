#!/bin/bash
mkdir -p gpurun_out
echo "== pytest sharded"; timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "virtual" > gpurun_out/pytest_r2c.txt 2>&1; tail -3 gpurun_out/pytest_r2c.txt; grep -E "^E  " gpurun_out/pytest_r2c.txt | head -10
grep -q "passed" gpurun_out/pytest_r2c.txt && ! grep -q "failed" gpurun_out/pytest_r2c.txt || exit 1
run() { echo "-- $*"; env "$@" timeout 90 python tools/exp_routed_local.py 8 4 16 2>gpurun_out/exp_routed.err | tee -a gpurun_out/exp_routed_local2.jsonl || tail -5 gpurun_out/exp_routed.err; }
run GPUHASH_SERVE_STAGED=1
run GPUHASH_SERVE_STAGED=1 GPUHASH_SERVE_CTAS_PER_SM=4 GPUHASH_SCATTER_CTAS_PER_SM=2 GPUHASH_GATHER_CTAS_PER_SM=1
run GPUHASH_SERVE_STAGED=1 GPUHASH_SERVE_CTAS_PER_SM=6 GPUHASH_SCATTER_CTAS_PER_SM=1 GPUHASH_GATHER_CTAS_PER_SM=1
run GPUHASH_SERVE_STAGED=0 GPUHASH_SERVE_CTAS_PER_SM=4 GPUHASH_SCATTER_CTAS_PER_SM=2 GPUHASH_GATHER_CTAS_PER_SM=2
run GPUHASH_SERVE_STAGED=0 GPUHASH_SERVE_CTAS_PER_SM=8 GPUHASH_SCATTER_CTAS_PER_SM=2 GPUHASH_GATHER_CTAS_PER_SM=2
