#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_ring.py tests/test_gpu_parity.py -m gpu -x -q -k "ring or staged or launch_variant" > gpurun_out/pytest_r2j.txt 2>&1; tail -3 gpurun_out/pytest_r2j.txt; grep -E "^E  " gpurun_out/pytest_r2j.txt | head -20
grep -q "passed" gpurun_out/pytest_r2j.txt && ! grep -q "failed\|error" gpurun_out/pytest_r2j.txt || exit 1
for q in -6 -5; do
echo "== bench qpt=$q"; GPUHASH_SEARCH_QPT=$q timeout 400 python bench.py --verbose --no-cpu --no-ops > gpurun_out/bench_q$q.json 2> gpurun_out/bench_q$q.err; grep -E "resident|ring|zero_copy" gpurun_out/bench_q$q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q$q.json')); print(d['value'], d['roofline']['achieved'], d['roofline']['bulk_launch'], d['e2e']['ring'])"
done
