#!/bin/bash
mkdir -p gpurun_out
echo "== pytest parity (pair insert/delete kernels are the default now)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -m gpu -x -q > gpurun_out/pytest_r2l.txt 2>&1; tail -3 gpurun_out/pytest_r2l.txt; grep -E "^E  " gpurun_out/pytest_r2l.txt | head -20
grep -q "passed" gpurun_out/pytest_r2l.txt && ! grep -q "failed\|error" gpurun_out/pytest_r2l.txt || exit 1
run() { echo "-- $*"; env "$@" timeout 300 python bench.py --verbose --no-cpu --no-ring > gpurun_out/b.json 2> gpurun_out/b.err; grep -E "resident" gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['roofline']['achieved'], {k:v['Mops/s'] for k,v in d['ops'].items() if isinstance(v,dict)}, d['e2e']['value'])"; }
run GPUHASH_UPDATE_PAIR=1
run GPUHASH_UPDATE_PAIR=0
run GPUHASH_WARP_BLOCK=128
run GPUHASH_WARP_BLOCK=256
