#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cycle_multi.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --verbose --no-cpu --no-ops --no-ring --no-ref-gpu > gpurun_out/r02_bench_cfg.json 2> gpurun_out/r02_bench_cfg.err; grep "resident\|config" gpurun_out/r02_bench_cfg.err | cut -c1-700
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg.json').read().strip().splitlines()[-1]); print(d['value'], d['timing']['strict_cycle_order']['Mops/s'])"
