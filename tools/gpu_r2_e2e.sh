#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --verbose --no-cpu --no-ops --no-ring --no-ref-gpu --no-configs > gpurun_out/r02_bench_e2e.json 2> gpurun_out/r02_bench_e2e.err; tail -4 gpurun_out/r02_bench_e2e.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_e2e.json').read().strip().splitlines()[-1]); print(json.dumps(d['e2e']['variants'],indent=1)); print(d['value'], d['timing']['strict_cycle_order'], d['e2e']['value'])"
