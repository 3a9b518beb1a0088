#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu.txt
echo "== bench N=1 (steps 20)"; timeout 1200 python bench.py --steps 20 --warmup 5 --verbose > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -12 gpurun_out/r02_bench_n1.err; cut -c1-600 gpurun_out/r02_bench_n1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ('value','config2','config3')}, indent=1)[:6000])
PY
