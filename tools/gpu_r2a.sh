#!/bin/bash
# staged (bulk-copy) search kernel: parity, then A/B against the plain four-lane kernel (resident + e2e)
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2a.txt 2>&1; tail -5 gpurun_out/pytest_r2a.txt
for st in 1 0; do
  echo "== bench staged=$st"
  GPUHASH_SEARCH_STAGED=$st timeout 600 python bench.py --no-cpu --verbose > gpurun_out/bench_r2a_$st.json 2> gpurun_out/bench_r2a_$st.err
  grep -E "resident|e2e" gpurun_out/bench_r2a_$st.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2a_$st.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["path"], "roof", d["roofline"]["achieved"], d["roofline"]["bulk_launch"])
PY
done
