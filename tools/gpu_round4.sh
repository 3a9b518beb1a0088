#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -vE "^\s*$" | tail -60 | tee gpurun_out/pytest_gpu4.txt
echo "== sweep search"; timeout 900 python tools/sweep.py search > gpurun_out/sweep4.jsonl 2> gpurun_out/sweep4.err; grep -E "search_shape|pipeline|miss|insert_fresh|delete" gpurun_out/sweep4.jsonl | grep -E '"n": (65536|16777216|4194304)|pipeline|miss|insert|delete' | cut -c1-220; tail -3 gpurun_out/sweep4.err
echo "== bench"; timeout 900 python bench.py --verbose > gpurun_out/bench4.json 2> gpurun_out/bench4.err; cat gpurun_out/bench4.json; tail -5 gpurun_out/bench4.err
