#!/bin/bash
mkdir -p gpurun_out
echo "== xchg tests"; timeout 900 python -m pytest tests/test_gpu_xchg.py -x -q -m gpu 2>&1 | tail -3
for g in 1 8; do
echo "== xchg local G=$g"; timeout 600 python tools/exp_xchg_local.py $g 64 34 12 2>&1 | tail -1 | tee -a gpurun_out/r02_xchg_local3.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['G'], d['mixed_us_per_rank_step'], d['search_us_per_rank_step'], d.get('alone_us'), d['mismatches'])"
done
