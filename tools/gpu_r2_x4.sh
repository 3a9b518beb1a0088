#!/bin/bash
mkdir -p gpurun_out
for ab in 16 32 64; do
GPUHASH_XCHG_ABLATE=$ab GPUHASH_XCHG_SHAPE=16x8 timeout 600 python tools/exp_xchg_local.py 1 64 34 12 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['env'].get('GPUHASH_XCHG_ABLATE'), d['mixed_us_per_rank_step'], d['search_us_per_rank_step'], d['alone_us'])"
done
