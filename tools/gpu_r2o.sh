#!/bin/bash
mkdir -p gpurun_out
for f in 1 0; do
echo "== bench graph, fused=$f"; BENCH_FUSED=$f timeout 300 python bench.py --verbose --no-cpu --no-ops --no-ring > gpurun_out/bf$f.json 2> gpurun_out/bf$f.err; grep -E "resident|zero" gpurun_out/bf$f.err
done
