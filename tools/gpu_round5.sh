#!/bin/bash
mkdir -p gpurun_out
echo "== flavours"; timeout 120 ./tools/gather_flavours 34 33554432 | grep -E "na.v8|bucket|line|ca.u32 \(4B" | tee gpurun_out/flavours5.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.txt 2>&1; tail -8 gpurun_out/pytest_gpu5.txt; grep -E "^E  |Error" gpurun_out/pytest_gpu5.txt | head -30
echo "== sweep search"; timeout 900 python tools/sweep.py search > gpurun_out/sweep5.jsonl 2> gpurun_out/sweep5.err; grep -E "search_shape|search_64k_shape|miss" gpurun_out/sweep5.jsonl | grep -E '"n": (65536|4194304|16777216)|64k_shape|miss' | cut -c1-200; tail -3 gpurun_out/sweep5.err
