#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --verbose --no-cpu --no-ops --no-ring --no-ref-gpu --config 2 > gpurun_out/r02_bench_cfg.json 2> gpurun_out/r02_bench_cfg.err; grep config gpurun_out/r02_bench_cfg.err | cut -c1-1800
