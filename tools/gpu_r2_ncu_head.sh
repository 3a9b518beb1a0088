#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:cycle_multi -s 12 -c 1 -f -o gpurun_out/r02_cycle_search python bench.py --steps 4 --warmup 3 --reps 1 --no-cpu --no-ops --no-ring --no-ref-gpu --no-configs 2>&1 | tail -3 | cut -c1-300
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-ops --no-ring --no-ref-gpu --no-configs > gpurun_out/r02_bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches.csv
