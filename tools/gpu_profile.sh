#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + full captures of the search and insert kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 40 --warmup 3 --graph 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_quad_kernel -s 20 -c 3 -f -o gpurun_out/r01_search_quad \
    python bench.py --steps 40 --warmup 3 --graph 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:insert_flat_kernel -s 40 -c 2 -f -o gpurun_out/r01_insert_flat \
    python bench.py --steps 40 --warmup 3 --graph 0 --no-cpu > gpurun_out/ncu_full_ins.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches.csv
echo "== final bench"; timeout 900 python bench.py --verbose > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; cut -c1-400 gpurun_out/bench_r01.json; tail -3 gpurun_out/bench_r01.err
echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/pytest_gpu_final.txt
