#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + full captures of the search and routed kernels
mkdir -p gpurun_out
echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/pytest_gpu_final.txt; grep -E "^E  |^FAILED" gpurun_out/pytest_gpu_final.txt | head
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --verbose > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; cut -c1-300 gpurun_out/bench_r01.json; tail -12 gpurun_out/bench_r01.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-ops --no-ring > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_warp_kernel -s 20 -c 3 -f -o gpurun_out/r01_search_warp \
    python bench.py --steps 40 --warmup 3 --graph 0 --no-cpu --no-ops --no-ring > gpurun_out/ncu_full.log 2>&1
GPUHASH_FORCE_SHARDED=1 GPUHASH_BENCH_QUICK=1 GPUHASH_GROUP=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"serve_search_staged|route_scatter_tiles_kernel<2>|route_gather_tiles" -s 6 -c 6 -f -o gpurun_out/r01_routed \
    python bench.py --steps 64 --warmup 16 --graph 0 > gpurun_out/ncu_routed.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches.csv
