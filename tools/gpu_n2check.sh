#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1024 --warmup 64 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
echo "rc=$? stdout lines: $(wc -l < gpurun_out/bench_r2_n2.json)"; cut -c1-200 gpurun_out/bench_r2_n2.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 50 --warmup 2 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref rc=$? stdout lines: $(wc -l < gpurun_out/bench_ref_n2.json)"; cut -c1-200 gpurun_out/bench_ref_n2.json
