#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 4" "16 6" "16 8"; do
  set -- $cfg
  GPUHASH_BENCH_QUICK=1 GPUHASH_GROUP=$1 GPUHASH_LANES=$2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1024 --warmup 64 2>gpurun_out/r2e.err | grep quick || tail -5 gpurun_out/r2e.err
done
GPUHASH_ROUTE_TILES=0 GPUHASH_SERVE_STAGED=0 GPUHASH_SERVE_CTAS_PER_SM=16 GPUHASH_BENCH_QUICK=1 GPUHASH_GROUP=16 GPUHASH_LANES=6 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1024 --warmup 64 2>gpurun_out/r2e.err | grep quick || tail -5 gpurun_out/r2e.err
