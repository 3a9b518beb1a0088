#!/bin/bash
mkdir -p gpurun_out
run() { echo "-- $*"; env "$@" GPUHASH_BENCH_QUICK=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1536 --warmup 64 2>gpurun_out/r2g.err | grep quick || tail -5 gpurun_out/r2g.err; }
run GPUHASH_LANES=8
run GPUHASH_LANES=12
run GPUHASH_LANES=8 GPUHASH_GROUP=32
run GPUHASH_LANES=8 GPUHASH_SERVE_CTAS_PER_SM=6
run GPUHASH_LANES=8 GPUHASH_SERVE_CTAS_PER_SM=3
run GPUHASH_LANES=8 GPUHASH_SERVE_STAGED=0 GPUHASH_SERVE_CTAS_PER_SM=4
run GPUHASH_LANES=8 GPUHASH_SCATTER_CTAS_PER_SM=4 GPUHASH_GATHER_CTAS_PER_SM=4
run GPUHASH_LANES=16 GPUHASH_GROUP=8
