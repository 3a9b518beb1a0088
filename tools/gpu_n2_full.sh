#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --verbose > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; grep -v "^\*\|OMP_NUM\|^$\|W1017" gpurun_out/r02_bench_n$N.err | tail -8; cut -c1-3000 gpurun_out/r02_bench_n$N.json
