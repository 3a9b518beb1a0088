#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1024 --warmup 64 --verbose > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "rc=$?"; cut -c1-2500 gpurun_out/bench_r2_n$N.json; grep -E "bench r0|Error|error" gpurun_out/bench_r2_n$N.err | tail -12
