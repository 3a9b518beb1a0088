#!/bin/bash
mkdir -p gpurun_out
echo "== sharded tests"; timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -4
echo "== bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --verbose > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; grep -v "^\*\|OMP_NUM\|^$\|W1017" gpurun_out/r02_bench_n2.err | tail -12; cut -c1-2500 gpurun_out/r02_bench_n2.json
