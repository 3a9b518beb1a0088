#!/bin/bash
mkdir -p gpurun_out
echo "== pytest sharded"; timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/pytest_gpu12.txt 2>&1; tail -4 gpurun_out/pytest_gpu12.txt; grep -E "^E  |Error" gpurun_out/pytest_gpu12.txt | head -10
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 50 --verbose > gpurun_out/bench12_n2.json 2> gpurun_out/bench12_n2.err; cut -c1-2500 gpurun_out/bench12_n2.json; grep -E "rank0\]|bench r0" gpurun_out/bench12_n2.err | tail -12
