#!/bin/bash
mkdir -p gpurun_out
echo "== pytest sharded"; timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/pytest_gpu17.txt 2>&1; tail -4 gpurun_out/pytest_gpu17.txt; grep -E "^E  |Error" gpurun_out/pytest_gpu17.txt | head -10
echo "== routed path on ONE GPU"
GPUHASH_FORCE_SHARDED=1 GPUHASH_BENCH_QUICK=1 timeout 300 python bench.py --steps 640 --warmup 32 2>gpurun_out/r17.err | grep quick || tail -5 gpurun_out/r17.err
echo "== N=2 quick"
for cfg in "16 6" "16 8"; do set -- $cfg
GPUHASH_BENCH_QUICK=1 GPUHASH_GROUP=$1 GPUHASH_LANES=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 48 2>gpurun_out/lanes.err | grep quick || tail -5 gpurun_out/lanes.err
done
