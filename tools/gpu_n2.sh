#!/bin/bash
mkdir -p gpurun_out
echo "== xchg tests (2 GPUs)"; timeout 900 python -m pytest tests/test_gpu_xchg.py -x -q -m gpu -k one_process 2>&1 | tail -4
run() { # name, env...
  name=$1; shift
  env "$@" GPUHASH_BENCH_QUICK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r02_n2_$name.err | tail -1 | tee gpurun_out/r02_n2_$name.json | cut -c1-600
  grep -v "^\*\|OMP_NUM\|^$\|W1017\|NCCL version" gpurun_out/r02_n2_$name.err | tail -5
}
run xchg4 GPUHASH_SHARD_MODE=xchg GPUHASH_XCHG_ROUTER_WARPS=4
run xchg8 GPUHASH_SHARD_MODE=xchg GPUHASH_XCHG_ROUTER_WARPS=8
run lanes GPUHASH_SHARD_MODE=lanes
