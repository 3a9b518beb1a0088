#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
echo "== sharded tests, warp serve"; GPUHASH_SERVE_MODE=warp timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" GPUHASH_BENCH_QUICK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02_n${N}_$name.err | tail -1 | tee gpurun_out/r02_n${N}_$name.json | cut -c1-420
  grep -v "^\*\|OMP_NUM\|^$\|W1017\|NCCL version" gpurun_out/r02_n${N}_$name.err | tail -3
}
run lanes_staged GPUHASH_SHARD_MODE=lanes GPUHASH_SERVE_MODE=staged
run lanes_warp GPUHASH_SHARD_MODE=lanes GPUHASH_SERVE_MODE=warp
run lanes_warp4 GPUHASH_SHARD_MODE=lanes GPUHASH_SERVE_MODE=warp GPUHASH_SERVE_CTAS_PER_SM=4
run lanes_warp2 GPUHASH_SHARD_MODE=lanes GPUHASH_SERVE_MODE=warp GPUHASH_SERVE_CTAS_PER_SM=2
run xchg16x8 GPUHASH_SHARD_MODE=xchg GPUHASH_XCHG_SHAPE=16x8
