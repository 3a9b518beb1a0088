#!/bin/bash
# 8 GPUs: both N > 1 modes of the bench (quick: value + parity only), then the real-NVLink tests at world 4 and 8
mkdir -p gpurun_out
N=${1:-8}
run() { # name, env...
  name=$1; shift
  env "$@" GPUHASH_BENCH_QUICK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02_n${N}_$name.err | tail -1 | tee gpurun_out/r02_n${N}_$name.json | cut -c1-500
  grep -v "^\*\|OMP_NUM\|^$\|W1017\|NCCL version" gpurun_out/r02_n${N}_$name.err | tail -3
}
run lanes GPUHASH_SHARD_MODE=lanes
run lanes16 GPUHASH_SHARD_MODE=lanes GPUHASH_LANES=16
run xchg16x8 GPUHASH_SHARD_MODE=xchg GPUHASH_XCHG_SHAPE=16x8
echo "== tests on $N GPUs"
timeout 900 python -m pytest tests/test_gpu_xchg.py tests/test_gpu_sharded.py -q -m gpu -k "one_process or sharded_index_on_gpus" 2>&1 | tail -4 | tee gpurun_out/r02_n${N}_pytest.txt
