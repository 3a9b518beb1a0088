#!/bin/bash
mkdir -p gpurun_out
N=8
run() { # name, env...
  name=$1; shift
  env "$@" GPUHASH_BENCH_QUICK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/sweep8_$name.err | tail -1 | tee gpurun_out/r02_n8_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['us_per_step'], d['value_Mops'], d['per_gpu_Mops'], d['mismatches'])"
}
run lanes20 GPUHASH_LANES=20
run lanes16b GPUHASH_LANES=16
