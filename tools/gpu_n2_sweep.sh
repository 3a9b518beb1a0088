#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
run() { # name, env...
  name=$1; shift
  env "$@" GPUHASH_BENCH_QUICK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/sweep_$name.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['us_per_step'], d['per_gpu_Mops'], d['mismatches'])"
}
run lanes16 GPUHASH_LANES=16
run lanes10 GPUHASH_LANES=10
run lanes20 GPUHASH_LANES=20
run lanes5 GPUHASH_LANES=5
