#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
run() { # name, env...
  name=$1; shift
  env "$@" GPUHASH_BENCH_QUICK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/sweep_$name.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['us_per_step'], d['per_gpu_Mops'], d['mismatches'])"
}
run base
run lanes12 GPUHASH_LANES=12
run lanes16 GPUHASH_LANES=16
run lanes4 GPUHASH_LANES=4
run sc4 GPUHASH_SCATTER_CTAS_PER_SM=4
run sc8 GPUHASH_SCATTER_CTAS_PER_SM=8
run ga4 GPUHASH_GATHER_CTAS_PER_SM=4
run ga8 GPUHASH_GATHER_CTAS_PER_SM=8
run sc4ga4 GPUHASH_SCATTER_CTAS_PER_SM=4 GPUHASH_GATHER_CTAS_PER_SM=4
run serve8 GPUHASH_SERVE_CTAS_PER_SM=8
run serve2 GPUHASH_SERVE_CTAS_PER_SM=2
run steps60 
