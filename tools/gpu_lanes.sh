#!/bin/bash
for lanes in 4 8 16 32 64; do
  for graph in 1 0; do
    GPUHASH_BENCH_QUICK=1 GPUHASH_LANES=$lanes timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 600 --warmup 30 --graph $graph 2>/dev/null | grep quick
  done
done
