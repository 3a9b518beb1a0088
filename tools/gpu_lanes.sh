#!/bin/bash
for cfg in "16 2" "16 4" "16 6" "16 8" "4 8" "1 32"; do
  set -- $cfg
  GPUHASH_BENCH_QUICK=1 GPUHASH_GROUP=$1 GPUHASH_LANES=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 48 2>gpurun_out/lanes.err | grep quick || tail -5 gpurun_out/lanes.err
done
echo "== full bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 50 --verbose > gpurun_out/bench15_n2.json 2> gpurun_out/bench15_n2.err; cut -c1-2800 gpurun_out/bench15_n2.json; grep -E "rank0\]|bench r0" gpurun_out/bench15_n2.err | tail -12
