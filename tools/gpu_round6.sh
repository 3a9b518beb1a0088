#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu6.txt 2>&1; tail -6 gpurun_out/pytest_gpu6.txt; grep -E "^E  |Error|FAILED" gpurun_out/pytest_gpu6.txt | head -30
echo "== bench N=1"; timeout 600 python bench.py --verbose --no-cpu > gpurun_out/bench6_n1.json 2> gpurun_out/bench6_n1.err; cut -c1-1500 gpurun_out/bench6_n1.json; tail -4 gpurun_out/bench6_n1.err
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 20 --verbose > gpurun_out/bench6_n2.json 2> gpurun_out/bench6_n2.err; cut -c1-2500 gpurun_out/bench6_n2.json; grep -vE "^\s*$|Warning|warn" gpurun_out/bench6_n2.err | tail -25
