#!/bin/bash
# NVLink data counters around one quick 2-GPU bench run (ncu cannot profile a multi-rank command)
mkdir -p gpurun_out
N=2
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvl_before.txt 2>&1
GPUHASH_BENCH_QUICK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/nvl_bench.err | tail -1 > gpurun_out/nvl_bench.json
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvl_after.txt 2>&1
head -8 gpurun_out/nvl_before.txt; echo ...; head -8 gpurun_out/nvl_after.txt; cut -c1-300 gpurun_out/nvl_bench.json
python - <<'PY'
import re
def tot(p):
    tx=rx=0
    for l in open(p):
        m=re.search(r'Data Tx:\s*(\d+)\s*KiB',l);  tx+= int(m.group(1)) if m else 0
        m=re.search(r'Data Rx:\s*(\d+)\s*KiB',l);  rx+= int(m.group(1)) if m else 0
    return tx,rx
b=tot('gpurun_out/nvl_before.txt'); a=tot('gpurun_out/nvl_after.txt')
print("GPU0 NVLink data during the run: Tx %.1f MiB, Rx %.1f MiB" % ((a[0]-b[0])/1024, (a[1]-b[1])/1024))
PY
