#!/bin/bash
mkdir -p gpurun_out
echo "== pytest new tests"; timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q -k "virtual or staged or zero_copy or launch_variant" > gpurun_out/pytest_r2b.txt 2>&1; tail -5 gpurun_out/pytest_r2b.txt; grep -E "^E  " gpurun_out/pytest_r2b.txt | head -20
echo "== routed local"
for cfg in "8 4 16" "2 4 16" "8 2 16" "8 6 4"; do
  timeout 300 python tools/exp_routed_local.py $cfg 2>gpurun_out/exp_routed.err | tee -a gpurun_out/exp_routed_local.jsonl || tail -5 gpurun_out/exp_routed.err
done
