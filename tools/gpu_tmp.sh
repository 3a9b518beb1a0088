#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q -k "virtual or compact_results_mode or sharded_index" 2>&1 | tail -3
EXP_ARGS="8 8 16" EXP_CYCLES=32 EXP_SKIP_PARTS=1 timeout 100 python tools/exp_routed_local.py 2>gpurun_out/exp_routed.err || tail -5 gpurun_out/exp_routed.err
