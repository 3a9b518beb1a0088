#!/bin/bash
mkdir -p gpurun_out
run() { echo "-- $*"; env "$@" EXP_SKIP_PARTS=1 timeout 100 python tools/exp_routed_local.py 2>gpurun_out/exp_routed.err | tee -a gpurun_out/exp_routed_local4.jsonl || tail -5 gpurun_out/exp_routed.err; }
run EXP_ARGS="8 4 16" EXP_STAGGER=1
run EXP_ARGS="8 8 16" EXP_CYCLES=32
run EXP_ARGS="8 8 16" EXP_CYCLES=32 EXP_STAGGER=1
run EXP_ARGS="8 8 8" EXP_CYCLES=32 EXP_STAGGER=1
run EXP_ARGS="8 12 4" EXP_CYCLES=48 EXP_STAGGER=1
