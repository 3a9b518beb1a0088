#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:xchg_step -s 521 -c 2 -f -o gpurun_out/r02_xchg_step python tools/exp_xchg_local.py 1 64 34 4 2>&1 | tail -5
