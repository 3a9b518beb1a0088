#!/bin/bash
mkdir -p gpurun_out
echo "== xchg tests"; timeout 900 python -m pytest tests/test_gpu_xchg.py -x -q -m gpu 2>&1 | tail -15
