#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m tests.golden.make_ref_golden sparse 2>&1 | tail -12
ls -la gpurun_out/ref_golden/
