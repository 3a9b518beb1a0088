#!/bin/bash
# round 2, call A: new one-launch cycle kernel -- parity, timing sweep, load-flavour probe
mkdir -p gpurun_out
echo "== pytest cycle_multi"; timeout 900 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu 2>&1 | tail -15
echo "== pytest cycle subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cycle or delete_insert or zero_copy or smoke" 2>&1 | tail -5
echo "== exp_cycles (3 CTAs/SM build)"; timeout 600 python tools/exp_cycles.py 34 20 > gpurun_out/r02_exp_cycles.jsonl 2> gpurun_out/r02_exp_cycles.err; cat gpurun_out/r02_exp_cycles.jsonl; tail -3 gpurun_out/r02_exp_cycles.err
echo "== exp_cycles (4 CTAs/SM build)"; GPUHASH_LIB=build/lib4/libgpuhash.so timeout 600 python tools/exp_cycles.py 34 20 > gpurun_out/r02_exp_cycles_4cta.jsonl 2>&1; grep -E '"W": 64|lone|per_batch' gpurun_out/r02_exp_cycles_4cta.jsonl
echo "== flavours"; timeout 300 ./tools/gather_flavours 34 33554432 > gpurun_out/r02_gather_flavours.jsonl 2>&1; tail -22 gpurun_out/r02_gather_flavours.jsonl
timeout 600 ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_sectors.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum \
    --clock-control none --csv --log-file gpurun_out/r02_flavours_ncu.csv ./tools/gather_flavours 34 4194304 > gpurun_out/r02_flavours_ncu.log 2>&1
ls -la gpurun_out/r02_flavours_ncu.csv
