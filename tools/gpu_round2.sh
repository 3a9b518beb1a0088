#!/bin/bash
# GPU call 2: L2 fetch granularity experiment, reference differential (serial first, short timeouts), parity suite re-run.
mkdir -p gpurun_out
echo "== exp l2 granularity"; timeout 600 python tools/exp_l2gran.py > gpurun_out/exp_l2gran.jsonl 2> gpurun_out/exp_l2gran.err; cat gpurun_out/exp_l2gran.jsonl; tail -3 gpurun_out/exp_l2gran.err
echo "== ncu dram sectors per gather at granularity 32/128"
cat > /tmp/g.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
import megakv_b200 as mk
from megakv_b200 import _native as N
L = mk.lib(); big = mk.DeviceBuffer(1 << 34, zero=True)
for gran in (128, 64, 32):
    L.gpuhash_set_l2_fetch_granularity(gran)
    ms = C.c_float(); N.check(L.gpuhash_roofline_gather(big.ptr, 1 << 34, 1 << 22, 1, 4, 1, C.byref(ms), None))
PY
timeout 300 ncu --metrics dram__sectors_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:gather_kernel --csv --log-file gpurun_out/ncu_gather_gran.csv python /tmp/g.py > /dev/null 2>&1
cat gpurun_out/ncu_gather_gran.csv | grep -E 'dram__sectors|time_duration|lts__t' | cut -d, -f5,13-15 | head -20
echo "== ref serial"; timeout 90 python -m tests.golden.make_ref_golden serial > gpurun_out/ref_serial.log 2>&1; echo rc=$?; tail -4 gpurun_out/ref_serial.log
echo "== ref batch";  timeout 45 python -m tests.golden.make_ref_golden batch > gpurun_out/ref_batch.log 2>&1; echo rc=$?; tail -4 gpurun_out/ref_batch.log
nvidia-smi --query-gpu=name,memory.used --format=csv
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu2.txt
