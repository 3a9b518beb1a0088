#!/bin/bash
mkdir -p gpurun_out
echo "== flavours 16GiB"; timeout 120 ./tools/gather_flavours 34 33554432 | tee gpurun_out/flavours_16g.jsonl
echo "== flavours under ncu (sectors per access, n=2^22)"
timeout 300 ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_sectors.sum,gpu__time_duration.sum \
  --clock-control none --csv --log-file gpurun_out/ncu_flavours.csv ./tools/gather_flavours 34 4194304 > gpurun_out/ncu_flavours.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/ncu_flavours.csv')) if len(r)>10]
hdr=rows[0]; mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID'); ki=hdr.index('Kernel Name')
cur=collections.OrderedDict()
for r in rows[1:]:
    cur.setdefault(r[ii],{'k':r[ki][:40]})[r[mi]]=r[vi]
for k,v in cur.items():
    if int(k)%3==1:
        n=4194304
        print(k, v['k'], 'dram/acc=%.2f'%(float(v['dram__sectors_read.sum'].replace(',',''))/n), 'lts/acc=%.2f'%(float(v['lts__t_sectors_srcunit_tex_op_read.sum'].replace(',',''))/n), 'xbar2l1/acc=%.2f'%(float(v.get('l1tex__m_xbar2l1tex_read_sectors.sum','0').replace(',',''))/n), 'us', float(v['gpu__time_duration.sum'].replace(',',''))/1e3)
PY
echo "== failing tests, full output"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multiset_exact or high_load or contention" 2>&1 | grep -vE "^\s*$" | tail -120 > gpurun_out/pytest_fail.txt; tail -60 gpurun_out/pytest_fail.txt
