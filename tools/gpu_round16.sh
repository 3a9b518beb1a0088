#!/bin/bash
mkdir -p gpurun_out
echo "== routed path on ONE GPU (all exchanges local): quick value + per-kernel times"
GPUHASH_FORCE_SHARDED=1 GPUHASH_BENCH_QUICK=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex.sum --clock-control none -k regex:"serve_search|gather|scatter_pub_kernel<2" -c 30 --csv --log-file gpurun_out/ncu_routed.csv python bench.py --steps 64 --warmup 16 --graph 0 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/ncu_routed.csv')) if len(r)>10]
hdr=rows[0]; mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID'); ki=hdr.index('Kernel Name'); gi=hdr.index('Grid Size')
cur=collections.OrderedDict()
for r in rows[1:]:
    cur.setdefault(r[ii],{'k':r[ki][:48],'grid':r[gi]})[r[mi]]=float(r[vi].replace(',',''))
for k,v in list(cur.items())[-20:]:
    print(k, v['k'], v['grid'], 'us=%.1f'%(v['gpu__time_duration.sum']/1e3), 'dramR_MB=%.1f'%(v['dram__bytes_read.sum']/1e6), 'dramW_MB=%.1f'%(v['dram__bytes_write.sum']/1e6), 'l2req_M=%.2f'%(v.get('lts__t_requests_srcunit_tex.sum',0)/1e6))
PY
