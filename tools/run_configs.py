#!/usr/bin/env python
"""BASELINE configs[2] / configs[3] legs of bench.py on their own (megakv_b200/bench_configs.py), e.g. under ncu:
   ncu --metrics lts__t_sector_hit_rate.pct,dram__sectors_read.sum,dram__sectors_write.sum,gpu__time_duration.sum \
       --clock-control none --csv --log-file gpurun_out/x.csv python tools/run_configs.py [steps=4] [config=0|2|3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megakv_b200 as mk
from megakv_b200 import _native as N
from megakv_b200 import bench_configs

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = mk.lib(); mk.require_gpu(); N.check(L.gpuhash_set_device(0))
args = argparse.Namespace(steps=steps, batches_per_step=64, config=cfg)
log = lambda m: print(m, file=sys.stderr, flush=True)
print(json.dumps(bench_configs.run(args, L, N, mk, 0, log, 47.4e9)))
