#!/usr/bin/env python
"""One-launch scheduler cycles (gpuhash_bench_cycles) vs the per-batch launch path (gpuhash_bench_resident) on the bench's
16 GiB table: Gops/s by batches per cycle and streams.  Usage: python tools/exp_cycles.py [mem_p] [steps]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megakv_b200 as mk
from megakv_b200 import _native as N

BATCH, N_SEARCH = 65536, 62259
N_INSERT = BATCH - N_SEARCH
mem_p = int(sys.argv[1]) if len(sys.argv) > 1 else 34
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = mk.lib(); mk.require_gpu(); L.gpuhash_set_device(0)
geom = mk.make_geom(mem_p)
table = mk.DeviceBuffer(L.gpuhash_table_bytes(C.byref(geom)), zero=True)
pop = (1 << mem_p) // 32
gen = mk.DeviceBuffer(12 << 24)
for first in range(0, pop, 1 << 24):
    n = min(1 << 24, pop - first)
    N.check(L.gpuhash_gen_inserts(gen.ptr, None, 1, first, n, None))
    N.check(L.gpuhash_insert_flat_ex(C.byref(geom), table.ptr, gen.ptr, n, None, 0, None))
N.check(L.gpuhash_device_sync())
W_MAX = 128
kd = (steps + 5) * W_MAX
sd, od, idd = mk.DeviceBuffer(8 * N_SEARCH * kd), mk.DeviceBuffer(8 * N_SEARCH * kd), mk.DeviceBuffer(12 * N_INSERT * kd)
N.check(L.gpuhash_gen_queries(sd.ptr, None, 1, pop, N_SEARCH * kd, 99, 0.0, 0.0, None))
N.check(L.gpuhash_gen_inserts(idd.ptr, None, 1, pop, N_INSERT * kd, None))
N.check(L.gpuhash_device_sync())


def cycles(W, streams, n_insert, reps=3):
    best = None
    for r in range(reps + 1):
        res = N.BenchResult()
        off = (r % 2) * 5 * W                                       # alternate the batches so a repeat is not an update run
        N.check(L.gpuhash_bench_cycles(C.byref(geom), table.ptr, sd.ptr + 8 * N_SEARCH * off, N_SEARCH, od.ptr + 8 * N_SEARCH * off,
                                       idd.ptr + 12 * N_INSERT * off, n_insert, W, steps, streams, C.byref(res)), "bench_cycles")
        if r and (best is None or res.total_ms < best):
            best = res.total_ms
    return best


only = os.environ.get("EXP_ONLY")
for W in ((64,) if only else (16, 64, 128)):
    for streams in ((1, 2, 3, 4) if only else (1, 2, 3)):
        for n_ins, name in ((N_INSERT, "mixed"), (0, "search")):
            ms = cycles(W, streams, n_ins)
            ops = steps * W * (N_SEARCH + n_ins)
            print(json.dumps({"exp": "cycles", "W": W, "streams": streams, "kind": name, "steps": steps, "ms": round(ms, 3),
                              "Gops": round(ops / ms / 1e6, 2), "us_per_step": round(ms / steps * 1e3, 1)}), flush=True)
if only:
    sys.exit(0)
# the round-1 shape for comparison: one search + one insert launch per batch, 64 streams, one graph
res = N.BenchResult()
for r in range(3):
    N.check(L.gpuhash_bench_resident(C.byref(geom), table.ptr, sd.ptr, N_SEARCH, od.ptr, idd.ptr, N_INSERT, steps * 64, 64, 1, C.byref(res)))
print(json.dumps({"exp": "per_batch_graph_64_streams", "batches": steps * 64, "ms": round(res.total_ms, 3),
                  "Gops": round(steps * 64 * BATCH / res.total_ms / 1e6, 2)}), flush=True)
# a lone cycle of one 64 K batch, call by call (latency shape)
res = N.BenchResult()
for r in range(3):
    N.check(L.gpuhash_bench_cycles(C.byref(geom), table.ptr, sd.ptr, N_SEARCH, od.ptr, idd.ptr, N_INSERT, 1, 200, 1, C.byref(res)))
print(json.dumps({"exp": "lone_batch_cycles_back_to_back", "us_per_batch": round(res.total_ms / 200 * 1e3, 2)}), flush=True)
