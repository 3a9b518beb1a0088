#!/usr/bin/env python
"""Mixed 95/5 batch of BASELINE configs[1]: streams x {one launch per batch, one per operation kind} x graph."""
import ctypes as C, json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import megakv_b200 as mk
from megakv_b200 import _native as N
L = mk.lib()
def emit(**kw): print(json.dumps(kw), flush=True)
NS, NI, K = 62259, 3277, 1024
t = mk.DeviceTable(34); geom = t.geom
pop = 1 << 29
buf = mk.DeviceBuffer(12 << 24)
for f in range(0, pop, 1 << 24):
    N.check(L.gpuhash_gen_inserts(buf.ptr, None, 1, f, 1 << 24, None)); N.check(L.gpuhash_insert_flat_ex(C.byref(geom), t.ptr, buf.ptr, 1 << 24, None, 0, None))
N.check(L.gpuhash_device_sync())
sd = mk.DeviceBuffer(8 * NS * K); od = mk.DeviceBuffer(8 * NS * K); idb = mk.DeviceBuffer(12 * NI * K)
N.check(L.gpuhash_gen_queries(sd.ptr, None, 1, pop, NS * K, 99, 0.0, 0.0, None)); N.check(L.gpuhash_device_sync())
nxt = pop
for fused in (1, 0):
    for streams in (4, 8, 16, 32):
        for graph in (1, 0):
            L.gpuhash_set_tuning(C.byref(N.Tune(0, 0, 4, fused)))
            best = 1e30
            for rep in range(3):
                N.check(L.gpuhash_gen_inserts(idb.ptr, None, 1, nxt, NI * K, None)); N.check(L.gpuhash_device_sync()); nxt += NI * K
                r = N.BenchResult()
                N.check(L.gpuhash_bench_resident(C.byref(geom), t.ptr, sd.ptr, NS, od.ptr, idb.ptr, NI, K, streams, graph, C.byref(r)))
                best = min(best, r.total_ms)
            emit(exp="mixed", fused=fused, streams=streams, graph=graph, us_per_step=round(best / K * 1e3, 3), Mops=round(65536 * K / best / 1e3, 1))
for streams in (8, 16, 32):
    r = N.BenchResult(); best = 1e30
    for rep in range(3):
        N.check(L.gpuhash_bench_resident(C.byref(geom), t.ptr, sd.ptr, NS, od.ptr, None, 0, K, streams, 1, C.byref(r))); best = min(best, r.total_ms)
    emit(exp="search_only", streams=streams, us_per_step=round(best / K * 1e3, 3), Mops=round(NS * K / best / 1e3, 1))
