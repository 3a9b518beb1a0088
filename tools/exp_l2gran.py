#!/usr/bin/env python
"""Experiment: does cudaLimitMaxL2FetchGranularity change what a random 32 B probe costs in DRAM traffic?
(ncu on the first run showed 8.25 DRAM sectors per search where the algorithm needs 3.25: the L2 fills whole
128 B lines.)  Also: gather rate vs table size (TLB reach)."""
import ctypes as C, json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import megakv_b200 as mk
from megakv_b200 import _native as N
L = mk.lib()
def emit(**kw): print(json.dumps(kw), flush=True)
mk.require_gpu()
big = mk.DeviceBuffer(1 << 34, zero=True)
emit(exp="default_granularity", bytes=L.gpuhash_get_l2_fetch_granularity())
def gather(nbytes, mode, ilp=4, n=1 << 26):
    ms = C.c_float(); N.check(L.gpuhash_roofline_gather(big.ptr, nbytes, n, mode, ilp, 3, C.byref(ms), None))
    return n / (ms.value / 1e3) / 1e9
for gran in (0, 32, 64, 128):
    if gran:
        rc = L.gpuhash_set_l2_fetch_granularity(gran)
        emit(exp="set_granularity", want=gran, rc=rc, now=L.gpuhash_get_l2_fetch_granularity())
    for mode in (0, 1, 2):
        emit(exp="gather", gran=gran or "default", mode=mode, table="16GiB", Gunits=round(gather(1 << 34, mode), 2))
    if gran in (0, 32):
        for p in (27, 28, 29, 30, 31, 32, 33, 34):
            emit(exp="gather_size", gran=gran or "default", table_log2=p, Gsectors=round(gather(1 << p, 1), 2))
big.free()
# bulk search + 64K pipeline at each granularity
t = mk.DeviceTable(34); geom = t.geom
pop = 1 << 29
buf = mk.DeviceBuffer(12 << 24)
for f in range(0, pop, 1 << 24):
    N.check(L.gpuhash_gen_inserts(buf.ptr, None, 1, f, 1 << 24, None)); N.check(L.gpuhash_insert_flat_ex(C.byref(geom), t.ptr, buf.ptr, 1 << 24, None, 0, None))
N.check(L.gpuhash_device_sync())
sd = mk.DeviceBuffer(8 << 24); od = mk.DeviceBuffer(8 << 24)
N.check(L.gpuhash_gen_queries(sd.ptr, None, 1, pop, 1 << 24, 99, 0.0, 0.0, None)); N.check(L.gpuhash_device_sync())
def resident(n, steps, streams, graph, reps=3):
    best = 1e30
    for _ in range(reps):
        r = N.BenchResult(); N.check(L.gpuhash_bench_resident(C.byref(geom), t.ptr, sd.ptr, n, od.ptr, None, 0, steps, streams, graph, C.byref(r))); best = min(best, r.total_ms)
    return best
for gran in (128, 64, 32):
    L.gpuhash_set_l2_fetch_granularity(gran)
    for qpt in (1, 4):
        for pf in (0, 1):
            L.gpuhash_set_tuning(C.byref(N.Tune(qpt, pf, 4)))
            ms = resident(1 << 24, 1, 1, 0)
            emit(exp="bulk_search", gran=gran, qpt=qpt, prefetch=pf, Mops=round((1 << 24) / ms / 1e3, 1))
    L.gpuhash_set_tuning(C.byref(N.Tune(0, 0, 4)))
    ms = resident(62259, 256, 8, 1)
    emit(exp="pipeline_64k", gran=gran, Mops=round(62259 * 256 / ms / 1e3, 1))
    N.check(L.gpuhash_gen_inserts(buf.ptr, None, 1, pop + gran * (1 << 24), 1 << 24, None)); N.check(L.gpuhash_device_sync())
    r = N.BenchResult(); N.check(L.gpuhash_bench_resident(C.byref(geom), t.ptr, None, 0, None, buf.ptr, 1 << 24, 1, 1, 0, C.byref(r)))
    emit(exp="bulk_insert", gran=gran, Mops=round((1 << 24) / r.total_ms / 1e3, 1))
emit(exp="done")
