#!/usr/bin/env python
"""Design-space sweep on one B200 (run under gpurun; writes JSON lines to stdout).

Answers the questions DESIGN.md needs numbers for:
  1. what is the random-32 B-sector ceiling of this GPU (the real roofline of a hash probe), and does DRAM move
     32 B or 64 B per random access (=> AoS bucket vs split signature/location arrays)?
  2. one thread per request with 256-bit row loads vs a 4-lane cooperative group with 128-bit loads; how many
     requests per thread; does the L2::64B prefetch hint on the signature row pay?
  3. how much of the kernel's bulk rate survives at the 64K-request batch of BASELINE configs[1], as a function
     of streams and CUDA-graph replay?
  4. insert / delete rates at low and high load; L2-resident (MEM_P 26) vs HBM (MEM_P 34).
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import megakv_b200 as mk  # noqa: E402
from megakv_b200 import _native as N  # noqa: E402

L = mk.lib()
SEED = 1


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timed_resident(geom, table, search_d, n_search, out_d, insert_d, n_insert, steps, streams, graph, reps=3):
    best = 1e30
    for _ in range(reps):
        res = N.BenchResult()
        N.check(L.gpuhash_bench_resident(C.byref(geom), table, search_d, n_search, out_d, insert_d, n_insert,
                                         steps, streams, graph, C.byref(res)))
        best = min(best, res.total_ms)
    return best


def preload(geom, table, pop, first=0):
    chunk = 1 << 24
    buf = mk.DeviceBuffer(12 * min(chunk, pop))
    t0 = time.time()
    for f in range(first, first + pop, chunk):
        n = min(chunk, first + pop - f)
        N.check(L.gpuhash_gen_inserts(buf.ptr, None, SEED, f, n, None))
        N.check(L.gpuhash_insert_flat_ex(C.byref(geom), table, buf.ptr, n, None, 0, None))
    N.check(L.gpuhash_device_sync())
    buf.free()
    return time.time() - t0


def gather_sweep(table, nbytes, tag):
    for mode, name in [(0, "sig sector of random 64B bucket"), (1, "random 32B sector"), (2, "whole random 64B bucket")]:
        for ilp in (1, 2, 4, 8):
            ms = C.c_float()
            n = 1 << 27
            N.check(L.gpuhash_roofline_gather(table, nbytes, n, mode, ilp, 3, C.byref(ms), None))
            units = n / (ms.value / 1e3)
            emit(exp="gather", table=tag, mode=name, ilp=ilp, ms=round(ms.value, 4),
                 Gunits_per_s=round(units / 1e9, 3), GBps_useful=round(units * (64 if mode == 2 else 32) / 1e9, 1))


def search_sweep(mem_p, tag, layout=N.LAYOUT_PAIRS):
    t = mk.DeviceTable(mem_p, layout=layout)
    tag = tag + ("/pairs" if layout == N.LAYOUT_PAIRS else "/reflayout")
    geom = t.geom
    pop = (1 << mem_p) // 8 // 4
    dt = preload(geom, t.ptr, pop)
    emit(exp="preload", table=tag, keys=pop, seconds=round(dt, 3), Mops=round(pop / dt / 1e6, 1))
    nmax = 1 << 24
    sd = mk.DeviceBuffer(8 * nmax); od = mk.DeviceBuffer(8 * nmax)
    N.check(L.gpuhash_gen_queries(sd.ptr, None, SEED, pop, nmax, 99, 0.0, 0.0, None))
    N.check(L.gpuhash_device_sync())
    old = N.Tune(); L.gpuhash_get_tuning(C.byref(old))
    # ---- 2: kernel shape, one launch at a time
    for n in (1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24):
        for qpt in (1, 2, 4, -4):
            for pf in (0,):
                L.gpuhash_set_tuning(C.byref(N.Tune(qpt, pf, 4)))
                reps = max(1, min(200, nmax // n))                 # reps * n requests must stay inside the buffers
                ms = timed_resident(geom, t.ptr, sd.ptr, n, od.ptr, None, 0, reps, 1, 0) / reps
                emit(exp="search_shape", table=tag, n=n, qpt=qpt, prefetch=pf, us_per_launch=round(ms * 1e3, 3),
                     Mops=round(n / ms / 1e3, 1), GBps_112=round(n * 112 / ms / 1e6, 1))
    # ---- 3: the 64K batch, pipelined
    L.gpuhash_set_tuning(C.byref(old))
    n = 62259
    steps = 256
    for streams in (1, 2, 4, 8, 16, 32):
        for graph in (0, 1):
            ms = timed_resident(geom, t.ptr, sd.ptr, n, od.ptr, None, 0, steps, streams, graph)
            emit(exp="search_64k_pipeline", table=tag, streams=streams, graph=graph, us_per_batch=round(ms / steps * 1e3, 3),
                 Mops=round(n * steps / ms / 1e3, 1), GBps_112=round(n * steps * 112 / ms / 1e6, 1))
    for qpt in (1, 2, -4):
        for pf in (0,):
            L.gpuhash_set_tuning(C.byref(N.Tune(qpt, pf, 4)))
            ms = timed_resident(geom, t.ptr, sd.ptr, n, od.ptr, None, 0, steps, 8, 1)
            emit(exp="search_64k_shape", table=tag, qpt=qpt, prefetch=pf, streams=8, graph=1,
                 Mops=round(n * steps / ms / 1e3, 1))
    L.gpuhash_set_tuning(C.byref(old))
    # miss-only traffic (2 sectors per request)
    N.check(L.gpuhash_gen_queries(sd.ptr, None, SEED + 12345, pop, 1 << 24, 7, 0.0, 0.0, None))
    ms = timed_resident(geom, t.ptr, sd.ptr, 1 << 24, od.ptr, None, 0, 1, 1, 0)
    emit(exp="search_miss_bulk", table=tag, n=1 << 24, Mops=round((1 << 24) / ms / 1e3, 1), GBps_80=round((1 << 24) * 80 / ms / 1e6, 1))
    # ---- 4: insert / delete
    ins = mk.DeviceBuffer(12 * (1 << 24))
    nxt = pop
    for n in (3277, 65536, 1 << 20, 1 << 24):
        N.check(L.gpuhash_gen_inserts(ins.ptr, None, SEED, nxt, n, None)); N.check(L.gpuhash_device_sync())
        ms = timed_resident(geom, t.ptr, None, 0, None, ins.ptr, n, 1, 1, 0, reps=1)
        emit(exp="insert_fresh", table=tag, n=n, load="0.25", us=round(ms * 1e3, 2), Mops=round(n / ms / 1e3, 1))
        a, b = C.c_void_p(), C.c_void_p()
        e0, e1 = L.gpuhash_event_create(), L.gpuhash_event_create()
        L.gpuhash_event_record(e0, None)
        N.check(L.gpuhash_delete_ex(C.byref(geom), ins.ptr, t.ptr, n, None, 0, None))
        L.gpuhash_event_record(e1, None)
        f = C.c_float(); N.check(L.gpuhash_event_elapsed_ms(e0, e1, C.byref(f)))
        emit(exp="delete_present", table=tag, n=n, us=round(f.value * 1e3, 2), Mops=round(n / f.value / 1e3, 1))
        nxt += n
    sd.free(); od.free(); ins.free(); t.free()


def fill_sweep(mem_p, algo, tag):
    """insert rate and what happens to requests as the table fills to 95 %"""
    t = mk.DeviceTable(mem_p, algo)
    slots = (1 << mem_p) // 8
    step = slots // 20
    buf = mk.DeviceBuffer(12 * step)
    st = mk.DeviceStats()
    e0, e1 = L.gpuhash_event_create(), L.gpuhash_event_create()
    prev = None
    for k in range(19):
        N.check(L.gpuhash_gen_inserts(buf.ptr, None, SEED, k * step, step, None)); N.check(L.gpuhash_device_sync())
        L.gpuhash_event_record(e0, None)
        N.check(L.gpuhash_insert_flat_ex(C.byref(t.geom), t.ptr, buf.ptr, step, st.ptr, 0, None))
        L.gpuhash_event_record(e1, None)
        f = C.c_float(); N.check(L.gpuhash_event_elapsed_ms(e0, e1, C.byref(f)))
        s = st.read()
        d = {k2: (s[k2] - (prev[k2] if prev else 0)) for k2 in s if k2 != "chain_hist"}
        prev = s
        emit(exp="fill", table=tag, algo=algo, load_after=round((k + 1) / 20, 2), n=step, Mops=round(step / f.value / 1e3, 1),
             to_b2=d["ins_to_b2"], displaced=d["ins_displaced"], dropped=d["ins_dropped"], overwritten=d["ins_overwritten"],
             cas_retry=d["ins_cas_retry"], gave_up=d["ins_gave_up"])
    emit(exp="fill_chain_hist", table=tag, algo=algo, chain_hist=prev and st.read()["chain_hist"])
    buf.free(); t.free()


def main():
    mk.require_gpu()
    only = set(sys.argv[1:]) or {"gather", "search", "fill"}
    sm, l2 = C.c_int(), C.c_int(); free, total = C.c_size_t(), C.c_size_t()
    N.check(L.gpuhash_device_info(0, C.byref(sm), C.byref(l2), C.byref(free), C.byref(total)))
    emit(exp="device", sm_count=sm.value, l2_bytes=l2.value, free_gib=round(free.value / 2**30, 1), build=L.gpuhash_build_info().decode())
    if "gather" in only:
        big = mk.DeviceBuffer(1 << 34, zero=True)
        gather_sweep(big.ptr, 1 << 34, "16GiB")
        gather_sweep(big.ptr, 1 << 26, "64MiB(L2)")
        big.free()
    if "search" in only:
        for layout in (N.LAYOUT_PAIRS, N.LAYOUT_REFERENCE):
            search_sweep(34, "MEM_P34", layout)
            search_sweep(26, "MEM_P26", layout)
    if "fill" in only:
        fill_sweep(26, N.CUCKOO, "MEM_P26")
        fill_sweep(26, N.TWO_CHOICE, "MEM_P26")
    emit(exp="done")


if __name__ == "__main__":
    main()
