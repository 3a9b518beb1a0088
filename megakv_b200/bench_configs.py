"""BASELINE.json configs[2] and configs[3] as extra keys of bench.py's JSON line (SURVEY.md 8(d) Config 3 and 4).

config2   HASH_2CHOICE (gpu_hash.cu:77-229), key ranks from the reference's own zipf generator (src/zipf.h:137-183, theta
          0.99, restated bit for bit in gpuhash_workload.cu), a table that fits in L2 (MEM_P 26 = 64 MiB) against one that
          does not (MEM_P 34); searches alone (both table layouts: the layout is what the L2-resident case is sensitive to,
          gpuhash_geom_init_auto picks it) and a 50/50 search / insert mix (inserts of ranked keys = updates in place, hot
          slots contended) as one-launch scheduler cycles.  Every search word of the last mixed step is checked.
config3   HASH_CUCKOO at 90 % load (gpu_hash.cu:333-425): the table is filled to 0.9 of its slots, then every step deletes
          the oldest keys and inserts as many fresh ones (delete -> insert per batch, the reference's in-stream order;
          src/items.c:94-106 is where its deletes come from).  Reported: updates/s, and from a second pass with the
          statistics counters on: eviction-chain length histogram, displaced, dropped, CAS retries, how many deletes found
          their key (a dropped key cannot be deleted any more).

Both legs are bounded (a few seconds of GPU time) and never raise into the headline: bench.py catches and reports.
"""
import ctypes as C
import json
import os
import time

import numpy as np

BATCH = 65536
SEED = 1
THETA = 0.99
HERE = os.path.dirname(os.path.abspath(__file__))


def known_zetan(pop, theta):
    """mehcached_zeta(pop, theta) (src/zipf.h:103-115) from the committed table (2^29 terms take ~20 s of numpy), else computed"""
    from . import keystream as ks
    try:
        with open(os.path.join(HERE, "zetan_table.json")) as f:
            tab = json.load(f)
        if abs(tab["theta"] - theta) < 1e-12 and str(pop) in tab["zetan"]:
            return float(tab["zetan"][str(pop)])
    except Exception:
        pass
    return ks.ref_zetan(pop, theta)


def _events(L, N):
    a, b = L.gpuhash_event_create(), L.gpuhash_event_create()

    def timed(fn):
        N.check(L.gpuhash_device_sync())
        N.check(L.gpuhash_event_record(a, None)); fn(); N.check(L.gpuhash_event_record(b, None))
        t = C.c_float(); N.check(L.gpuhash_event_elapsed_ms(a, b, C.byref(t)))
        return t.value / 1e3

    def close():
        L.gpuhash_event_destroy(a); L.gpuhash_event_destroy(b)
    return timed, close


def _preload(L, N, mk, geom, table, first, count, stats=None):
    chunk = 1 << 24
    gen = mk.DeviceBuffer(12 * min(chunk, max(count, 1)))
    for lo in range(first, first + count, chunk):
        n = min(chunk, first + count - lo)
        N.check(L.gpuhash_gen_inserts(gen.ptr, None, SEED, lo, n, None))
        N.check(L.gpuhash_insert_flat_ex(C.byref(geom), table.ptr, gen.ptr, n, stats.ptr if stats else None, 0, None))
    N.check(L.gpuhash_device_sync())
    gen.free()


def config2(args, L, N, mk, log, sector_rate):
    steps = max(4, min(args.steps, 20))
    W = args.batches_per_step
    n_s = n_i = BATCH // 2
    free_, total_ = C.c_size_t(), C.c_size_t()
    N.check(L.gpuhash_device_info(getattr(args, "_dev", 0), None, None, C.byref(free_), C.byref(total_)))
    out = {"workload": f"configs[2]: HASH_2CHOICE, zipf theta {THETA} by the reference's generator (src/zipf.h), load factor 0.25; "
                       f"searches alone (one launch of 2^22) and a 50/50 search/insert mix as one-launch cycles of {W} batches of 64K",
           "tables": []}
    timed, close = _events(L, N)
    for mem_p in (26, 34):
        if (1 << mem_p) + (6 << 30) > free_.value:
            out["tables"].append({"mem_p": mem_p, "skipped": "not enough free device memory next to the headline table"})
            continue
        auto = N.Geom()
        N.check(L.gpuhash_geom_init_auto(C.byref(auto), mem_p, N.TWO_CHOICE))
        pop = (1 << mem_p) // 32
        zetan = known_zetan(pop, THETA)
        row = {"mem_p": mem_p, "table_MiB": (1 << mem_p) >> 20, "population": pop,
               "layout_auto": "reference bytes (one thread per request, location word on a hit only)" if auto.layout == N.LAYOUT_REFERENCE else "pairs",
               "search_only_Mops": {}}
        n_bulk = 1 << 22
        sel_b = mk.DeviceBuffer(8 * n_bulk); res_b = mk.DeviceBuffer(8 * n_bulk); sel_u = mk.DeviceBuffer(8 * n_bulk)
        N.check(L.gpuhash_gen_requests_ref_zipf(sel_b.ptr, None, SEED, pop, n_bulk, 4242, 0, THETA, zetan, 0, None))
        for layout, name in ((N.LAYOUT_PAIRS, "pairs"), (N.LAYOUT_REFERENCE, "reference")):
            if mem_p > 30 and layout != auto.layout:
                continue                                              # the 16 GiB table is built once, in the layout chosen for it
            geom = N.Geom.from_buffer_copy(bytes(auto)); geom.layout = layout
            table = mk.DeviceBuffer(L.gpuhash_table_bytes(C.byref(geom)), zero=True)
            _preload(L, N, mk, geom, table, 0, pop)
            run = lambda: N.check(L.gpuhash_search_ex(C.byref(geom), sel_b.ptr, res_b.ptr, table.ptr, n_bulk, None, None))
            timed(run)
            t = min(timed(run) for _ in range(3))
            row["search_only_Mops"][name] = round(n_bulk / t / 1e6, 1)
            N.check(L.gpuhash_gen_queries(sel_u.ptr, None, SEED, pop, n_bulk, 4243, 0.0, 0.0, None))
            run_u = lambda: N.check(L.gpuhash_search_ex(C.byref(geom), sel_u.ptr, res_b.ptr, table.ptr, n_bulk, None, None))
            timed(run_u)
            row.setdefault("search_only_uniform_Mops", {})[name] = round(n_bulk / min(timed(run_u) for _ in range(3)) / 1e6, 1)
            if True:
                # ---- the mix (both layouts where both tables are built): K steps of W batches of (32768 searches + 32768 updates)
                kd = steps * W
                s_d = mk.DeviceBuffer(8 * n_s * kd); o_d = mk.DeviceBuffer(8 * n_s * kd); e_d = mk.DeviceBuffer(4 * n_s * kd)
                i_d = mk.DeviceBuffer(12 * n_i * kd)
                N.check(L.gpuhash_gen_requests_ref_zipf(s_d.ptr, e_d.ptr, SEED, pop, n_s * kd, 99, 0, THETA, zetan, 0, None))
                N.check(L.gpuhash_gen_requests_ref_zipf(i_d.ptr, None, SEED, pop, n_i * kd, 777, 0, THETA, zetan, 1, None))
                res = N.BenchResult()
                for _ in range(3):
                    N.check(L.gpuhash_bench_cycles(C.byref(geom), table.ptr, s_d.ptr, n_s, o_d.ptr, i_d.ptr, n_i, W, steps, 2, C.byref(res)),
                            "gpuhash_bench_cycles")
                row.setdefault("mixed_Mops", {})[name] = round(steps * W * (n_s + n_i) / (res.total_ms / 1e3) / 1e6, 1)
                row.setdefault("mixed_ms_per_step", {})[name] = round(res.total_ms / steps, 4)
                last = (steps - 1) * W * n_s
                got = np.empty(2 * W * n_s, dtype=np.uint32); exp = np.empty(W * n_s, dtype=np.uint32)
                N.check(L.gpuhash_d2h(got.ctypes.data, o_d.ptr + 8 * last, got.nbytes, None))
                N.check(L.gpuhash_d2h(exp.ctypes.data, e_d.ptr + 4 * last, exp.nbytes, None)); N.check(L.gpuhash_device_sync())
                o0, o1 = got[0::2], got[1::2]
                good = ((o0 == exp) & ((o1 == 0) | (o1 == exp))) | ((o1 == exp) & (o0 == 0))
                row["searches_checked"] = row.get("searches_checked", 0) + int(len(exp)); row["mismatches"] = row.get("mismatches", 0) + int((~good).sum())
                row["distinct_keys_in_a_step"] = int(len(np.unique(exp)))
                for b in (s_d, o_d, e_d, i_d):
                    b.free()
            table.free()
        row["search_frac_of_probe_ceiling"] = {k: round(v * 1e6 * 2 / sector_rate, 3) for k, v in row["search_only_Mops"].items()} if mem_p > 30 else None
        sel_b.free(); res_b.free(); sel_u.free()
        out["tables"].append(row)
        log(f"config2 MEM_P {mem_p}: {row}")
    close()
    return out


def config3(args, L, N, mk, log):
    mem_p = 30
    W = args.batches_per_step
    steps = max(4, min(args.steps, 20))
    n_u = BATCH // 2                                                 # per batch: 32768 deletes + 32768 inserts
    geom = N.Geom()
    N.check(L.gpuhash_geom_init(C.byref(geom), mem_p, N.CUCKOO))
    slots = (1 << mem_p) // 8
    live = int(slots * 0.9)
    table = mk.DeviceBuffer(L.gpuhash_table_bytes(C.byref(geom)), zero=True)
    st_fill = mk.DeviceStats()
    t0 = time.time()
    _preload(L, N, mk, geom, table, 0, live, st_fill)
    fill = st_fill.read()
    log(f"config3: filled 2^{mem_p} B to 0.9 ({live} keys) in {time.time() - t0:.2f} s; dropped {fill['ins_dropped']}")
    per_step = W * n_u
    kd = 2 * steps + 2                                               # a timed pass and a counted pass (+ warm-up), every step its own keys
    d_d = mk.DeviceBuffer(12 * per_step * kd); i_d = mk.DeviceBuffer(12 * per_step * kd)
    # step k deletes the keys step k-1 inserted (step 0: the last keys of the fill) and inserts fresh ones: the load factor stays
    # at 0.9 (a delete only misses if its key was dropped in between).  Deleting the OLDEST keys instead lets the table creep
    # to 100 %: after a few million drops most of them are gone already, their deletes free nothing, and every second insert
    # ends in a drop (measured: 0.45 drops per insert, 42 % of the deletes find their key).
    N.check(L.gpuhash_gen_inserts(d_d.ptr, None, SEED, live - per_step, per_step * kd, None))
    N.check(L.gpuhash_gen_inserts(i_d.ptr, None, SEED, live, per_step * kd, None))
    ws = mk.DeviceBuffer(L.gpuhash_cycle_workspace_bytes(W), zero=True)
    descs_d = mk.DeviceBuffer(C.sizeof(N.Batch) * W * kd)
    all_descs = (N.Batch * (W * kd))()
    for k in range(kd):
        for w in range(W):
            off = 12 * n_u * (k * W + w)
            all_descs[k * W + w] = N.Batch(None, None, d_d.ptr + off, i_d.ptr + off, 0, n_u, n_u, 0)
    descs_d.upload(np.frombuffer(bytes(all_descs), dtype=np.uint8))
    host_descs = [(N.Batch * W)(*all_descs[k * W:(k + 1) * W]) for k in range(kd)]

    def run_steps(first, count, stats):
        for k in range(first, first + count):
            N.check(L.gpuhash_cycle_multi_ex(C.byref(geom), table.ptr, host_descs[k], descs_d.ptr + C.sizeof(N.Batch) * W * k, W, 0, ws.ptr,
                                             stats.ptr if stats else None, None), "gpuhash_cycle_multi_ex")

    timed, close = _events(L, N)
    run_steps(0, 2, None)                                            # warm-up
    t = timed(lambda: run_steps(2, steps, None))
    st = mk.DeviceStats()
    t_counted = timed(lambda: run_steps(2 + steps, steps, st))
    close()
    err = L.gpuhash_cycle_error(1)
    s = st.read()
    n_ins = steps * per_step
    # occupancy after the churn, counted on the host (pair layout: word 2l of a bucket is slot l's signature)
    tab = table.download(np.uint32)
    occupancy = float((tab[0::2] != 0).mean())
    del tab
    # how many of the keys that should be live are still findable (a dropped key is gone): a sample of the newest inserts
    n_chk = 1 << 20
    newest_first = live + per_step * kd - n_chk
    sel = mk.DeviceBuffer(8 * n_chk); res = mk.DeviceBuffer(8 * n_chk)
    N.check(L.gpuhash_gen_inserts(None, sel.ptr, SEED, newest_first, n_chk, None))
    N.check(L.gpuhash_search_ex(C.byref(geom), sel.ptr, res.ptr, table.ptr, n_chk, None, None)); N.check(L.gpuhash_device_sync())
    r = res.download(np.uint32)
    want = np.arange(newest_first + 1, newest_first + n_chk + 1, dtype=np.uint64).astype(np.uint32)
    found = (r[0::2] == want) | (r[1::2] == want)
    wrong = ((r[0::2] != 0) & (r[0::2] != want)) | ((r[1::2] != 0) & (r[1::2] != want))
    out = {
        "workload": f"configs[3]: HASH_CUCKOO, table 2^{mem_p} bytes filled to load factor 0.9 ({live} keys), then steady churn: per step "
                    f"{W} batches x ({n_u} deletes of the keys the previous step inserted -> {n_u} inserts of fresh keys), one launch per step",
        "mem_p": mem_p, "load_factor": 0.9, "steps": steps, "updates_per_step": 2 * per_step,
        "churn_Mops": round(steps * 2 * per_step / t / 1e6, 1), "ms_per_step": round(t / steps * 1e3, 4),
        "churn_Mops_with_counters_on": round(steps * 2 * per_step / t_counted / 1e6, 1),
        "fill": {"inserted": live, "dropped": fill["ins_dropped"], "displaced": fill["ins_displaced"], "to_bucket2": fill["ins_to_b2"],
                 "cas_retries": fill["ins_cas_retry"], "chain_hist": fill["chain_hist"]},
        "churn_counters": {"inserts": n_ins, "chain_hist": s["chain_hist"], "displaced": s["ins_displaced"], "dropped": s["ins_dropped"],
                           "to_bucket2": s["ins_to_b2"], "cas_retries": s["ins_cas_retry"], "gave_up": s["ins_gave_up"],
                           "displaced_per_insert": round(s["ins_displaced"] / n_ins, 4), "dropped_per_insert": round(s["ins_dropped"] / n_ins, 6),
                           "deletes": n_ins, "deletes_that_found_their_key": s["del_requests_hit"]},
        "occupancy_after_churn": round(occupancy, 4),
        "newest_keys_findable": round(float(found.mean()), 6), "searches_answered_by_another_key_with_the_same_signature": int(wrong.sum()),
        "phase_wait_timeouts": int(err),
    }
    for b in (table, d_d, i_d, ws, descs_d, sel, res, st, st_fill):
        b.free()
    log(f"config3: {out}")
    return out


def run(args, L, N, mk, local_rank, log, sector_rate):
    args._dev = local_rank
    res = {}
    for name, fn in (("config2", lambda: config2(args, L, N, mk, log, sector_rate)), ("config3", lambda: config3(args, L, N, mk, log))):
        if getattr(args, "config", 0) and name != f"config{args.config}":
            continue
        try:
            res[name] = fn()
        except Exception as e:
            res[name] = {"failed": repr(e)}
    return res
