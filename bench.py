#!/usr/bin/env python
"""bench.py -- batched search/insert Mops/s of the hash-index hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): HASH_CUCKOO, table 2^34 bytes (>> L2), preloaded to load factor 0.25
(2^29 uniform keys), then K steps; one step = one batch of 65 536 requests = 62 259 searches (95 %, keys drawn
uniformly from the preloaded population, so every search hits) + 3 277 inserts (5 %, fresh keys).  Per batch:
one gpu_hash_search-equivalent launch, then one insert launch on the same stream (the reference's in-stream
order, mega_scheduler.c:392-502); batches go round-robin over S = 64 streams like the reference's per-worker streams
(mega_scheduler.c:276-280).  Every step has its own input/output arrays in HBM (K * 1 MiB >> L2).

  value     whole-job Mops/s with the batches already resident in HBM (CUDA events around the K steps)
  e2e       the same K steps through the host-buffer C ABI (gpuhash_index_submit): pinned host -> device ->
            kernels -> pinned host inside the timed region
  roofline  search kernel only, same batches/streams: algorithmic bytes (8 B request + 2 x 32 B
            signature sectors + 32 B location sector per hit bucket + 8 B result) / time, against
            MEASURED_PEAKS.json hbm_gbs; plus one bulk launch and the measured random-32 B-sector ceiling
  cpu_baseline  oracle (C restatement of the reference's algorithm), one core, bounded sample of the same steps

  --impl reference : the reference's algorithm on the host cores (oracle, all threads) -- same metric/config.
"""
import argparse
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # as many hardware queues as the driver offers (before CUDA starts)
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 65536
N_SEARCH = 62259            # 95 %
N_INSERT = BATCH - N_SEARCH  # 3277 = 5 %
SEED = 1


def quiet_stdout():
    """Libraries print to stdout (NCCL: "NCCL version ..."); the contract is ONE JSON line there.  Everything written to
    fd 1 from now on goes to stderr; emit() writes the line to the real stdout.  (The saved descriptor lives on the sys
    module: this file is both __main__ and, for the multi-GPU arm, the imported module `bench`.)"""
    if getattr(sys, "_megakv_real_stdout", None) is None:
        sys.stdout.flush()
        sys._megakv_real_stdout = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    fd = getattr(sys, "_megakv_real_stdout", None)
    os.write(fd if fd is not None else 1, data)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (pynvml, 5 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._run = [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "applications_clocks_setting": 0x2}
        while self._run:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self._run = True
            self.t = threading.Thread(target=self._loop, daemon=True); self.t.start()
        return self

    def __exit__(self, *a):
        if self.nv:
            self._run = False; self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU arms

def cpu_workload(mem_p, preload_log2, steps, threads, log, warm=0):
    """The same step on the host: oracle table of 2^mem_p bytes preloaded with 2^preload_log2 keys, then `steps`
    batches of 62 259 searches (all hits) + 3 277 fresh inserts.  Returns (Mops/s, seconds, cores)."""
    from oracle import pyoracle as po
    o = po.Oracle(mem_p, po.CUCKOO)
    pop = 1 << preload_log2
    t0 = time.time()
    chunk = 1 << 22
    for first in range(0, pop, chunk):
        iel, _ = po.keys(SEED, first, min(chunk, pop - first))
        o.insert_mt(iel, 8)
    log(f"cpu: preloaded 2^{preload_log2} keys into a 2^{mem_p} B table in {time.time() - t0:.1f} s")
    rng = np.random.default_rng(7)
    all_sel = []
    steps += warm
    idx = rng.integers(0, pop, size=(steps, N_SEARCH), dtype=np.int64)
    from megakv_b200 import keystream as ks                     # request derivation only (numpy)
    for s in range(steps):
        all_sel.append(ks.keys_to_requests(ks._keys_at(SEED, idx[s])))
    ins = [po.keys(SEED, pop + s * N_INSERT, N_INSERT)[0] for s in range(steps)]
    t0 = po.now()
    for s in range(steps):
        if s == warm:
            t0 = po.now()
        if threads == 1:
            out = o.search(all_sel[s]); o.insert(ins[s])
        else:
            out = o.search_mt(all_sel[s], threads); o.insert_mt(ins[s], threads)
    dt = po.now() - t0
    assert ((out[0::2] != 0) | (out[1::2] != 0)).all()
    return (steps - warm) * BATCH / dt / 1e6, dt, threads


def host_mem_p(want):
    """largest table the host can hold next to everything else (the oracle table lives in host RAM)"""
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 8 << 30
    p = want
    while p > 20 and (1 << p) * 1.3 > avail:
        p -= 1
    return p


def run_reference_arm(args, log):
    cores = os.cpu_count() or 1
    mem_p = host_mem_p(args.mem_p)
    steps = max(1, min(args.steps, 512))
    warm = max(0, min(args.warmup, 8))
    val, dt, thr = cpu_workload(mem_p, min(26, mem_p - 7), steps, cores, log, warm=warm)
    line = {
        "impl": "reference", "metric": "batched search/insert Mops/s (95/5 GET/SET, uniform keys)",
        "value": round(val, 3), "unit": "Mops/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": round(1e3 * dt / steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(mem_p, args, note="CPU: oracle (C restatement of gpu_hash.cu), all host threads; "
                                  f"table 2^{mem_p} B preloaded with 2^{min(26, mem_p - 7)} keys"),
        "cpu_baseline": {"value": round(val, 3), "unit": "Mops/s", "cores": thr, "kind": "port",
                         "sample": f"{steps} steps of 65536 requests after {warm} warm-up steps (search_mt + insert_mt), wall clock"},
        "e2e": {"value": round(val, 3), "unit": "Mops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(mem_p, args, note=None):
    c = {"workload": f"configs[1]: 1xB200 search-heavy 95/5 GET/SET, uniform keys, HASH_CUCKOO, table 2^{mem_p} bytes, "
                     f"batch 64K signatures ({N_SEARCH} searches + {N_INSERT} inserts per step)",
         "mem_p": mem_p, "algo": "HASH_CUCKOO", "load_factor": 0.25, "batch": BATCH,
         "streams": args.streams, "cuda_graph": bool(args.graph),
         "cache": "every step has its own request/result arrays (K x 1 MiB > L2) and the table is 16 GiB >> 126 MB L2"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------ GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mem-p", type=int, default=34)
    ap.add_argument("--streams", type=int, default=64,
                    help="batches in flight for the resident leg (8: 12.7, 16: 16.0, 32: 18.1, 64: 19.9, 128: 20.5 Gops/s); "
                         "the e2e leg uses at most 32 workers (more only add host-link contention)")
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ops", action="store_true", help="skip the per-operation bulk launches")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference's own search kernel (oracle/_ref)")
    ap.add_argument("--no-ring", action="store_true", help="skip the persistent-kernel ring variant of the e2e leg")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def log(msg):
        if args.verbose or os.environ.get("BENCH_VERBOSE"):
            print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    quiet_stdout()
    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, log)
        return 0

    if world > 1 or os.environ.get("GPUHASH_FORCE_SHARDED"):      # (the env switch: routed path on one GPU, for profiling)
        from megakv_b200 import sharded_bench                     # multi-GPU arm: sharded index + all-to-all routing
        return sharded_bench.main(args, rank, world, local_rank, log)

    import megakv_b200 as mk
    from megakv_b200 import _native as N
    L = mk.lib()
    mk.require_gpu()
    N.check(L.gpuhash_set_device(local_rank))

    steps, warm = max(1, args.steps), max(3, args.warmup)
    mem_p = args.mem_p
    free, total = C.c_size_t(), C.c_size_t()
    N.check(L.gpuhash_device_info(local_rank, None, None, C.byref(free), C.byref(total)))
    while (1 << mem_p) + (8 << 30) > free.value and mem_p > 26:
        mem_p -= 1
    S = max(1, min(args.streams, 128))
    ix = L.gpuhash_index_create(mem_p, N.CUCKOO, min(S, 32), N_SEARCH, N_INSERT, 1)
    if not ix:
        raise mk.GpuHashError("gpuhash_index_create failed")
    geom = L.gpuhash_index_geom(ix).contents
    table = L.gpuhash_index_table(ix)

    # ---- preload to load factor 0.25 with device-generated keys
    pop = (1 << mem_p) // 8 // 4
    chunk = 1 << 24
    gen_d = mk.DeviceBuffer(12 * chunk)
    t0 = time.time()
    for first in range(0, pop, chunk):
        n = min(chunk, pop - first)
        N.check(L.gpuhash_gen_inserts(gen_d.ptr, None, SEED, first, n, None))
        N.check(L.gpuhash_insert_flat_ex(C.byref(geom), table, gen_d.ptr, n, None, 0, None))
    N.check(L.gpuhash_device_sync())
    log(f"preloaded {pop} keys (LF 0.25) into 2^{mem_p} B in {time.time() - t0:.2f} s")
    gen_d.free()

    # ---- K_d distinct batches resident in HBM
    kd = min(steps + warm, 4096)
    search_d = mk.DeviceBuffer(8 * N_SEARCH * kd)
    out_d = mk.DeviceBuffer(8 * N_SEARCH * kd)
    insert_d = mk.DeviceBuffer(12 * N_INSERT * kd)
    N.check(L.gpuhash_gen_queries(search_d.ptr, None, SEED, pop, N_SEARCH * kd, 99, 0.0, 0.0, None))
    N.check(L.gpuhash_gen_inserts(insert_d.ptr, None, SEED, pop, N_INSERT * kd, None))
    N.check(L.gpuhash_device_sync())

    def resident(first, count, n_search=N_SEARCH, n_insert=N_INSERT):
        """`count` steps starting at batch `first` (wrapping inside the kd resident batches); returns seconds"""
        total_ms, done = 0.0, 0
        while done < count:
            b0 = (first + done) % kd
            c = min(count - done, kd - b0)
            res = N.BenchResult()
            N.check(L.gpuhash_bench_resident(C.byref(geom), table,
                                             search_d.ptr + 8 * N_SEARCH * b0, n_search, out_d.ptr + 8 * N_SEARCH * b0,
                                             insert_d.ptr + 12 * N_INSERT * b0, n_insert, c, S, args.graph, C.byref(res)),
                    "gpuhash_bench_resident")
            total_ms += res.total_ms; done += c
        return total_ms / 1e3

    # replayed as a graph, one launch per operation kind overlaps better across streams than the single-launch cycle
    # (tools/exp_mixed.py: 19.3 vs 14.1 Gops/s); issued call by call it is the other way round (11.4 vs 13.5)
    fused_resident = int(os.environ["BENCH_FUSED"]) if os.environ.get("BENCH_FUSED") else (0 if args.graph else 1)
    L.gpuhash_set_tuning(C.byref(N.Tune(0, 0, 4, fused_resident)))
    sampler = ClockSampler(local_rank)
    resident(0, warm)                                               # W untimed warm-up steps
    with sampler:
        t_val = resident(warm, steps)                               # EXACTLY K timed steps
    value = steps * BATCH / t_val / 1e6
    log(f"resident: {steps} steps in {t_val * 1e3:.2f} ms -> {value:.1f} Mops/s")

    # ---- sanity on the timed output: every search of the last timed batch hit (loc != 0 in one of the words)
    last = (warm + steps - 1) % kd
    chk = np.empty(2 * N_SEARCH, dtype=np.uint32)
    N.check(L.gpuhash_d2h(chk.ctypes.data, out_d.ptr + 8 * N_SEARCH * last, chk.nbytes, None)); N.check(L.gpuhash_device_sync())
    hit_frac = float(((chk[0::2] != 0) | (chk[1::2] != 0)).mean())
    hits_per_search = float(((chk[0::2] != 0).sum() + (chk[1::2] != 0).sum()) / N_SEARCH)
    assert hit_frac > 0.999, f"timed searches did not hit: {hit_frac}"

    # ---- roofline: the search kernel alone on the same batches and streams
    peak, peak_src = peaks()
    bytes_per_search = 8 + 2 * 32 + 32 * hits_per_search + 8
    resident(0, min(warm, 50), n_insert=0)
    with sampler:
        t_s = resident(warm, steps, n_insert=0)
    achieved = steps * N_SEARCH * bytes_per_search / t_s / 1e9
    # one bulk launch (2^24 requests) -- the kernel without launch/ramp effects
    bulk_n = min(1 << 24, N_SEARCH * kd)
    res = N.BenchResult()
    for _ in range(3):
        N.check(L.gpuhash_bench_resident(C.byref(geom), table, search_d.ptr, bulk_n, out_d.ptr, None, 0, 1, 1, 0, C.byref(res)))
    bulk_gbs = bulk_n * bytes_per_search / (res.total_ms / 1e3) / 1e9
    bulk_mops = bulk_n / (res.total_ms / 1e3) / 1e6
    # measured random-sector ceiling on the same allocation
    ms = C.c_float()
    N.check(L.gpuhash_roofline_gather(table, 1 << mem_p, 1 << 27, 0, 4, 3, C.byref(ms), None))
    sector_rate = (1 << 27) / (ms.value / 1e3)                       # 32 B sectors per second
    traffic, traffic_src = None, None                                # DRAM bytes per launch of that kernel, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": round(N_SEARCH * bytes_per_search),
            "kernel": "gh::search_warp_kernel<pairs>", "peak_source": peak_src, "bytes_per_search": round(bytes_per_search, 2),
            "launches": steps, "avg_launch_us_effective": round(t_s / steps * 1e6, 3),
            "bulk_launch": {"requests": bulk_n, "GB/s": round(bulk_gbs, 1), "Mops/s": round(bulk_mops, 1),
                            "frac": round(bulk_gbs / peak, 4)},
            "random_sector_probe": {"Gsectors/s": round(sector_rate / 1e9, 2), "GB/s": round(sector_rate * 32 / 1e9, 1),
                                    "search_frac_of_probe": round(steps * N_SEARCH * (2 + hits_per_search) / t_s / sector_rate, 4),
                                    "bulk_frac_of_probe": round(bulk_n * (2 + hits_per_search) / (res.total_ms / 1e3) / sector_rate, 4)}}

    # ---- the reference's OWN search kernel on the same GPU, same table, same batches (SURVEY 8d iii): gpu_hash.cu compiled
    #      where it lies in legacy-warp mode (oracle/Makefile -> oracle/_ref/, the only way its __ballot assembles), driven
    #      with the reference's launch shape (24576 threads, 256 per block, mega.c:163-165) and the caller's memset
    #      (mega_scheduler.c:406).  A reported baseline and a full-size parity check (its result words == ours), nothing more;
    #      its batch insert kernel hangs on sm_100, so only search is timed.
    ref_gpu = None
    ref_so = os.path.join(ROOT, "oracle", "_ref", f"libgpuhash_ref_cuckoo_{mem_p}.so")
    if os.path.exists(ref_so) and not args.no_ref_gpu and not args.no_cpu:    # part of the baseline leg: the only one that may run oracle/
        try:
            R = C.CDLL(ref_so)
            R.gpu_hash_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
            R.gpu_hash_search.restype = None
            ours = np.empty(2 * N_SEARCH, dtype=np.uint32)
            N.check(L.gpuhash_search_ex(C.byref(geom), search_d.ptr, out_d.ptr, table, N_SEARCH, None, None)); N.check(L.gpuhash_device_sync())
            N.check(L.gpuhash_d2h(ours.ctypes.data, out_d.ptr, ours.nbytes, None)); N.check(L.gpuhash_device_sync())
            N.check(L.gpuhash_table_convert(C.byref(geom), table, N.LAYOUT_REFERENCE, None)); N.check(L.gpuhash_device_sync())
            try:
                nb = min(kd, 200)

                def ref_batches():
                    for b in range(nb):
                        N.check(L.gpuhash_dev_memset(out_d.ptr + 8 * N_SEARCH * b, 0, 8 * N_SEARCH, None))
                        R.gpu_hash_search(search_d.ptr + 8 * N_SEARCH * b, out_d.ptr + 8 * N_SEARCH * b, table, N_SEARCH, 24576, 256, None)
                ref_batches(); N.check(L.gpuhash_device_sync())
                t0 = time.perf_counter(); ref_batches(); N.check(L.gpuhash_device_sync()); t_ref = time.perf_counter() - t0
                # and the way the reference's scheduler issues them: one stream per worker, 16 at most (mega_scheduler.c:276-280)
                st16 = [L.gpuhash_stream_create() for _ in range(16)]

                def ref_batches_16():
                    for b in range(nb):
                        s_ = st16[b % 16]
                        N.check(L.gpuhash_dev_memset(out_d.ptr + 8 * N_SEARCH * b, 0, 8 * N_SEARCH, s_))
                        R.gpu_hash_search(search_d.ptr + 8 * N_SEARCH * b, out_d.ptr + 8 * N_SEARCH * b, table, N_SEARCH, 24576, 256, s_)
                ref_batches_16(); N.check(L.gpuhash_device_sync())
                t0 = time.perf_counter(); ref_batches_16(); N.check(L.gpuhash_device_sync()); t_ref16 = time.perf_counter() - t0
                for s_ in st16:
                    L.gpuhash_stream_destroy(s_)
                theirs = np.empty(2 * N_SEARCH, dtype=np.uint32)
                N.check(L.gpuhash_d2h(theirs.ctypes.data, out_d.ptr, theirs.nbytes, None)); N.check(L.gpuhash_device_sync())
                ref_gpu = {"what": "pzrq/megakv hash_search (gpu_hash.cu:28-75) as compiled by oracle/Makefile (compute_60 PTX -> sm_100), launch shape "
                                   "24576 x 256 + memset per batch, one stream, device-resident batches, host wall clock over "
                                   f"{nb} batches",
                           "Mops/s": round(nb * N_SEARCH / t_ref / 1e6, 1), "us_per_batch": round(t_ref / nb * 1e6, 2),
                           "Mops/s_16_streams": round(nb * N_SEARCH / t_ref16 / 1e6, 1),
                           "results_equal_ours": bool(np.array_equal(ours, theirs))}
                log(f"reference GPU search kernel: {ref_gpu['Mops/s']} Mops/s on one stream, {ref_gpu['Mops/s_16_streams']} on 16; "
                    f"results equal ours: {ref_gpu['results_equal_ours']}")
            finally:
                as_ref = N.Geom.from_buffer_copy(bytes(geom)); as_ref.layout = N.LAYOUT_REFERENCE
                N.check(L.gpuhash_table_convert(C.byref(as_ref), table, geom.layout, None)); N.check(L.gpuhash_device_sync())
        except Exception as e:                                       # a baseline must never take the product line down
            ref_gpu = {"failed": str(e)}

    # ---- every operation on its own, uniform and zipf(0.99) keys, one bulk launch each (north star: search, insert and
    #      delete Mops/s, absolute and against the random-access roofline).  Algorithmic sectors per op: 3.
    ops = {}
    if not args.no_ops:
        from megakv_b200 import keystream as ks
        ev_a, ev_b = L.gpuhash_event_create(), L.gpuhash_event_create()

        def timed(fn):
            N.check(L.gpuhash_device_sync())
            N.check(L.gpuhash_event_record(ev_a, None)); N.check(fn()); N.check(L.gpuhash_event_record(ev_b, None))
            t = C.c_float(); N.check(L.gpuhash_event_elapsed_ms(ev_a, ev_b, C.byref(t)))
            return t.value / 1e3

        n_ops = min(1 << 22, N_SEARCH * kd)
        zn = ks.zetan(pop, 0.99)
        req_d = mk.DeviceBuffer(12 * n_ops)
        sel_d = mk.DeviceBuffer(8 * n_ops); res_d = mk.DeviceBuffer(8 * n_ops)

        def report(name, secs):
            ops[name] = {"Mops/s": round(n_ops / secs / 1e6, 1), "frac_of_sector_roofline": round(n_ops * 3 / secs / sector_rate, 3)}

        for dist_name, theta, z in (("uniform", 0.0, 0.0), ("zipf0.99", 0.99, zn)):
            N.check(L.gpuhash_gen_queries(sel_d.ptr, None, SEED, pop, n_ops, 4242, theta, z, None))
            timed(lambda: L.gpuhash_search_ex(C.byref(geom), sel_d.ptr, res_d.ptr, table, n_ops, None, None))
            report(f"search_{dist_name}", timed(lambda: L.gpuhash_search_ex(C.byref(geom), sel_d.ptr, res_d.ptr, table, n_ops, None, None)))
        # insert: fresh uniform keys (claims), then zipf draws from the population (updates in place, hot slots contended)
        fresh0 = pop + N_INSERT * (kd + 4 * 1024) + (1 << 26)
        N.check(L.gpuhash_gen_inserts(req_d.ptr, None, SEED, fresh0, n_ops, None))
        report("insert_uniform", timed(lambda: L.gpuhash_insert_flat_ex(C.byref(geom), table, req_d.ptr, n_ops, None, 0, None)))
        report("delete_uniform", timed(lambda: L.gpuhash_delete_ex(C.byref(geom), req_d.ptr, table, n_ops, None, 0, None)))   # removes them again
        N.check(L.gpuhash_gen_requests(req_d.ptr, SEED, pop, n_ops, 777, 0.99, zn, None))
        report("insert_zipf0.99", timed(lambda: L.gpuhash_insert_flat_ex(C.byref(geom), table, req_d.ptr, n_ops, None, 0, None)))
        report("delete_zipf0.99", timed(lambda: L.gpuhash_delete_ex(C.byref(geom), req_d.ptr, table, n_ops, None, 0, None)))
        N.check(L.gpuhash_insert_flat_ex(C.byref(geom), table, req_d.ptr, n_ops, None, 0, None))   # put the deleted hot keys back
        N.check(L.gpuhash_device_sync())
        ops["requests_per_launch"] = n_ops
        req_d.free(); sel_d.free(); res_d.free()
        L.gpuhash_event_destroy(ev_a); L.gpuhash_event_destroy(ev_b)

    # ---- e2e: pinned host buffers through gpuhash_index_submit (H2D + D2H inside the timed region)
    ke = min(steps, 1024)
    hs = L.gpuhash_host_alloc(8 * N_SEARCH * ke); ho = L.gpuhash_host_alloc(8 * N_SEARCH * ke); hi = L.gpuhash_host_alloc(12 * N_INSERT * ke)
    if not (hs and ho and hi):
        raise mk.GpuHashError("pinned host allocation failed")
    N.check(L.gpuhash_d2h(hs, search_d.ptr, 8 * N_SEARCH * ke, None)); N.check(L.gpuhash_device_sync())
    ho_np = np.ctypeslib.as_array(C.cast(ho, C.POINTER(C.c_uint32)), shape=(2 * N_SEARCH * ke,))
    next_key = [pop + N_INSERT * (kd + ke)]

    def fresh_inserts():
        """new keys for every e2e pass, so inserts stay inserts (not updates)"""
        N.check(L.gpuhash_gen_inserts(insert_d.ptr, None, SEED, next_key[0], N_INSERT * ke, None))
        N.check(L.gpuhash_d2h(hi, insert_d.ptr, 12 * N_INSERT * ke, None)); N.check(L.gpuhash_device_sync())
        next_key[0] += N_INSERT * ke

    def e2e_pass(count, graph):
        t, done = 0.0, 0
        while done < count:
            c = min(ke, count - done)
            fresh_inserts()
            r = N.BenchResult()
            N.check(L.gpuhash_bench_e2e(ix, hs, N_SEARCH, ho, hi, N_INSERT, c, graph, C.byref(r)), "gpuhash_bench_e2e")
            t += r.total_ms / 1e3; done += c
        return t

    # four ways through the same host-buffer call: staging copies or zero-copy (kernels read/write the pinned host
    # buffers over PCIe themselves), each launched call by call or replayed as one CUDA graph per pass
    variants = {}
    tune_bench = N.Tune(); L.gpuhash_get_tuning(C.byref(tune_bench))
    for zero_copy, graph, fused, name in [(0, 0, None, "staged"), (0, 1, None, "staged+graph"), (1, 0, None, "zero_copy"),
                                          (1, 1, None, "zero_copy+graph"), (1, 1, 1, "zero_copy+graph+one_launch")]:
        # one_launch: the whole cycle of a worker (searches, then inserts) is ONE kernel (gpuhash_cycle_ex)
        L.gpuhash_set_tuning(C.byref(N.Tune(0, 0, 4, tune_bench.fused_cycle if fused is None else fused)))
        L.gpuhash_index_set_zero_copy(ix, zero_copy)
        e2e_pass(min(ke, max(3, warm)), graph)                       # warm-up
        ho_np[:] = 0
        with sampler:
            w0 = time.time(); t_e = e2e_pass(steps, graph); wall = time.time() - w0
        ok = float(((ho_np[0::2] != 0) | (ho_np[1::2] != 0)).mean())
        assert ok > 0.999, f"e2e ({name}) results did not come back: {ok}"
        variants[name] = {"Mops/s": round(steps * BATCH / t_e / 1e6, 1), "ms": round(t_e * 1e3, 3), "wall_ms": round(wall * 1e3, 2)}
        log(f"e2e {name}: {steps} steps in {t_e * 1e3:.2f} ms (wall {wall * 1e3:.1f}) -> {steps * BATCH / t_e / 1e6:.1f} Mops/s")
    L.gpuhash_set_tuning(C.byref(tune_bench))
    # the consumer only ever takes one of the two result words (mega_send.c:411-414): let the device choose and send 4 B
    # per search back instead of 8.  Reported next to the others, not as the headline: it changes what search_out holds.
    L.gpuhash_index_set_zero_copy(ix, 1); L.gpuhash_index_set_compact_results(ix, 1)
    e2e_pass(min(ke, max(3, warm)), 1)
    ho_np[:] = 0
    with sampler:
        t_c = e2e_pass(steps, 1)
    okc = float((ho_np[: N_SEARCH * min(ke, steps)] != 0).mean())
    assert okc > 0.999, f"e2e (compact) results did not come back: {okc}"
    compact_info = {"Mops/s": round(steps * BATCH / t_c / 1e6, 1), "d2h_bytes_per_step": 4 * N_SEARCH,
                    "what": "zero_copy+graph with one result word per search (gpuhash_index_set_compact_results)"}
    log(f"e2e zero_copy+graph, compact results: {steps * BATCH / t_c / 1e6:.1f} Mops/s")
    L.gpuhash_index_set_compact_results(ix, 0)
    L.gpuhash_index_set_zero_copy(ix, 0)
    # fifth way: no launches at all -- descriptor rings in pinned memory feeding the persistent kernel (north star (c)).
    # Timed by the host's wall clock (there is no launch to bracket with events), so it carries the submit loop too.
    ring_info = None
    if not args.no_ring:
        q = L.gpuhash_ring_create(C.byref(geom), table, 8, 4, 4, 2000)
        if q:
            try:
                def ring_pass(count, reps=0):
                    t, done, rtt = 0.0, 0, C.c_float(0)
                    while done < count:
                        c = min(ke, count - done)
                        fresh_inserts()
                        r = N.BenchResult()
                        N.check(L.gpuhash_bench_ring(q, hs, N_SEARCH, ho, hi, N_INSERT, c, C.byref(r), reps if done == 0 else 0, C.byref(rtt)),
                                "gpuhash_bench_ring")
                        t += r.total_ms / 1e3; done += c
                    return t, rtt.value
                ring_pass(min(ke, max(3, warm)))
                ho_np[:] = 0
                with sampler:
                    t_e, rtt = ring_pass(steps, 64)
                ok = float(((ho_np[0::2] != 0) | (ho_np[1::2] != 0)).mean())
                assert ok > 0.999, f"e2e (ring) results did not come back: {ok}"
                variants["ring"] = {"Mops/s": round(steps * BATCH / t_e / 1e6, 1), "ms": round(t_e * 1e3, 3), "wall_ms": round(t_e * 1e3, 2),
                                    "timing": "host wall clock"}
                ring_info = {"rings": 8, "slots": 4, "ctas_per_ring": L.gpuhash_ring_ctas_per_ring(q),
                             "round_trip_us_one_64K_search_batch": round(rtt, 1)}
                log(f"e2e ring: {steps} steps in {t_e * 1e3:.2f} ms -> {steps * BATCH / t_e / 1e6:.1f} Mops/s; lone batch round trip {rtt:.1f} us")
            finally:
                L.gpuhash_ring_destroy(q)
            # the same lone batch through the launch path: submit + sync, zero-copy
            L.gpuhash_index_set_zero_copy(ix, 1)
            lat = []
            for i in range(64):
                a = time.perf_counter()
                N.check(L.gpuhash_index_submit(ix, 0, hs + 8 * N_SEARCH * (i % ke), N_SEARCH, ho + 8 * N_SEARCH * (i % ke), None, 0, None, 0))
                N.check(L.gpuhash_index_sync(ix))
                lat.append((time.perf_counter() - a) * 1e6)
            L.gpuhash_index_set_zero_copy(ix, 0)
            ring_info["round_trip_us_launch_path"] = round(float(np.median(lat)), 1)
    best = max(variants, key=lambda k: variants[k]["Mops/s"])
    e2e_val, wall_e = variants[best]["Mops/s"], variants[best]["wall_ms"] / 1e3

    # ---- CPU baseline: the oracle on one core, bounded sample of the same steps
    cpu = None
    if not args.no_cpu:
        try:
            cmem = host_mem_p(mem_p)
            cval, cdt, _ = cpu_workload(cmem, min(26, cmem - 7), 2500, 1, log, warm=20)
            cpu = {"value": round(cval, 3), "unit": "Mops/s", "cores": 1, "kind": "port",
                   "sample": f"2500 steps of the same 65536-request batch on a 2^{cmem} B host table preloaded with "
                             f"2^{min(26, cmem - 7)} keys, oracle/gpuhash_oracle.c, {cdt:.1f} s"}
        except Exception as e:                                        # never let the checker's environment kill the GPU line
            cpu = {"value": None, "unit": "Mops/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}

    if cpu is not None and ref_gpu is not None:
        cpu["reference_gpu_search_kernel"] = ref_gpu                 # same baseline leg: the reference's kernel on this GPU
    line = {
        "metric": "batched search/insert Mops/s (95/5 GET/SET, uniform keys)",
        "value": round(value, 1), "unit": "Mops/s", "n_gpus": 1, "steps": steps, "warmup": warm,
        "ms_per_step": round(t_val / steps * 1e3, 6), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(mem_p, args),
        "e2e": {"value": round(e2e_val, 1), "unit": "Mops/s", "h2d_bytes_per_step": 8 * N_SEARCH + 12 * N_INSERT,
                "d2h_bytes_per_step": 8 * N_SEARCH, "wall_ms": round(wall_e * 1e3, 2), "workers": min(S, 32),
                "path": best, "variants": variants, "ring": ring_info, "compact_results": compact_info},
        "gpu_launches": (1 if fused_resident else 2) * steps,
        "roofline": roof,
        "ops": ops,
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
        "search_hit_fraction": round(hit_frac, 5),
    }
    emit(line)
    L.gpuhash_host_free(hs); L.gpuhash_host_free(ho); L.gpuhash_host_free(hi)
    L.gpuhash_index_destroy(ix)
    return 0


if __name__ == "__main__":
    sys.exit(main())
