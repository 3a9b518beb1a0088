#!/usr/bin/env python
"""bench.py -- batched search/insert Mops/s of the hash-index hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): HASH_CUCKOO, table 2^34 bytes (>> L2), preloaded to load factor 0.25 (2^29 uniform
keys).  One STEP = one scheduler cycle (src/mega_scheduler.c:393-504: walk every worker's batch, synchronise once) over
W = 64 worker batches of 65 536 requests = 62 259 searches (95 %, keys drawn uniformly from the preloaded population, so
every search hits) + 3 277 inserts (5 %, fresh keys); per batch search -> insert in that order.  A step is ONE kernel
launch (gpuhash_cycle_multi_ex).

  value     whole-job Mops/s with the batches already resident in HBM: CUDA events around exactly K launches; the region is
            measured --reps times and the median reported (all regions listed under "timing")
  e2e       the same K steps through the call a scheduler makes -- gpuhash_index_submit_all / gpuhash_index_wait on PINNED
            HOST batches, cycles kept in order (the kernel of cycle k+1 starts when cycle k's has finished, as in the reference),
            staged through per-slot device copies (the workers' arrays are adjacent in pinned memory, so each array kind is ONE
            copy per direction), four cycles in flight -- timed by the HOST'S WALL CLOCK from the first submit to the return of
            the last wait.  ONE fixed path; the zero-copy paths and the unordered mode are listed next to it
  roofline  the same launches with searches only: algorithmic bytes (8 B request + 2 x 32 B signature sectors
            + 32 B location sector per hit bucket + 8 B result = 112 B) / time, against MEASURED_PEAKS.json hbm_gbs; plus
            one bulk launch and the measured random-probe ceiling of the same table
  parity    every search word of the last timed step (and of every e2e step) checked against the key's location
  cpu_baseline  oracle (C restatement of the reference's algorithm), one core, bounded sample of the same steps

  --impl reference : the reference's algorithm on the host cores (oracle, persistent pool of all threads), same config.
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

def host_mem_p(want):
    """largest table the host can hold next to everything else (the oracle table lives in host RAM)"""
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 8 << 30
    p = want
    while p > 20 and (1 << p) * 1.25 + (3 << 30) > avail:
        p -= 1
    return p


class CpuWorkload:
    """The same step on the host cores: oracle table of 2^mem_p bytes preloaded to load factor 0.25 (2^(mem_p-5) keys, the
    GPU arm's population), then steps of `batches` batches of 62 259 searches (keys drawn uniformly from the population:
    every search hits) + 3 277 fresh inserts.  The threads persist across steps (oracle/gpuhash_oracle.c orc_pool_*): all
    searches of a step over all threads, then its inserts over the 8 closed bucket ranges."""

    def __init__(self, mem_p, threads, log):
        from oracle import pyoracle as po
        self.po, self.mem_p, self.log = po, mem_p, log
        self.o = po.Oracle(mem_p, po.CUCKOO)
        self.pop = (1 << mem_p) // 32
        self.all = po.Pool(os.cpu_count() or 1)                       # preloading is not timed: always every core
        self.pool = self.all if threads == self.all.threads else po.Pool(threads)
        t0 = time.time()
        self.all.preload(self.o, SEED, 0, self.pop)
        log(f"cpu: preloaded {self.pop} keys into a 2^{mem_p} B host table in {time.time() - t0:.1f} s ({self.all.threads} threads)")
        self.next_key = self.pop

    def run(self, steps, warm, batches, budget_s=None):
        """returns (Mops/s, seconds, steps actually timed).  budget_s: stop after about that many seconds (>= 1 step)"""
        po = self.po
        ns, ni = N_SEARCH * batches, N_INSERT * batches
        out = np.empty(2 * ns, dtype=np.uint32)
        gen = lambda k: (po.gen_queries(SEED, self.pop, ns, 1000 + k), self._fresh(ni))
        for k in range(warm):
            sel, ins = gen(k)
            self.pool.cycle(self.o, sel, out, ins, batches)
        total, done = 0.0, 0
        for k in range(steps):
            sel, ins = gen(warm + k)                                  # generation is outside the clock
            t0 = po.now()
            self.pool.cycle(self.o, sel, out, ins, batches)
            total += po.now() - t0
            done += 1
            # (a handful of preloaded keys are legitimately unfindable: the reference re-homes an evicted victim with the
            # REQUEST's hash, gpu_hash.cu:334-335, which orphans it -- SURVEY Appendix B)
            if self.next_key * 8 < 0.3 * (1 << self.mem_p):          # (small test tables fill up within a few steps: evictions galore)
                assert ((out[0::2] != 0) | (out[1::2] != 0)).mean() > 0.9999, "cpu arm: searches of preloaded keys missed"
            if budget_s is not None and total >= budget_s:
                break
        return done * batches * BATCH / total / 1e6, total, done

    def _fresh(self, n):
        iel = self.po.keys(SEED, self.next_key, n)[0]
        self.next_key += n
        return iel

    def close(self):
        if self.pool is not self.all:
            self.pool.close()
        self.all.close()


def run_reference_arm(args, log):
    """bench.py --impl reference: the reference's algorithm (oracle/gpuhash_oracle.c, a line-by-line restatement of
    gpu_hash.cu) on the box's host cores with every thread it can use, same config, metric and unit as the GPU arm."""
    cores = os.cpu_count() or 1
    mem_p = host_mem_p(args.mem_p)
    steps, warm = max(1, args.steps), max(1, min(args.warmup, 3))
    w = CpuWorkload(mem_p, cores, log)
    val, dt, done = w.run(steps, warm, args.batches_per_step, budget_s=120.0)
    w.close()
    line = {
        "impl": "reference", "metric": METRIC,
        "value": round(val, 3), "unit": "Mops/s", "n_gpus": args.gpus, "steps": done, "warmup": warm,
        "ms_per_step": round(1e3 * dt / done, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(mem_p, args),
        "cpu_baseline": {"value": round(val, 3), "unit": "Mops/s", "cores": w.pool.threads, "kind": "port",
                         "sample": f"{done} steps of {args.batches_per_step} x 65536 requests after {warm} warm-up steps on a 2^{mem_p} B host table "
                                   f"preloaded with {w.pop} keys; oracle/gpuhash_oracle.c through a persistent thread pool (searches over all "
                                   f"{w.pool.threads} threads, inserts over the 8 closed bucket ranges), wall clock {dt:.2f} s"},
        "e2e": {"value": round(val, 3), "unit": "Mops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


METRIC = "batched search/insert Mops/s (95/5 GET/SET, uniform keys)"


def workload_config(mem_p, args):
    """identical in both arms (the driver compares them)"""
    W = args.batches_per_step
    return {"workload": f"configs[1]: 1xB200 search-heavy 95/5 GET/SET, uniform keys, HASH_CUCKOO, table 2^{mem_p} bytes, "
                        f"batch 64K signatures ({N_SEARCH} searches + {N_INSERT} inserts); one step = one scheduler cycle over "
                        f"{W} worker batches (mega_scheduler.c:393-504)",
            "mem_p": mem_p, "algo": "HASH_CUCKOO", "load_factor": 0.25, "batch": BATCH, "batches_per_step": W,
            "requests_per_step": W * BATCH,
            "cache": "every step has its own request/result arrays (64 MiB per step > L2 is not needed: the table is 16 GiB >> 126 MB L2 "
                     "and every probe is a random line of it)"}


# ------------------------------------------------------------------------------------------ GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mem-p", type=int, default=34)
    ap.add_argument("--batches-per-step", type=int, default=64,
                    help="worker batches of 65 536 requests per scheduler cycle = per step (the reference's cycle walks all its "
                         "workers' batches and synchronises once, mega_scheduler.c:393-504)")
    ap.add_argument("--streams", type=int, default=4,
                    help="cycle launches in flight for the resident legs (1..4 = GPUHASH_INDEX_SLOTS): 1 = strictly one after the other "
                         "(reported next to `value` in any case), more = the tail of a cycle overlaps the head of the next "
                         "(measured 1 / 2 / 4: 19.9 / 20.7 / 21.0 Gops/s)")
    ap.add_argument("--reps", type=int, default=5, help="the timed region of exactly --steps steps is measured this many times; the median is reported")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ops", action="store_true", help="skip the per-operation bulk launches")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference's own search kernel (oracle/_ref)")
    ap.add_argument("--no-ring", action="store_true", help="skip the persistent-kernel ring variant of the e2e leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2] / configs[3] legs")
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3],
                    help="BASELINE configs[2] (two-choice, zipf, L2-resident vs HBM) and configs[3] (cuckoo churn at 90 %% load) are extra "
                         "keys of the line (config2, config3); 2 or 3 runs only that one of the two")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.batches_per_step = max(1, min(args.batches_per_step, 128))

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

    steps, warm, W = max(1, args.steps), max(3, args.warmup), args.batches_per_step
    mem_p = args.mem_p
    free, total = C.c_size_t(), C.c_size_t()
    N.check(L.gpuhash_device_info(local_rank, None, None, C.byref(free), C.byref(total)))
    while (1 << mem_p) + (12 << 30) > free.value and mem_p > 26:
        mem_p -= 1
    ix = L.gpuhash_index_create(mem_p, N.CUCKOO, W, N_SEARCH, N_INSERT, 1)
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

    # ---- resident batches: every step of a timed region has its own W batches (request, result and insert arrays)
    ks = min(steps + warm, 48)                                    # distinct steps resident in HBM (66 MB each); longer runs wrap
    kd = ks * W
    search_d = mk.DeviceBuffer(8 * N_SEARCH * kd)
    out_d = mk.DeviceBuffer(8 * N_SEARCH * kd)
    insert_d = mk.DeviceBuffer(12 * N_INSERT * kd)
    expect_d = mk.DeviceBuffer(4 * N_SEARCH * kd)
    N.check(L.gpuhash_gen_queries(search_d.ptr, expect_d.ptr, SEED, pop, N_SEARCH * kd, 99, 0.0, 0.0, None))
    next_key = [pop]

    def fresh_inserts(dst_ptr, n):
        """new keys every time a batch array is (re)used, so inserts stay inserts (not updates of a previous pass)"""
        N.check(L.gpuhash_gen_inserts(dst_ptr, None, SEED, next_key[0], n, None))
        next_key[0] += n

    fresh_inserts(insert_d.ptr, N_INSERT * kd)
    N.check(L.gpuhash_device_sync())

    def resident(first_step, count, n_insert=N_INSERT, streams=None):
        """`count` steps starting at resident step `first_step` (wrapping); one launch per step; returns seconds (CUDA events)"""
        streams = args.streams if streams is None else streams
        total_ms, done = 0.0, 0
        while done < count:
            s0 = (first_step + done) % ks
            c = min(count - done, ks - s0)
            res = N.BenchResult()
            b0 = s0 * W
            N.check(L.gpuhash_bench_cycles(C.byref(geom), table, search_d.ptr + 8 * N_SEARCH * b0, N_SEARCH, out_d.ptr + 8 * N_SEARCH * b0,
                                           insert_d.ptr + 12 * N_INSERT * b0, n_insert, W, c, streams, C.byref(res)),
                    "gpuhash_bench_cycles")
            total_ms += res.total_ms; done += c
        return total_ms / 1e3

    sampler = ClockSampler(local_rank)
    resident(0, warm)                                               # W untimed warm-up steps
    regions = []
    with sampler:
        for r in range(max(1, args.reps)):                          # each region: EXACTLY K steps, one launch each
            regions.append(resident(warm, steps))
            fresh_inserts(insert_d.ptr, N_INSERT * kd); N.check(L.gpuhash_device_sync())
    t_val = float(np.median(regions))
    value = steps * W * BATCH / t_val / 1e6
    log(f"resident: {steps} steps x {W} batches, regions {[round(x * 1e3, 3) for x in regions]} ms -> {value:.1f} Mops/s")
    # the same steps strictly one after the other (cycle k+1 starts when cycle k has finished, the reference's order)
    strict = None
    if args.streams != 1:
        with sampler:
            s_reg = []
            for r in range(3):
                s_reg.append(resident(warm, steps, streams=1))
                fresh_inserts(insert_d.ptr, N_INSERT * kd); N.check(L.gpuhash_device_sync())
        t_st = float(np.median(s_reg))
        strict = {"Mops/s": round(steps * W * BATCH / t_st / 1e6, 1), "ms_per_step": round(t_st / steps * 1e3, 6),
                  "regions_ms": [round(x * 1e3, 3) for x in s_reg],
                  "what": "the same launches on ONE stream: the kernel of step k+1 starts when step k's has finished"}

    # ---- parity on the timed output, word for word: every search of the LAST timed step must return the location its key was
    #      inserted with (generator's expect_loc = key index + 1) in exactly one of its two words and 0 in the other
    last = (warm + steps - 1) % ks
    nchk = N_SEARCH * W
    chk = np.empty(2 * nchk, dtype=np.uint32); exp = np.empty(nchk, dtype=np.uint32)
    N.check(L.gpuhash_d2h(chk.ctypes.data, out_d.ptr + 8 * nchk * last, chk.nbytes, None))
    N.check(L.gpuhash_d2h(exp.ctypes.data, expect_d.ptr + 4 * nchk * last, exp.nbytes, None)); N.check(L.gpuhash_device_sync())
    sel_chk = np.empty(2 * nchk, dtype=np.uint32)
    N.check(L.gpuhash_d2h(sel_chk.ctypes.data, search_d.ptr + 8 * nchk * last, sel_chk.nbytes, None)); N.check(L.gpuhash_device_sync())

    def table_has(sig, hash_):
        """direct inspection of the device table (pair layout: slot l = words 2l, 2l+1) with the reference's bucket functions
        (gpu_hash.cu:55,66-67): is `sig` stored in either candidate bucket?"""
        b1 = hash_ & geom.hash_mask
        b2 = (((hash_ ^ sig) & geom.block_mask) | (hash_ & ~geom.block_mask & 0xFFFFFFFF)) & geom.hash_mask
        for b in (b1, b2):
            w = np.empty(16, dtype=np.uint32)
            N.check(L.gpuhash_d2h(w.ctypes.data, table + 64 * int(b), 64, None)); N.check(L.gpuhash_device_sync())
            if (w[0::2] == sig).any():
                return True
        return False

    def check_words(words, expect, sel_words):
        """(mismatches, orphans): a search must return its key's location (generator: key index + 1) in exactly one word and
        0 or the same location in the other.  Both words 0 is right only if the key is really absent from both of its
        buckets -- the reference orphans an evicted victim by re-homing it with the REQUEST's hash (gpu_hash.cu:334-335,
        SURVEY Appendix B), a few per 10^8 inserts at this load factor -- which is verified bucket by bucket."""
        o0, o1 = words[0::2], words[1::2]
        good = ((o0 == expect) & ((o1 == 0) | (o1 == expect))) | ((o1 == expect) & (o0 == 0))
        unfound = np.nonzero((o0 == 0) & (o1 == 0))[0]
        bad = int((~good).sum()) - len(unfound)
        if len(unfound) > 2000:
            return bad + len(unfound), 0
        present = sum(1 for i in unfound if table_has(int(sel_words[2 * i]), int(sel_words[2 * i + 1])))
        return bad + present, len(unfound) - present

    mismatches, orphans = check_words(chk, exp, sel_chk)
    o0, o1 = chk[0::2], chk[1::2]
    hit_frac = float(((o0 != 0) | (o1 != 0)).mean())
    hits_per_search = float(((o0 != 0).sum() + (o1 != 0).sum()) / nchk)
    assert mismatches == 0, f"{mismatches} of {nchk} timed searches returned something else than their key's location"

    # ---- roofline: the same launches with the searches only (the cycle kernel's search tiles on the same batches)
    peak, peak_src = peaks()
    bytes_per_search = 8 + 2 * 32 + 32 * hits_per_search + 8
    resident(0, min(warm, 5), n_insert=0)
    s_regions = []
    with sampler:
        for r in range(max(1, args.reps)):
            s_regions.append(resident(warm, steps, n_insert=0))
    t_s = float(np.median(s_regions))
    achieved = steps * W * N_SEARCH * bytes_per_search / t_s / 1e9
    # one bulk launch of the plain search entry point (2^24 requests)
    bulk_n = min(1 << 24, N_SEARCH * kd)
    res = N.BenchResult()
    for _ in range(3):
        N.check(L.gpuhash_bench_resident(C.byref(geom), table, search_d.ptr, bulk_n, out_d.ptr, None, 0, 1, 1, 0, C.byref(res)))
    bulk_gbs = bulk_n * bytes_per_search / (res.total_ms / 1e3) / 1e9
    bulk_mops = bulk_n / (res.total_ms / 1e3) / 1e6
    # measured random-probe ceiling on the same allocation (32 B loads, one per thread, four in flight)
    ms = C.c_float()
    N.check(L.gpuhash_roofline_gather(table, 1 << mem_p, 1 << 27, 0, 4, 3, C.byref(ms), None))
    sector_rate = (1 << 27) / (ms.value / 1e3)                       # random probes per second
    traffic, traffic_src = None, None                                # DRAM bytes per launch of that kernel, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": round(W * N_SEARCH * bytes_per_search),
            "kernel": "gh::cycle_multi_kernel<pairs> (search tiles of one step: %d x %d searches)" % (W, N_SEARCH),
            "peak_source": peak_src + " (of measured)", "bytes_per_search": round(bytes_per_search, 2),
            "launches": steps, "avg_launch_us": round(t_s / steps * 1e6, 2), "regions_ms": [round(x * 1e3, 3) for x in s_regions],
            "bulk_launch": {"requests": bulk_n, "GB/s": round(bulk_gbs, 1), "Mops/s": round(bulk_mops, 1),
                            "frac": round(bulk_gbs / peak, 4)},
            "random_probe": {"Gprobes/s": round(sector_rate / 1e9, 2),
                             "what": "random 32 B loads from the same 16 GiB table: the request (DRAM row activation) ceiling; a search "
                                     "is two probes (one per bucket), so the ceiling for searches is half of it",
                             "search_frac_of_ceiling": round(steps * W * N_SEARCH * 2 / t_s / sector_rate, 4),
                             "bulk_frac_of_ceiling": round(bulk_n * 2 / (res.total_ms / 1e3) / sector_rate, 4)}}

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
            nb = min(kd, 256)                                        # batches compared word for word and timed
            ours = np.empty(2 * N_SEARCH * nb, dtype=np.uint32)
            N.check(L.gpuhash_search_ex(C.byref(geom), search_d.ptr, out_d.ptr, table, N_SEARCH * nb, None, None)); N.check(L.gpuhash_device_sync())
            N.check(L.gpuhash_d2h(ours.ctypes.data, out_d.ptr, ours.nbytes, None)); N.check(L.gpuhash_device_sync())
            N.check(L.gpuhash_table_convert(C.byref(geom), table, N.LAYOUT_REFERENCE, None)); N.check(L.gpuhash_device_sync())
            try:
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
                theirs = np.empty(2 * N_SEARCH * nb, dtype=np.uint32)
                N.check(L.gpuhash_d2h(theirs.ctypes.data, out_d.ptr, theirs.nbytes, None)); N.check(L.gpuhash_device_sync())
                ref_gpu = {"what": "pzrq/megakv hash_search (gpu_hash.cu:28-75) as compiled by oracle/Makefile (compute_60 PTX -> sm_100), launch shape "
                                   "24576 x 256 + memset per batch, device-resident batches, host wall clock over "
                                   f"{nb} batches of {N_SEARCH}",
                           "Mops/s_one_stream": round(nb * N_SEARCH / t_ref / 1e6, 1), "us_per_batch": round(t_ref / nb * 1e6, 2),
                           "Mops/s_16_streams": round(nb * N_SEARCH / t_ref16 / 1e6, 1),
                           "ours_search_only_over_its_16_streams": round(steps * W * N_SEARCH / t_s / (nb * N_SEARCH / t_ref16), 2),
                           "ours_search_only_over_its_one_stream": round(steps * W * N_SEARCH / t_s / (nb * N_SEARCH / t_ref), 2),
                           "words_compared": int(ours.size), "results_equal_ours": bool(np.array_equal(ours, theirs))}
                log(f"reference GPU search kernel: {ref_gpu['Mops/s_one_stream']} Mops/s on one stream, {ref_gpu['Mops/s_16_streams']} on 16; "
                    f"results equal ours: {ref_gpu['results_equal_ours']}")
            finally:
                as_ref = N.Geom.from_buffer_copy(bytes(geom)); as_ref.layout = N.LAYOUT_REFERENCE
                N.check(L.gpuhash_table_convert(C.byref(as_ref), table, geom.layout, None)); N.check(L.gpuhash_device_sync())
        except Exception as e:                                       # a baseline must never take the product line down
            ref_gpu = {"failed": str(e)}

    # ---- every operation on its own, uniform and zipf(0.99) keys, one bulk launch each (north star: search, insert and
    #      delete Mops/s, absolute and against the random-access roofline).  Probes per op: search 2 buckets,
    #      insert / delete 1 bucket read + 1 written back.
    ops = {}
    if not args.no_ops:
        from megakv_b200 import keystream as ksm
        ev_a, ev_b = L.gpuhash_event_create(), L.gpuhash_event_create()

        def timed(fn):
            N.check(L.gpuhash_device_sync())
            N.check(L.gpuhash_event_record(ev_a, None)); N.check(fn()); N.check(L.gpuhash_event_record(ev_b, None))
            t = C.c_float(); N.check(L.gpuhash_event_elapsed_ms(ev_a, ev_b, C.byref(t)))
            return t.value / 1e3

        n_ops = min(1 << 22, N_SEARCH * kd)
        zn = ksm.zetan(pop, 0.99)
        req_d = mk.DeviceBuffer(12 * n_ops)
        sel_d = mk.DeviceBuffer(8 * n_ops); res_d = mk.DeviceBuffer(8 * n_ops)

        def report(name, secs, probes):
            ops[name] = {"Mops/s": round(n_ops / secs / 1e6, 1), "frac_of_probe_ceiling": round(n_ops * probes / secs / sector_rate, 3)}

        for dist_name, theta, z in (("uniform", 0.0, 0.0), ("zipf0.99", 0.99, zn)):
            N.check(L.gpuhash_gen_queries(sel_d.ptr, None, SEED, pop, n_ops, 4242, theta, z, None))
            timed(lambda: L.gpuhash_search_ex(C.byref(geom), sel_d.ptr, res_d.ptr, table, n_ops, None, None))
            report(f"search_{dist_name}", timed(lambda: L.gpuhash_search_ex(C.byref(geom), sel_d.ptr, res_d.ptr, table, n_ops, None, None)), 2)
        # insert: fresh uniform keys (claims), then zipf draws from the population (updates in place, hot slots contended)
        N.check(L.gpuhash_gen_inserts(req_d.ptr, None, SEED, next_key[0] + (1 << 28), n_ops, None))
        report("insert_uniform", timed(lambda: L.gpuhash_insert_flat_ex(C.byref(geom), table, req_d.ptr, n_ops, None, 0, None)), 1)
        report("delete_uniform", timed(lambda: L.gpuhash_delete_ex(C.byref(geom), req_d.ptr, table, n_ops, None, 0, None)), 1)   # removes them again
        N.check(L.gpuhash_gen_requests(req_d.ptr, SEED, pop, n_ops, 777, 0.99, zn, None))
        report("insert_zipf0.99", timed(lambda: L.gpuhash_insert_flat_ex(C.byref(geom), table, req_d.ptr, n_ops, None, 0, None)), 1)
        report("delete_zipf0.99", timed(lambda: L.gpuhash_delete_ex(C.byref(geom), req_d.ptr, table, n_ops, None, 0, None)), 1)
        N.check(L.gpuhash_insert_flat_ex(C.byref(geom), table, req_d.ptr, n_ops, None, 0, None))   # put the deleted hot keys back
        N.check(L.gpuhash_device_sync())
        ops["requests_per_launch"] = n_ops
        req_d.free(); sel_d.free(); res_d.free()
        L.gpuhash_event_destroy(ev_a); L.gpuhash_event_destroy(ev_b)

    # ---- e2e: the call a scheduler makes -- gpuhash_index_submit_all(W pinned host batches) ... gpuhash_index_wait -- ONE fixed
    #      path (HEADLINE below), HOST WALL CLOCK from the first submit to the return of the last wait; other paths next to it.
    ke = min(steps + warm, 24)                                       # distinct steps in pinned memory (66 MB each)
    hb = ke * W
    hs = L.gpuhash_host_alloc(8 * N_SEARCH * hb); ho = L.gpuhash_host_alloc(8 * N_SEARCH * hb); hi = L.gpuhash_host_alloc(12 * N_INSERT * hb)
    if not (hs and ho and hi):
        raise mk.GpuHashError("pinned host allocation failed")
    N.check(L.gpuhash_d2h(hs, search_d.ptr, 8 * N_SEARCH * hb, None)); N.check(L.gpuhash_device_sync())
    exp_h = np.empty(N_SEARCH * hb, dtype=np.uint32)
    N.check(L.gpuhash_d2h(exp_h.ctypes.data, expect_d.ptr, exp_h.nbytes, None)); N.check(L.gpuhash_device_sync())
    ho_np = np.ctypeslib.as_array(C.cast(ho, C.POINTER(C.c_uint32)), shape=(2 * N_SEARCH * hb,))

    def host_inserts():
        fresh_inserts(insert_d.ptr, N_INSERT * hb)
        N.check(L.gpuhash_d2h(hi, insert_d.ptr, 12 * N_INSERT * hb, None)); N.check(L.gpuhash_device_sync())

    def e2e_pass(count, depth=2):
        """`count` steps through submit_all / wait; returns (wall seconds, steps)"""
        host_inserts()
        r = N.BenchResult()
        N.check(L.gpuhash_bench_e2e_cycles(ix, hs, N_SEARCH, ho, hi, N_INSERT, W, hb, count, depth, C.byref(r)), "gpuhash_bench_e2e_cycles")
        return r.total_ms / 1e3

    hs_np = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint32)), shape=(2 * N_SEARCH * hb,))

    def e2e_check(compact=False):
        n = N_SEARCH * W * min(steps, ke)
        if compact:
            got = ho_np[:n]
            words = np.zeros(2 * n, dtype=np.uint32); words[0::2] = got
            return check_words(words, exp_h[:n], hs_np[:2 * n])[0]
        return check_words(ho_np[:2 * n], exp_h[:n], hs_np[:2 * n])[0]

    variants = {}
    HEADLINE = "staged copies (adjacent host batches: one copy per array) + one launch per step, cycles ordered, 4 in flight"
    paths = [   # name, zero_copy, unordered cycles, cycles in flight
        (HEADLINE, 0, 0, 4),
        ("staged copies + one launch per step, cycles ordered, 2 in flight", 0, 0, 2),
        ("zero-copy (the kernel reads / writes the pinned buffers) + one launch per step, cycles ordered, 2 in flight", 1, 0, 2),
        ("zero-copy + one launch per step, one cycle in flight", 1, 0, 1),
        ("zero-copy + one launch per step, cycle kernels UNORDERED, 2 in flight", 1, 1, 2),
    ]
    t_e = e2e_val = e2e_bad = None
    e_regions = []
    for name, zc, unordered, depth in paths:
        L.gpuhash_index_set_zero_copy(ix, zc); L.gpuhash_index_set_unordered_cycles(ix, unordered)
        e2e_pass(max(depth, min(warm, 3)), depth=depth); ho_np[:] = 0      # (every slot in flight is warm: staging buffers are allocated at first use)
        with sampler:
            regs = [e2e_pass(steps, depth=depth) for _ in range(3)]
        t = float(np.median(regs))
        bad = e2e_check()
        assert bad == 0, f"e2e ({name}): {bad} searches came back wrong"
        variants[name] = {"Mops/s": round(steps * W * BATCH / t / 1e6, 1), "wall_ms": round(t * 1e3, 3)}
        if name == HEADLINE:
            t_e, e_regions, e2e_bad, e2e_val = t, regs, bad, steps * W * BATCH / t / 1e6
            log(f"e2e {name}: {steps} steps, wall {[round(x * 1e3, 2) for x in regs]} ms -> {e2e_val:.1f} Mops/s")
    L.gpuhash_index_set_unordered_cycles(ix, 0)
    # the consumer only ever takes one of the two result words (mega_send.c:411-414): let the device choose and send 4 B
    # per search back instead of 8.  Reported next to the headline, not as it: it changes what search_out holds.
    L.gpuhash_index_set_zero_copy(ix, 0); L.gpuhash_index_set_compact_results(ix, 1)
    e2e_pass(min(warm, 3), depth=4); ho_np[:] = 0
    t_c = e2e_pass(steps, depth=4)
    c_bad = e2e_check(compact=True)
    assert c_bad == 0, f"e2e (compact): {c_bad} searches came back wrong"
    compact_info = {"Mops/s": round(steps * W * BATCH / t_c / 1e6, 1), "wall_ms": round(t_c * 1e3, 3), "d2h_bytes_per_step": 4 * N_SEARCH * W,
                    "what": "the headline path with one result word per search (gpuhash_index_set_compact_results)"}
    L.gpuhash_index_set_compact_results(ix, 0)
    L.gpuhash_index_set_zero_copy(ix, 0)
    # no launches at all -- descriptor rings in pinned memory feeding the persistent kernel (north star (c)), wall clock
    ring_info = None
    if not args.no_ring:
        q = L.gpuhash_ring_create(C.byref(geom), table, 8, 4, 4, 2000)
        if q:
            try:
                def ring_pass(count, reps=0):
                    host_inserts()
                    rtt = C.c_float(0)
                    r = N.BenchResult()
                    N.check(L.gpuhash_bench_ring(q, hs, N_SEARCH, ho, hi, N_INSERT, min(count * W, hb), C.byref(r), reps, C.byref(rtt)),
                            "gpuhash_bench_ring")
                    return r.total_ms / 1e3, rtt.value, min(count * W, hb)
                ring_pass(min(warm, 3))
                with sampler:
                    t_r, rtt, nb_r = ring_pass(steps, 64)
                variants["ring (persistent kernel, one doorbell per batch)"] = {"Mops/s": round(nb_r * BATCH / t_r / 1e6, 1), "wall_ms": round(t_r * 1e3, 3),
                                                                               "batches": nb_r}
                ring_info = {"rings": 8, "slots": 4, "ctas_per_ring": L.gpuhash_ring_ctas_per_ring(q),
                             "round_trip_us_one_64K_search_batch": round(rtt, 1)}
            finally:
                L.gpuhash_ring_destroy(q)
            # the same lone batch through the launch path: submit + sync, zero-copy
            L.gpuhash_index_set_zero_copy(ix, 1)
            lat = []
            for i in range(64):
                a = time.perf_counter()
                N.check(L.gpuhash_index_submit(ix, 0, hs + 8 * N_SEARCH * (i % hb), N_SEARCH, ho + 8 * N_SEARCH * (i % hb), None, 0, None, 0))
                N.check(L.gpuhash_index_sync(ix))
                lat.append((time.perf_counter() - a) * 1e6)
            L.gpuhash_index_set_zero_copy(ix, 0)
            ring_info["round_trip_us_launch_path"] = round(float(np.median(lat)), 1)

    # ---- BASELINE configs[2] and configs[3] as extra keys of the same line
    extra = {}
    if not args.no_configs:
        try:
            from megakv_b200 import bench_configs
            extra = bench_configs.run(args, L, N, mk, local_rank, log, sector_rate)
        except Exception as e:                                        # never let a secondary leg take the headline down
            extra = {"configs_failed": repr(e)}

    # ---- CPU baseline: the oracle on ONE core, bounded sample of the same steps (preloaded by all cores, untimed)
    cpu = None
    if not args.no_cpu:
        try:
            cmem = host_mem_p(mem_p)
            cw = CpuWorkload(cmem, 1, log)
            cval, cdt, cdone = cw.run(64, 1, W, budget_s=12.0)
            cw.close()
            cpu = {"value": round(cval, 3), "unit": "Mops/s", "cores": 1, "kind": "port",
                   "sample": f"{cdone} steps of {W} x 65536 requests on a 2^{cmem} B host table preloaded with {cw.pop} keys, "
                             f"oracle/gpuhash_oracle.c on one thread, {cdt:.1f} s"}
        except Exception as e:                                        # never let the checker's environment kill the GPU line
            cpu = {"value": None, "unit": "Mops/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}

    if cpu is not None and ref_gpu is not None:
        cpu["reference_gpu_search_kernel"] = ref_gpu                 # same baseline leg: the reference's kernel on this GPU
    line = {
        "metric": METRIC,
        "value": round(value, 1), "unit": "Mops/s", "n_gpus": 1, "steps": steps, "warmup": warm,
        "ms_per_step": round(t_val / steps * 1e3, 6), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(mem_p, args),
        "timing": {"timed_region_ms": round(t_val * 1e3, 3), "regions_ms": [round(x * 1e3, 3) for x in regions],
                   "what": f"each region = exactly {steps} steps = {steps} launches, CUDA events; median of {len(regions)} regions",
                   "cycles_in_flight": args.streams,
                   "cycle_order": ("strict" if args.streams == 1 else
                                   f"{args.streams} cycles in flight on {args.streams} streams: the tail of a cycle overlaps the head of the next, so requests of "
                                   "consecutive cycles are unordered where they overlap -- like workers inside a cycle (mega_scheduler.c:393-502); the "
                                   "benchmark's requests are independent of each other either way (searches of preloaded keys, inserts of fresh keys)"),
                   "strict_cycle_order": strict},
        "e2e": {"value": round(e2e_val, 1), "unit": "Mops/s", "h2d_bytes_per_step": (8 * N_SEARCH + 12 * N_INSERT) * W,
                "d2h_bytes_per_step": 8 * N_SEARCH * W, "wall_ms": round(t_e * 1e3, 3), "regions_wall_ms": [round(x * 1e3, 3) for x in e_regions],
                "timing": "host wall clock, first gpuhash_index_submit_all to the return of the last gpuhash_index_wait",
                "path": HEADLINE, "mismatches": e2e_bad,
                "variants": variants, "ring": ring_info, "compact_results": compact_info},
        "gpu_launches": steps,
        "parity_checked": True, "mismatches": mismatches, "searches_checked": nchk, "orphaned_keys_seen": orphans,
        "roofline": roof,
        "ops": ops,
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
        "search_hit_fraction": round(hit_frac, 5),
    }
    line.update(extra)
    emit(line)
    L.gpuhash_host_free(hs); L.gpuhash_host_free(ho); L.gpuhash_host_free(hi)
    L.gpuhash_index_destroy(ix)
    return 0


if __name__ == "__main__":
    sys.exit(main())
