"""N > 1 arm of bench.py: the sharded index (BASELINE.json configs[4]).

Weak scaling: every rank owns one 2^mem_p-byte shard (logical table 2^(mem_p + log2 N) bytes, preloaded to load
factor 0.25 through the routed insert path itself) and issues its own batch of 65 536 requests per step
(62 259 searches for keys drawn uniformly from the WHOLE population -> (N-1)/N of them leave the GPU, + 3 277
inserts of fresh keys).  Both exchanges are inside the timed region.  The fused path (peer stores over NVLink +
flags, no host sync, the K steps replayed as one CUDA graph) is the product; the same steps through NCCL
all_to_all_single are timed next to it as `nccl_baseline`.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

BATCH, N_SEARCH = 65536, 62259
N_INSERT = BATCH - N_SEARCH
SEED = 1


def main(args, rank, world, local_rank, log):
    import torch
    import torch.distributed as dist
    import megakv_b200 as mk
    from megakv_b200 import _native as N
    from megakv_b200.sharded import ShardPlan, ShardedIndex, CudaShardBackend
    import bench as B

    torch.cuda.set_device(local_rank)
    N.check(mk.lib().gpuhash_set_device(local_rank))
    if "RANK" not in os.environ:                                  # single process (profiling the routed kernels on one GPU)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = mk.lib()
    dev = torch.device("cuda", local_rank)
    steps, warm = max(1, args.steps), max(3, args.warmup)
    mem_p_shard = args.mem_p
    free, total = C.c_size_t(), C.c_size_t()
    N.check(L.gpuhash_device_info(local_rank, None, None, C.byref(free), C.byref(total)))
    while (1 << mem_p_shard) + (12 << 30) > free.value and mem_p_shard > 26:
        mem_p_shard -= 1
    log2w = world.bit_length() - 1
    plan = ShardPlan(min(mem_p_shard + log2w, 38), world)
    cap = 1 << 20
    be = CudaShardBackend(plan, rank, cap)                        # bulk lane: preload in 1 M-request batches
    ix = ShardedIndex(be, plan, exchange="p2p")
    # S lanes = S batches in flight per GPU, each with its own inboxes/staging/flags and stream (the sharded
    # counterpart of the reference's one-stream-per-worker, mega_scheduler.c:276-280)
    # One exchange routes GROUP consecutive 64 K batches of this GPU at once -- what the reference's scheduler cycle does
    # with the batches of all its workers (mega_scheduler.c:392-502 loops over cpu_worker_num <= 16 buffers per cycle).
    # A graph node costs ~2 us of front-end time here and a routed batch needs ten of them, so per-batch exchanges are
    # node-bound (2 GPUs: 22 us per 64 K batch however many lanes); per-cycle exchanges are not.
    # 64 batches per exchange (2 GPUs, 8 lanes: 16 batches 25.6, 32: 31.1, 64: 33.6 Gops/s -- every kernel and flag wait
    # has a fixed cost); a short run (--steps below 128) is cut into two exchanges rather than into many small ones, and
    # takes only as many lanes as it has exchanges
    GROUP = int(os.environ.get('GPUHASH_GROUP', 0)) or min(64, max(1, (steps + 1) // 2))
    S = max(1, min(int(os.environ.get('GPUHASH_LANES', 8)), 16, -(-steps // GROUP)))
    quick = bool(os.environ.get('GPUHASH_BENCH_QUICK'))
    lanes = [ShardedIndex(CudaShardBackend(plan, rank, GROUP * BATCH, table=be.table), plan, exchange="p2p") for _ in range(S)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    # GPUHASH_SPLIT_UPDATES=1 (experiment): the update exchange of a cycle runs next to its search exchange -- own (small)
    # inboxes, flags and stream per lane; only its serve kernel waits (event) for this rank's search serve kernel of the
    # same cycle, which keeps the reference's in-stream order search -> insert at every owner for every origin.
    # Measured on 2 GPUs, 8 lanes: 33.0 vs 33.3 Gops/s without -- the other lanes fill those gaps already.  Off.
    split_updates = os.environ.get('GPUHASH_SPLIT_UPDATES', '0') != '0'
    ulanes = [ShardedIndex(CudaShardBackend(plan, rank, GROUP * N_INSERT, table=be.table), plan, exchange="p2p") for _ in range(S)] if split_updates else []
    ustreams = [torch.cuda.Stream(device=dev) for _ in range(S)] if split_updates else []
    ixc = ShardedIndex(lanes[0].be, plan, exchange="collective")  # same table and buffers, NCCL exchange (baseline)

    # ---- preload through the routed insert path: rank r inserts key indices r, r + world, ... in chunks
    pop = (1 << plan.mem_p_total) // 8 // 4
    per_rank = pop // world
    gen = torch.empty((cap, 3), dtype=torch.int32, device=dev)
    t0 = time.time()
    for first in range(0, per_rank, cap):
        n = min(cap, per_rank - first)
        N.check(L.gpuhash_gen_inserts(gen.data_ptr(), None, SEED, rank * per_rank + first, n, be._stream()))
        ix.insert(gen[:n])
    torch.cuda.synchronize(); dist.barrier()
    log(f"preloaded {pop} keys over {world} shards in {time.time() - t0:.2f} s (p2p err={be.p2p_error()})")

    # ---- resident batches (whole cycles of GROUP batches)
    kd = -(-min(steps + warm, 2048) // GROUP) * GROUP
    sel = torch.empty((kd, N_SEARCH, 2), dtype=torch.int32, device=dev)
    ins = torch.empty((kd, N_INSERT, 3), dtype=torch.int32, device=dev)
    out = torch.empty((kd, N_SEARCH, 2), dtype=torch.int32, device=dev)
    N.check(L.gpuhash_gen_queries(sel.data_ptr(), None, SEED, per_rank * world, N_SEARCH * kd, 99 + rank, 0.0, 0.0, be._stream()))
    N.check(L.gpuhash_gen_inserts(ins.data_ptr(), None, SEED, pop + rank * (1 << 26), N_INSERT * kd, be._stream()))
    torch.cuda.synchronize()

    sel_f, ins_f, out_f = sel.view(-1, 2), ins.view(-1, 3), out.view(-1, 2)

    def cycles_of(first, count):
        """(first batch, number of batches <= GROUP) for exactly `count` batches starting at `first`, never wrapping"""
        i = 0
        while i < count:
            b = (first + i) % kd
            g = min(GROUP, count - i, kd - b)
            yield b, g
            i += g

    def run_steps(index, first, count, with_insert=True):
        """exactly `count` batches, up to GROUP of them per exchange"""
        if index is not None:                                           # one lane, one stream (NCCL baseline)
            for b, g in cycles_of(first, count):
                index.search(sel_f[b * N_SEARCH:(b + g) * N_SEARCH], out_f[b * N_SEARCH:(b + g) * N_SEARCH])
                if with_insert:
                    index.insert(ins_f[b * N_INSERT:(b + g) * N_INSERT])
            return
        cur = torch.cuda.current_stream()
        for st in streams + ustreams:
            st.wait_stream(cur)
        for c, (b, g) in enumerate(cycles_of(first, count)):
            k = c % S
            sq, oq, iq = sel_f[b * N_SEARCH:(b + g) * N_SEARCH], out_f[b * N_SEARCH:(b + g) * N_SEARCH], ins_f[b * N_INSERT:(b + g) * N_INSERT]
            if with_insert and split_updates:
                ev = torch.cuda.Event()
                with torch.cuda.stream(streams[k]):
                    lanes[k].search(sq, oq, after_serve=lambda: ev.record(torch.cuda.current_stream()))
                with torch.cuda.stream(ustreams[k]):
                    ulanes[k].insert(iq, before_serve=lambda: torch.cuda.current_stream().wait_event(ev))
                continue
            with torch.cuda.stream(streams[k]):
                lanes[k].search(sq, oq)
                if with_insert:
                    lanes[k].insert(iq)
        for st in streams + ustreams:
            cur.wait_stream(st)

    def timed(index, first, count, graph, with_insert=True):
        """seconds for `count` steps, max over ranks"""
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run_steps(index, first, count, with_insert)
            torch.cuda.synchronize(); dist.barrier()
            e0.record(); g.replay(); e1.record()
        else:
            e0.record(); run_steps(index, first, count, with_insert); e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def phase_profile(count=40):
        """one lane, call by call, CUDA events around every launch: where a routed step spends its time (us, this rank)"""
        lane, be_l = lanes[0], lanes[0].be
        names = ["search.scatter+publish", "search.serve", "search.gather", "insert.scatter+publish", "insert.serve"]
        acc = [0.0] * 5
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        A = be_l.arena.ptr
        for i in range(count):
            b = (i * GROUP) % kd
            sq, iq, oq = sel_f[b * N_SEARCH:(b + GROUP) * N_SEARCH], ins_f[b * N_INSERT:(b + GROUP) * N_INSERT], out_f[b * N_SEARCH:(b + GROUP) * N_SEARCH]
            torch.cuda.synchronize(); dist.barrier()
            ev[0].record()
            be_l._p2p_scatter(lane, sq, 2, True); ev[1].record()
            be_l._p2p_serve(lane, 0, 2 * sq.shape[0]); ev[2].record()
            be_l._p2p_gather(lane, oq.shape[0], oq); ev[3].record()
            be_l._p2p_scatter(lane, iq, 3, False); ev[4].record()
            be_l._p2p_serve(lane, 1, 2 * iq.shape[0]); ev[5].record()
            torch.cuda.synchronize()
            for k in range(5):
                acc[k] += ev[k].elapsed_time(ev[k + 1]) * 1e3
        return {n: round(a / count, 2) for n, a in zip(names, acc)}

    use_graph = bool(args.graph)
    try:
        timed(None, 0, warm, use_graph)                                 # warm-up
    except Exception as e:                                              # graph capture is an optimisation, not a dependency
        log(f"graph capture failed ({e}); running eagerly")
        use_graph = False
        timed(None, 0, warm, False)
    sampler = B.ClockSampler(local_rank)
    with sampler:
        t_val = timed(None, warm, steps, use_graph)
    value = world * steps * BATCH / t_val / 1e6
    err = be.p2p_error() + sum(l.be.p2p_error() for l in lanes + ulanes)
    assert err == 0, "a flag wait timed out"
    chk = out[(warm + steps - 1) % kd].cpu().numpy().view(np.uint32)
    hit = float(((chk[:, 0] != 0) | (chk[:, 1] != 0)).mean())
    assert hit > 0.999, f"searches did not hit: {hit}"

    if quick:
        if rank == 0:
            B.emit({"quick": True, "n_gpus": world, "lanes": S, "group": GROUP, "graph": use_graph, "wait_mode": L.gpuhash_wait_mode(), "value_Mops": round(value, 1),
                    "us_per_step": round(t_val / steps * 1e6, 2)})
        dist.barrier(); dist.destroy_process_group()
        return 0
    with sampler:
        t_s = timed(None, warm, steps, use_graph, with_insert=False)      # search kernel path only (roofline)
    phases = phase_profile()
    # NCCL baseline on fewer steps (host sync per exchange)
    kb = min(steps, 10 * GROUP)
    timed(ixc, 0, GROUP, False)
    t_nccl = timed(ixc, warm, kb, False)

    # e2e: pinned host -> routed lookup -> pinned host, every exchange of GROUP batches, two ways:
    #   staged     H2D copy, routed search + insert on device buffers, D2H copy
    #   zero_copy  the scatter kernel reads the requests from the pinned host arrays itself and the gather kernel writes
    #              the results into the pinned host array (coalesced 256 B per warp over PCIe): no staging pass
    # each replayed as one CUDA graph (the exchanges of a fixed set of pinned batch buffers), eager as a fallback
    ke = min(-(-steps // GROUP) * GROUP, 16 * GROUP)
    hs = torch.empty((ke * N_SEARCH, 2), dtype=torch.int32).pin_memory(); hs.copy_(sel_f[: ke * N_SEARCH].cpu())
    hi = torch.empty((ke * N_INSERT, 3), dtype=torch.int32).pin_memory()
    ho = torch.empty((ke * N_SEARCH, 2), dtype=torch.int32).pin_memory()
    Se = min(S, 4)
    ds = torch.empty((Se, GROUP * N_SEARCH, 2), dtype=torch.int32, device=dev); di = torch.empty((Se, GROUP * N_INSERT, 3), dtype=torch.int32, device=dev)
    do = torch.empty((Se, GROUP * N_SEARCH, 2), dtype=torch.int32, device=dev)
    e2e_next = [pop + rank * (1 << 26) + N_INSERT * kd]

    def fresh_host_inserts():
        N.check(L.gpuhash_gen_inserts(ins.data_ptr(), None, SEED, e2e_next[0], N_INSERT * min(ke, kd), be._stream()))
        hi[: N_INSERT * min(ke, kd)].copy_(ins_f[: N_INSERT * min(ke, kd)].cpu())
        e2e_next[0] += N_INSERT * ke

    def e2e_issue(count, zero_copy):
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        c, i = 0, 0
        while i < count:
            b = i % ke
            g = min(GROUP, count - i, ke - b)
            k = c % (S if zero_copy else Se)
            hsl, hil, hol = hs[b * N_SEARCH:(b + g) * N_SEARCH], hi[b * N_INSERT:(b + g) * N_INSERT], ho[b * N_SEARCH:(b + g) * N_SEARCH]
            with torch.cuda.stream(streams[k]):
                if zero_copy:
                    lanes[k].search(hsl, hol); lanes[k].insert(hil)
                else:
                    ds[k][: g * N_SEARCH].copy_(hsl, non_blocking=True)
                    di[k][: g * N_INSERT].copy_(hil, non_blocking=True)
                    lanes[k].search(ds[k][: g * N_SEARCH], do[k][: g * N_SEARCH]); lanes[k].insert(di[k][: g * N_INSERT])
                    hol.copy_(do[k][: g * N_SEARCH], non_blocking=True)
            c += 1; i += g
        for st in streams:
            cur.wait_stream(st)

    def e2e(count, zero_copy, graph):
        fresh_host_inserts()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                e2e_issue(count, zero_copy)
            torch.cuda.synchronize(); dist.barrier()
            e0.record(); g.replay(); e1.record()
        else:
            e0.record(); e2e_issue(count, zero_copy); e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    e_steps = min(steps, ke)
    e2e_variants = {}
    for zero_copy, name in ((0, "staged+graph"), (1, "zero_copy+graph")):
        g_ok = use_graph
        try:
            e2e(min(ke, 2 * GROUP), zero_copy, g_ok)                   # warm-up
        except Exception as e:
            log(f"e2e {name}: graph capture failed ({e}); eager")
            g_ok = False
            e2e(min(ke, 2 * GROUP), zero_copy, False)
        ho.zero_()
        with sampler:
            t_e = e2e(e_steps, zero_copy, g_ok)
        got = ho[(e_steps - 1) * N_SEARCH: e_steps * N_SEARCH].numpy().view(np.uint32)
        ok_frac = float(((got[:, 0] != 0) | (got[:, 1] != 0)).mean())
        assert ok_frac > 0.999, f"e2e ({name}) results did not come back: {ok_frac}"
        e2e_variants[name if g_ok else name.replace("+graph", "")] = round(world * e_steps * BATCH / t_e / 1e6, 1)
        log(f"e2e {name}: {e_steps} steps in {t_e * 1e3:.2f} ms")
    e2e_path = max(e2e_variants, key=e2e_variants.get)
    e2e_val = e2e_variants[e2e_path]
    assert be.p2p_error() + sum(l.be.p2p_error() for l in lanes) == 0, "a flag wait timed out"

    if rank == 0:
        peak, peak_src = B.peaks()
        bytes_per_search = 8 + 2 * 32 + 32 * 1.0 + 8
        achieved = steps * N_SEARCH * bytes_per_search / t_s / 1e9      # per GPU
        cfg = B.workload_config(plan.mem_p_shard, args)
        cfg["workload"] = (f"configs[4]: {world}xB200 sharded index, logical table 2^{plan.mem_p_total} bytes "
                           f"(2^{plan.mem_p_shard} per GPU), keys routed by the top {log2w} bucket-index bits over NVLink; "
                           f"per GPU and step {N_SEARCH} searches + {N_INSERT} inserts")
        cfg["streams"] = S                                              # one stream per lane (exchange in flight)
        cfg.update({"mem_p_total": plan.mem_p_total, "exchange": "peer stores + flags (fused)", "cuda_graph": use_graph, "lanes": S,
                    "update_exchange": "own stream per lane, serve ordered behind the search serve" if split_updates else "same stream as the searches",
                    "batches_per_exchange": GROUP, "wait_mode": "stream mem-ops" if L.gpuhash_wait_mode() == 1 else "kernel",
                    "parallelism": f"shard{world}"})
        line = {
            "metric": "batched search/insert Mops/s (95/5 GET/SET, uniform keys)", "value": round(value, 1), "unit": "Mops/s",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": round(t_val / steps * 1e3, 6),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg,
            "e2e": {"value": round(e2e_val, 1), "unit": "Mops/s", "h2d_bytes_per_step": 8 * N_SEARCH + 12 * N_INSERT,
                    "d2h_bytes_per_step": 8 * N_SEARCH, "steps": e_steps, "path": e2e_path, "variants": e2e_variants},
            "gpu_launches": -(-steps // GROUP) * 5,                      # per rank: scatter, serve, gather + insert scatter, serve per exchange
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": None, "kernel": "serve_search_staged_kernel (per GPU, routed)", "peak_source": peak_src,
                         "note": "search path only, includes both NVLink exchanges"},
            "nccl_baseline": {"value": round(world * kb * BATCH / t_nccl / 1e6, 1), "unit": "Mops/s", "steps": kb,
                              "what": "same steps, exchanges through torch.distributed all_to_all_single"},
            "phase_us_one_lane": phases,
            "cpu_baseline": None, "clocks": sampler.summary(), "search_hit_fraction": round(hit, 5),
        }
        B.emit(line)
    dist.barrier()
    dist.destroy_process_group()
    return 0
