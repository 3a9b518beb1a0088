#!/usr/bin/env python
"""What does the routed (sharded) path cost per GPU?  G virtual ranks on ONE GPU (sharded.LocalCluster): same kernels,
flags and buffers as the multi-process fused path, no NVLink.  Prints one JSON line per configuration:

  routed      S lanes (streams), each replaying exchanges of G x n searches (+ 5 % inserts), one CUDA graph
  scatter / serve / gather   the same kernels alone, R launches back to back over the lanes' buffers (pipelined
              rate of each kernel type: where the per-GPU gap to the direct path comes from)

usage: python tools/exp_routed_local.py [G=8] [lanes=4] [batches_per_exchange=16] [mem_p_total=34]
"""
import ctypes as C
import json
import os
import sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import megakv_b200 as mk
from megakv_b200 import _native as N
from megakv_b200.sharded import ShardPlan, LocalCluster

BATCH, N_SEARCH = 65536, 62259
N_INSERT = BATCH - N_SEARCH


def main():
    if os.environ.get("EXP_ARGS"):
        sys.argv[1:] = os.environ["EXP_ARGS"].split()
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    GROUP = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    mem_p_total = int(sys.argv[4]) if len(sys.argv) > 4 else 34
    cycles = int(os.environ.get("EXP_CYCLES", 24))
    L = mk.lib()
    torch.cuda.set_device(0); N.check(L.gpuhash_set_device(0))
    dev = torch.device("cuda", 0)
    plan = ShardPlan(mem_p_total, G)
    n_s, n_i = GROUP * N_SEARCH, GROUP * N_INSERT
    lanes = [LocalCluster(plan, cap=GROUP * BATCH)]
    for _ in range(S - 1):
        lanes.append(LocalCluster(plan, cap=GROUP * BATCH, tables=lanes[0].tables))
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    def stream_ptr():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # preload to LF 0.25 through the routed insert path
    pop = (1 << mem_p_total) // 8 // 4
    chunk = GROUP * BATCH
    gen = [torch.empty((chunk, 3), dtype=torch.int32, device=dev) for _ in range(G)]
    first = 0
    while first < pop:
        reqs = []
        for r in range(G):
            n = max(0, min(chunk, pop - first))
            if n:
                N.check(L.gpuhash_gen_inserts(gen[r].data_ptr(), None, 1, first, n, stream_ptr()))
            reqs.append(gen[r][:n]); first += n
        lanes[0].insert(reqs)
    torch.cuda.synchronize()
    assert lanes[0].error() == 0

    # per lane and rank: one resident request set (searches hit, inserts fresh)
    sel = [[torch.empty((n_s, 2), dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(S)]
    ins = [[torch.empty((n_i, 3), dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(S)]
    out = [[torch.empty((n_s, 2), dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(S)]
    for k in range(S):
        for r in range(G):
            N.check(L.gpuhash_gen_queries(sel[k][r].data_ptr(), None, 1, pop, n_s, 1000 + 17 * k + r, 0.0, 0.0, stream_ptr()))
            N.check(L.gpuhash_gen_inserts(ins[k][r].data_ptr(), None, 1, pop + (k * G + r) * n_i * (cycles + 8), n_i, stream_ptr()))
    torch.cuda.synchronize()

    def timed_graph(fn, reps=1):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()                       # warm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(reps):
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 1e3)
        return best

    def fan(fn_lane, count):
        """count calls of fn_lane(k) round-robin over the S lane streams, joined into the current stream"""
        cur = torch.cuda.current_stream()                          # the capturing stream while a graph is being recorded
        for st in streams:
            st.wait_stream(cur)
        for c in range(count):
            with torch.cuda.stream(streams[c % S]):
                fn_lane(c % S)
        for st in streams:
            cur.wait_stream(st)

    stagger = bool(int(os.environ.get("EXP_STAGGER", "0")))
    started = {}

    def routed(k, with_insert=True):
        """EXP_STAGGER=1: lane k's first exchange starts when lane k-1's first scatter phase is done, so the lanes do not
        run their phases in lockstep (all serving, then all scattering, ...)"""
        hook = None
        if stagger and k not in started:
            if k > 0 and (k - 1) in started:
                torch.cuda.current_stream().wait_event(started[k - 1])
            ev = torch.cuda.Event()
            started[k] = ev
            hook = lambda: ev.record(torch.cuda.current_stream())
        lanes[k].search(sel[k], out[k], after_scatter=hook)
        if with_insert:
            lanes[k].insert(ins[k])

    res = {"env": {k_: v_ for k_, v_ in os.environ.items() if k_.startswith("GPUHASH_")}, "G": G, "lanes": S, "batches_per_exchange": GROUP, "mem_p_total": mem_p_total, "searches_per_exchange": G * n_s}
    started.clear()
    t = timed_graph(lambda: fan(lambda k: routed(k), cycles))
    res["routed_Mops"] = round(cycles * G * GROUP * BATCH / t / 1e6, 1)
    started.clear()
    t = timed_graph(lambda: fan(lambda k: routed(k, False), cycles))
    res["routed_search_only_Mops"] = round(cycles * G * n_s / t / 1e6, 1)
    chk = out[0][0].cpu().numpy()
    res["hit"] = float(((chk[:, 0] != 0) | (chk[:, 1] != 0)).mean())
    assert sum(l.error() for l in lanes) == 0

    # the kernels alone (flags are already satisfied: no waits are issued here)
    def scatter_only(k):
        cl = lanes[k]
        for r in range(G):
            be = cl.be[r]; cl.ix[r].seq += 1
            if be.route_tiles:
                N.check(L.gpuhash_route_scatter_tiles(sel[k][r].data_ptr(), n_s, 2, plan.hash_mask_total, plan.log2, be.pp_peer_inbox,
                                                      be.arena.ptr + be.off_cnt2, be.perm.data_ptr(), be.cap, r, be.pp_peer_cnt, be.pp_peer_reqf,
                                                      be.arena.ptr + be.off_ticket, cl.ix[r].seq, stream_ptr()))
            else:
                N.check(L.gpuhash_route_scatter_pub(sel[k][r].data_ptr(), n_s, 2, plan.hash_mask_total, plan.log2, be.pp_peer_inbox,
                                                    be.arena.ptr + be.off_cnt2, be.perm.data_ptr(), be.cap, r, be.pp_peer_cnt, be.pp_peer_reqf,
                                                    be.arena.ptr + be.off_ticket, cl.ix[r].seq, None, None, stream_ptr()))

    def serve_only(k):
        cl = lanes[k]
        for r in range(G):
            be = cl.be[r]
            N.check(L.gpuhash_serve(C.byref(be.geom), be.table.ptr, 0, plan.log2, be.pp_my_inbox, be.arena.ptr + be.off_cnt,
                                    be.pp_origin_stage, 2 * n_s, None, None, r, be.pp_peer_resf, be.arena.ptr + be.off_ticket + 4,
                                    cl.ix[r].seq, None, stream_ptr()))

    def gather_only(k):
        cl = lanes[k]
        for r in range(G):
            cl.be[r]._p2p_gather(cl.ix[r], n_s, out[k][r])             # its flag wait is already satisfied

    # scatter-only LAST: it advances the sequence numbers without serves, after which a full exchange would wait forever
    for name, fn in (("serve", serve_only), ("gather", gather_only), ("scatter", scatter_only)) if not os.environ.get("EXP_SKIP_PARTS") else ():
        t = timed_graph(lambda: fan(fn, cycles))
        res[name + "_Mops"] = round(cycles * G * n_s / t / 1e6, 1)
        res[name + "_us_per_1M"] = round(t / (cycles * G * n_s) * 1e12, 2)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
