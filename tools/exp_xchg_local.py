#!/usr/bin/env python
"""What does the fused exchange kernel (gpuhash_xchg.cu) cost per GPU?  G virtual ranks on ONE GPU
(sharded.LocalExchangeCluster): same kernel, flags, slots and arena as one process per GPU, no NVLink.  Every step
launches the kernel once per virtual rank, back to back on one stream, so  per-GPU rate = requests of one rank per step /
(time per step / G).  Prints one JSON line: mixed (95/5) and search-only steps, `steps` exchanges + the two draining
launches per rank inside the timed region, as one CUDA graph.

usage: python tools/exp_xchg_local.py [G=8] [batches_per_exchange=64] [mem_p_total=34] [steps=12]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import megakv_b200 as mk
from megakv_b200 import _native as N
from megakv_b200.sharded import ShardPlan, LocalExchangeCluster

BATCH, N_SEARCH = 65536, 62259
N_INSERT = BATCH - N_SEARCH


def main():
    if os.environ.get("EXP_ARGS"):
        sys.argv[1:] = os.environ["EXP_ARGS"].split()
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    mem_p_total = int(sys.argv[3]) if len(sys.argv) > 3 else 34
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    L = mk.lib()
    torch.cuda.set_device(0); N.check(L.gpuhash_set_device(0))
    dev = torch.device("cuda", 0)
    plan = ShardPlan(mem_p_total, G)
    n_s, n_i = W * N_SEARCH, W * N_INSERT
    cl = LocalExchangeCluster(plan, cap_search=max(n_s, 1 << 20), cap_update=max(n_i, 1 << 20))

    def sp():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    pop = (1 << mem_p_total) // 8 // 4                                 # load factor 0.25
    chunk = 1 << 20
    gen = [torch.empty((chunk, 3), dtype=torch.int32, device=dev) for _ in range(G)]
    first = 0
    while first < pop:
        reqs = []
        for r in range(G):
            n = max(0, min(chunk, pop - first))
            if n:
                N.check(L.gpuhash_gen_inserts(gen[r].data_ptr(), None, 1, first, n, sp()))
            reqs.append(gen[r][:n]); first += n
        cl.step(inserts=reqs)
        torch.cuda.synchronize()                                         # gen[] is reused by the next round
    cl.flush(); torch.cuda.synchronize()
    assert cl.error() == 0

    R = 3                                                                # resident request sets per rank
    sel = [[torch.empty((n_s, 2), dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(R)]
    exp = [[torch.empty((n_s,), dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(R)]
    out = [[torch.empty((n_s, 2), dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(R)]
    ins = [[torch.empty((n_i, 3), dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(steps + 4)]
    for k in range(R):
        for r in range(G):
            N.check(L.gpuhash_gen_queries(sel[k][r].data_ptr(), exp[k][r].data_ptr(), 1, pop, n_s, 1000 + 17 * k + r, 0.0, 0.0, sp()))
    nxt = [pop]

    def fresh():
        for k in range(len(ins)):
            for r in range(G):
                N.check(L.gpuhash_gen_inserts(ins[k][r].data_ptr(), None, 1, nxt[0], n_i, sp())); nxt[0] += n_i
        torch.cuda.synchronize()

    def run(count, with_insert):
        for c in range(count):
            cl.step(sel[c % R], out[c % R], None, ins[c] if with_insert else None)
        cl.flush()

    def timed(count, with_insert):
        fresh()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run(count, with_insert)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3

    timed(3, True)                                                       # warm-up
    res = {"exp": "xchg_local", "env": {k_: v_ for k_, v_ in os.environ.items() if k_.startswith("GPUHASH_")}, "G": G,
           "batches_per_exchange": W, "mem_p_total": mem_p_total, "steps": steps, "requests_per_rank_and_step": W * BATCH}
    t = min(timed(steps, True) for _ in range(2))
    res["mixed_us_per_rank_step"] = round(t / steps / G * 1e6, 1)
    res["mixed_Mops_per_gpu"] = round(steps * W * BATCH / (t / G) / 1e6, 1)
    t = min(timed(steps, False) for _ in range(2))
    res["search_us_per_rank_step"] = round(t / steps / G * 1e6, 1)
    res["search_Mops_per_gpu"] = round(steps * n_s / (t / G) / 1e6, 1)
    # each kind of work alone: a step with requests followed by two without -> launch 1 only scatters, 2 only serves, 3 only gathers
    if G == 1:
        acc = [0.0, 0.0, 0.0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        rounds = 5
        for i in range(rounds):
            torch.cuda.synchronize()
            ev[0].record(); cl.step(sel[i % R], out[i % R])
            ev[1].record(); cl.step()
            ev[2].record(); cl.step()
            ev[3].record(); torch.cuda.synchronize()
            if i:
                for k in range(3):
                    acc[k] += ev[k].elapsed_time(ev[k + 1]) * 1e3 / (rounds - 1)
        res["alone_us"] = {"scatter": round(acc[0], 1), "serve_search": round(acc[1], 1), "gather": round(acc[2], 1)}
    bad = 0
    for k in range(min(R, steps)):
        for r in range(G):
            o, e = out[k][r], exp[k][r]
            good = ((o[:, 0] == e) & ((o[:, 1] == 0) | (o[:, 1] == e))) | ((o[:, 1] == e) & (o[:, 0] == 0))
            bad += int((~good).sum())
    res["mismatches"] = bad
    assert cl.error() == 0
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
