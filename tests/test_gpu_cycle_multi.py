"""GPU parity of the one-launch scheduler cycle over ALL workers (gpuhash_cycle_multi_ex, gpuhash_index_submit_all).

Reference semantics (src/mega_scheduler.c:393-504): per worker search -> delete -> insert in stream order, workers
unordered against each other, one synchronisation per cycle.  The oracle is applied worker by worker in that order; the
test workloads give every worker its own keys, so the result does not depend on how the workers interleave.
"""
import ctypes as C
import threading

import numpy as np
import pytest

import megakv_b200 as mk
from megakv_b200 import _native as N
from oracle import pyoracle as po
from tests import helpers as H

pytestmark = pytest.mark.gpu

LAYOUTS = [mk.LAYOUT_PAIRS, mk.LAYOUT_REFERENCE]


@pytest.fixture(params=LAYOUTS, ids=["pairs", "reflayout"])
def layout(request):
    return request.param


def sorted_pairs(words):
    return np.sort(np.asarray(words).reshape(-1, 2), axis=1)


def make_cycle(rng, live_by_worker, loc0, n_fresh, n_del):
    """per worker: searches for half of its live keys + its fresh keys (which must miss), deletes, fresh inserts"""
    batches, loc = [], loc0
    for live in live_by_worker:
        fresh = H.random_requests(rng, n_fresh, loc_base=loc); loc += n_fresh
        dele = live[rng.permutation(len(live))[:n_del]] if len(live) else live[:0]
        sel = np.concatenate([H.to_sel(live[::2]), H.to_sel(fresh)])
        batches.append({"search": sel, "delete": dele, "insert": fresh})
    return batches, loc


@pytest.mark.parametrize("zero_copy", [0, 1], ids=["staged", "zero_copy"])
@pytest.mark.parametrize("workers", [1, 5, 16])
def test_submit_all_matches_oracle(gpu, layout, workers, zero_copy, rng):
    mem_p = 23                                                        # load factor < 0.1: no bucket ever fills, so no worker can evict another's key
    ix = mk.GpuHashIndex(mem_p, workers=workers, max_search=1 << 15, max_insert=1 << 13, max_delete=1 << 13, layout=layout)
    ix.enable_stats(True)
    o = po.Oracle(mem_p)
    live = [H.random_requests(rng, 3000 + 37 * w, loc_base=1 + 100000 * w) for w in range(workers)]
    for w in range(workers):
        ix.insert(live[w]); o.insert(live[w])
    ix.L.gpuhash_index_set_zero_copy(ix.h, zero_copy)               # from here on every host buffer handed in is pinned
    loc0 = 100000 * workers + 1
    for cyc in range(4):
        n_fresh = [700, 64, 1, 0][cyc]                              # ragged sizes: partial tiles, one request, empty parts
        n_del = [500, 0, 17, 129][cyc]
        batches, loc0 = make_cycle(rng, live, loc0, n_fresh, n_del)
        want = [o.search(b["search"]) for b in batches]             # searches see the table before this cycle's updates
        zeroed = 0
        for b in batches:
            zeroed += o.delete(b["delete"]); o.insert(b["insert"])
        st0 = ix.stats()
        if zero_copy:                                               # the kernel reads/writes these host arrays itself: pin them
            ticket, outs = submit_all_pinned(ix, batches)
        else:
            ticket, outs = ix.submit_all(batches)
        ix.wait(ticket)
        for w in range(workers):
            assert np.array_equal(sorted_pairs(outs[w]), sorted_pairs(want[w])), f"cycle {cyc} worker {w}"
            if n_fresh:
                assert not np.asarray(outs[w]).reshape(-1, 2)[-n_fresh:].any()   # this cycle's inserts are not visible yet
        assert ix.stats()["del_zeroed"] - st0["del_zeroed"] == zeroed
        assert o.digest(table=ix.dump()) == o.digest()
        for w, b in enumerate(batches):
            keep = ~np.isin(live[w]["loc"], b["delete"]["loc"])
            live[w] = np.concatenate([live[w][keep], b["insert"]])
        ix._keep.clear()
    ix.close()


class Pinned:
    """numpy view of cudaHostAlloc'd memory (the reference's batch buffers are pinned, mega_recv.c:154-156,176)"""

    def __init__(self, nbytes):
        self.n = max(int(nbytes), 16)
        self.ptr = N.lib().gpuhash_host_alloc(self.n)
        assert self.ptr
        self.u8 = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(self.n,))

    def free(self):
        if self.ptr:
            N.lib().gpuhash_host_free(self.ptr); self.ptr = None


def submit_all_pinned(ix, batches, keep=None):
    descs = (N.Batch * len(batches))()
    outs, pins = [], []
    for w, b in enumerate(batches):
        arrs = []
        for key, dt in (("search", mk.SEL_DT), ("delete", mk.IEL_DT), ("insert", mk.IEL_DT)):
            a = np.ascontiguousarray(b.get(key) if b.get(key) is not None else np.empty(0, dt), dtype=dt)
            p = Pinned(a.nbytes); p.u8[: a.nbytes] = a.view(np.uint8).reshape(-1)
            arrs.append((a, p)); pins.append(p)
        (s, ps), (d, pd), (i, pi) = arrs
        po_ = Pinned(8 * len(s)); po_.u8[:] = 0xEE; pins.append(po_)
        descs[w] = N.Batch(ps.ptr if len(s) else None, po_.ptr if len(s) else None, pd.ptr if len(d) else None,
                           pi.ptr if len(i) else None, len(s), len(d), len(i), 0)
        outs.append(po_.u8[: 8 * len(s)].view(np.uint32))
    ticket = ix.L.gpuhash_index_submit_all(ix.h, descs, len(batches))
    assert ticket >= 0, ticket
    ix._keep.append(pins)
    return ticket, outs


def test_cycle_multi_device_descriptors_and_compact(gpu, layout, rng):
    """gpuhash_cycle_multi_ex on caller-owned device buffers, misaligned (8 B but not 16 B) request arrays, compact results"""
    L = N.lib()
    mem_p = 23                                                        # load factor < 0.05: placement does not depend on the insert order
    t = mk.DeviceTable(mem_p, layout=layout)
    o = po.Oracle(mem_p)
    base = H.random_requests(rng, 20000)
    ins_d = mk.DeviceBuffer.from_host(base)
    mk.insert_flat_ex(t.geom, t, ins_d, len(base)); mk.device_sync()
    o.insert(base)
    W = 7
    ws = mk.DeviceBuffer(L.gpuhash_cycle_workspace_bytes(W), zero=True)
    for compact in (0, 1):
        descs = (N.Batch * W)()
        keep, sels, fresh_all = [], [], []
        for w in range(W):
            n = [1, 63, 64, 65, 1000, 4097, 0][w]
            sel = H.to_sel(base[rng.integers(0, len(base), n)])
            sel["sig"][::5] ^= 0x5a5a5a5a                            # some misses
            raw = np.zeros(8 * (n + 1), dtype=np.uint8)               # shift by one request: 8 B aligned, not 16 B
            raw[8:] = sel.view(np.uint8).reshape(-1)
            in_d = mk.DeviceBuffer.from_host(raw) if n else mk.DeviceBuffer(16)
            out_d = mk.DeviceBuffer(8 * (n + 2)); out_d.upload(np.full(2 * (n + 2), 0xDEADBEEF, dtype=np.uint32))
            fresh = H.random_requests(rng, 50 + w, loc_base=10**6 + 1000 * (w + 10 * compact))
            f_d = mk.DeviceBuffer.from_host(fresh)
            keep += [in_d, out_d, f_d]; sels.append(sel); fresh_all.append(fresh)
            descs[w] = N.Batch(in_d.ptr + 8 if n else None, out_d.ptr + 8 if n else None, None, f_d.ptr, n, 0, len(fresh), 0)
        d_d = mk.DeviceBuffer.from_host(np.frombuffer(bytes(descs), dtype=np.uint8))
        N.check(L.gpuhash_cycle_multi_ex(C.byref(t.geom), t.ptr, descs, d_d.ptr, W, compact, ws.ptr, None, None))
        mk.device_sync()
        assert L.gpuhash_cycle_error(1) == 0
        assert not ws.download(np.uint32).any()                      # the launch left its workspace zero
        for w in range(W):
            n = len(sels[w])
            want = o.search(sels[w]).reshape(-1, 2)
            got = keep[3 * w + 1].download(np.uint32)
            assert got[0] == 0xDEADBEEF and got[1] == 0xDEADBEEF     # nothing in front of the output array was touched
            if compact:
                assert np.array_equal(got[2: 2 + n], np.where(want[:, 0] != 0, want[:, 0], want[:, 1]))
                assert (got[2 + n: 2 + n + 2] == 0xDEADBEEF).all()
            else:
                assert np.array_equal(got[2: 2 + 2 * n].reshape(-1, 2), want)
                assert (got[2 + 2 * n:] == 0xDEADBEEF).all()
        for f in fresh_all:
            o.insert(f)
        assert o.digest(table=t.dump_reference()) == o.digest()


def test_cycle_same_key_search_delete_insert_order(gpu, layout, rng):
    """one worker searches, deletes and re-inserts THE SAME keys in one cycle: the search must see the old location, the
    table must end with the new one (search -> delete -> insert, mega_scheduler.c:392-502)"""
    mem_p = 23                                                        # load factor < 0.03: every key stays findable
    ix = mk.GpuHashIndex(mem_p, workers=3, max_search=1 << 15, max_insert=1 << 15, max_delete=1 << 15, layout=layout)
    o = po.Oracle(mem_p)
    base = [H.random_requests(rng, 9000, loc_base=1 + 20000 * w) for w in range(3)]
    for w in range(3):
        ix.insert(base[w]); o.insert(base[w])
    for rep in range(6):
        batches = []
        for w in range(3):
            new = base[w].copy(); new["loc"] += 7 + rep
            batches.append({"search": H.to_sel(base[w]), "delete": base[w], "insert": new})
        ticket, outs = ix.submit_all(batches)
        ix.wait(ticket)
        for w in range(3):
            got = np.asarray(outs[w]).reshape(-1, 2)
            assert (((got[:, 0] == base[w]["loc"]) | (got[:, 1] == base[w]["loc"]))).all(), f"rep {rep} worker {w}: stale or future value seen"
            o.delete(base[w]); o.insert(batches[w]["insert"])
            base[w] = batches[w]["insert"]
        assert o.digest(table=ix.dump()) == o.digest()
        ix._keep.clear()
    ix.close()


def test_concurrent_cycle_launches_do_not_deadlock(gpu, rng):
    """32 cycle kernels in flight at once on 32 streams (each far more CTAs than one wave when run alone), issued from four
    host threads, plus a long-running kernel hogging SMs: the phase waits only ever target tiles already held by running
    warps (atomic tickets), so every launch finishes and every result is right.  The stateless entry point takes its
    workspace from the shared pool (atomic slot counter)."""
    L = N.lib()
    mem_p = 24                                                        # load factor stays below 0.2: no evictions, the multiset is order-free
    t = mk.DeviceTable(mem_p)
    o = po.Oracle(mem_p)
    n_launch, n_s, n_i = 32, 40000, 6000
    base = H.random_requests(rng, 150000)
    b_d = mk.DeviceBuffer.from_host(base)
    mk.insert_flat_ex(t.geom, t, b_d, len(base)); mk.device_sync()
    o.insert(base)
    streams = [L.gpuhash_stream_create() for _ in range(n_launch)]
    sels, fresh, bufs = [], [], []
    for k in range(n_launch):
        sel = H.to_sel(base[rng.integers(0, len(base), n_s)])
        f = H.random_requests(rng, n_i, loc_base=10**6 + k * n_i)
        sels.append(sel); fresh.append(f)
        bufs.append((mk.DeviceBuffer.from_host(sel), mk.DeviceBuffer(8 * n_s), mk.DeviceBuffer.from_host(f)))
    want = [o.search(s) for s in sels]
    errs = []

    def issue(ks):
        try:
            for k in ks:
                for rep in range(3):                                  # same stream: three cycles back to back (searches only repeat)
                    N.check(L.gpuhash_cycle_ex(C.byref(t.geom), t.ptr, bufs[k][0].ptr, n_s, bufs[k][1].ptr, None, 0,
                                               bufs[k][2].ptr if rep == 0 else None, n_i if rep == 0 else 0, None, None, 0, None, streams[k]))
        except Exception as e:                                        # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=issue, args=(range(q, n_launch, 4),)) for q in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    mk.device_sync()
    assert not errs, errs
    assert L.gpuhash_cycle_error(1) == 0, "a phase wait timed out"
    for k in range(n_launch):
        got = bufs[k][1].download(np.uint32)
        # searches of launch k may or may not see inserts of OTHER launches (fresh keys are never searched): exact
        assert np.array_equal(sorted_pairs(got), sorted_pairs(want[k])), f"launch {k}"
    for f in fresh:
        o.insert(f)
    assert o.digest(table=t.dump_reference()) == o.digest()
    for s_ in streams:
        L.gpuhash_stream_destroy(s_)


def test_cycle_legacy_segments_two_lanes(gpu, layout, rng):
    """gpu_delete_insert semantics through gpuhash_cycle_ws_ex with a caller-owned workspace: device-side segment counts,
    an empty segment, a segment pointer table in device memory (mega_recv.c:148-149, mega_scheduler.c:484-494)"""
    L = N.lib()
    mem_p = 23                                                        # low load factor: placement is order-free, every delete hits
    t = mk.DeviceTable(mem_p, layout=layout)
    o = po.Oracle(mem_p)
    base = H.random_requests(rng, 25000)
    b_d = mk.DeviceBuffer.from_host(base)
    mk.insert_flat_ex(t.geom, t, b_d, len(base)); mk.device_sync(); o.insert(base)
    ws = mk.DeviceBuffer(L.gpuhash_cycle_workspace_bytes(1), zero=True)
    st = mk.DeviceStats()
    zeroed = 0
    for rep in range(3):
        dele = base[rng.permutation(len(base))[:3000]]
        fresh = H.random_requests(rng, 5000, loc_base=10**6 + 10000 * rep)
        blocks = mk.split_insert_blocks(fresh, 8); blocks[rep] = blocks[rep][:0]
        segs = mk.InsertSegments(blocks)
        sel = H.to_sel(base[:4000])
        s_d, o_d, d_d = mk.DeviceBuffer.from_host(sel), mk.DeviceBuffer(8 * len(sel)), mk.DeviceBuffer.from_host(dele)
        want = o.search(sel)
        N.check(L.gpuhash_cycle_ws_ex(C.byref(t.geom), t.ptr, s_d.ptr, len(sel), o_d.ptr, d_d.ptr, len(dele), None, 0,
                                      segs.ptrs.ptr, segs.nums.ptr, 8, 0, ws.ptr, st.ptr, None))
        mk.device_sync()
        assert np.array_equal(sorted_pairs(o_d.download(np.uint32)), sorted_pairs(want))
        zeroed += o.delete(dele); o.insert_blocks(blocks)
        assert o.digest(table=t.dump_reference()) == o.digest()
        base = np.concatenate([base[~np.isin(base["loc"], dele["loc"])]] + blocks)      # (the emptied segment's keys never went in)
    assert st.read()["del_zeroed"] == zeroed == 9000


def test_submit_all_staged_coalesces_adjacent_host_batches(gpu, layout, rng):
    """Staged mode with all workers' arrays carved out of ONE host block (worker w+1's array starts where worker w's ends):
    gpuhash_index_submit_all moves each array kind with one copy per run of adjacent batches.  Runs are broken on purpose
    (a gap after worker 2, an empty batch, a worker with no searches), four cycles in flight (one per slot), results and
    table against the oracle."""
    mem_p, workers = 23, 7
    ix = mk.GpuHashIndex(mem_p, workers=workers, max_search=1 << 14, max_insert=1 << 12, max_delete=1 << 12, layout=layout)
    o = po.Oracle(mem_p)
    live = [H.random_requests(rng, 2500 + 11 * w, loc_base=1 + 100000 * w) for w in range(workers)]
    for w in range(workers):
        ix.insert(live[w]); o.insert(live[w])
    loc0 = 100000 * workers + 1
    pending = []
    for cyc in range(6):
        batches, loc0 = make_cycle(rng, live, loc0, [300, 64, 1, 0, 257, 33][cyc], [200, 0, 17, 129, 64, 5][cyc])
        if cyc == 3:
            batches[4]["search"] = batches[4]["search"][:0]
        want = [o.search(b["search"]) for b in batches]
        for b in batches:
            o.delete(b["delete"]); o.insert(b["insert"])
        # one block per array kind; worker 3 starts 24 bytes late so that its run is not adjacent to worker 2's
        def carve(key, dt, item):
            total = sum(len(b[key]) for b in batches) * item + 24
            block = np.zeros(total, dtype=np.uint8)
            views, off = [], 0
            for w, b in enumerate(batches):
                if w == 3:
                    off += 24
                a = np.ascontiguousarray(b[key], dtype=dt).view(np.uint8).reshape(-1)
                block[off:off + len(a)] = a
                views.append((off, len(b[key])))
                off += len(a)
            return block, views
        sblk, sv = carve("search", mk.SEL_DT, 8); dblk, dv_ = carve("delete", mk.IEL_DT, 12); iblk, iv = carve("insert", mk.IEL_DT, 12)
        oblk = np.full(len(sblk), 0xEE, dtype=np.uint8)
        descs = (N.Batch * workers)()
        for w in range(workers):
            descs[w] = N.Batch(sblk.ctypes.data + sv[w][0] if sv[w][1] else None, oblk.ctypes.data + sv[w][0] if sv[w][1] else None,
                               dblk.ctypes.data + dv_[w][0] if dv_[w][1] else None, iblk.ctypes.data + iv[w][0] if iv[w][1] else None,
                               sv[w][1], dv_[w][1], iv[w][1], 0)
        ticket = ix.L.gpuhash_index_submit_all(ix.h, descs, workers)
        assert ticket >= 0, ticket
        pending.append((ticket, oblk, sv, want, (sblk, dblk, iblk, descs)))
        if len(pending) == 4 or cyc == 5:                            # up to GPUHASH_INDEX_SLOTS cycles in flight
            for (tk, ob, svv, wnt, _keep) in pending:
                ix.wait(tk)
                for w in range(workers):
                    got = ob[svv[w][0]: svv[w][0] + 8 * svv[w][1]].view(np.uint32)
                    assert np.array_equal(sorted_pairs(got), sorted_pairs(wnt[w])), f"cycle worker {w}"
            pending = []
        for w, b in enumerate(batches):
            keep = ~np.isin(live[w]["loc"], b["delete"]["loc"])
            live[w] = np.concatenate([live[w][keep], b["insert"]])
    assert o.digest(table=ix.dump()) == o.digest()
    ix.close()


def test_unordered_cycles_switch_keeps_independent_cycles_exact(gpu, rng):
    """gpuhash_index_set_unordered_cycles(1): consecutive cycle kernels may overlap; cycles whose requests do not depend on
    each other (searches of keys that were there before, inserts of fresh keys -- the benchmark's workload) stay exact."""
    mem_p, workers = 23, 4
    ix = mk.GpuHashIndex(mem_p, workers=workers, max_search=1 << 14, max_insert=1 << 12, max_delete=1 << 12)
    assert ix.L.gpuhash_index_set_unordered_cycles(ix.h, 1) == 0
    o = po.Oracle(mem_p)
    old = [H.random_requests(rng, 4000, loc_base=1 + 100000 * w) for w in range(workers)]
    for w in range(workers):
        ix.insert(old[w]); o.insert(old[w])
    tickets, loc0 = [], 10**6
    for cyc in range(4):                                             # four cycles in flight, nothing waited for in between
        batches = []
        for w in range(workers):
            fresh = H.random_requests(rng, 500, loc_base=loc0); loc0 += 500
            batches.append({"search": H.to_sel(old[w][cyc::4]), "insert": fresh})
            o.insert(fresh)
        tickets.append((ix.submit_all(batches), [o.search(b["search"]) for b in batches]))
    for (ticket, outs), want in tickets:
        ix.wait(ticket)
        for w in range(workers):
            assert np.array_equal(sorted_pairs(outs[w]), sorted_pairs(want[w]))
    assert o.digest(table=ix.dump()) == o.digest()
    ix.close()
