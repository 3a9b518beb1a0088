"""GPU suite: the launch-free scheduler cycle (gpuhash_ring_*, north_star (c)) against the oracle.

One persistent kernel, descriptor rings in pinned host memory: per ring the batches run strictly in order and each as
search -> delete -> insert (mega_scheduler.c:392-502 per stream); rings are unordered against each other.  Checked:
every search word of every batch, slot reuse with more batches than slots, request arrays on odd 8 B boundaries, the
kernel parking itself when idle and resuming, an explicit park with work pending, and the final table as a multiset."""
import ctypes as C
import time

import numpy as np
import pytest

import megakv_b200 as mk
from megakv_b200 import _native as N
from oracle import pyoracle as po
from tests import helpers as H

pytestmark = pytest.mark.gpu


class Pinned:
    """one cudaHostAlloc carved into arrays (the reference's batch buffers are pinned the same way, mega_recv.c:154-176)"""

    def __init__(self, nbytes):
        self.L = N.lib()
        self.base = self.L.gpuhash_host_alloc(nbytes)
        assert self.base
        self.nbytes, self.used = nbytes, 0

    def take(self, words, odd=False):
        """uint32 array of `words` words, 16 B aligned -- or, odd=True, 8 B past a 16 B boundary"""
        self.used = (self.used + 15) & ~15
        if odd:
            self.used += 8
        assert self.used + 4 * words <= self.nbytes
        addr = self.base + self.used
        self.used += 4 * words
        arr = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint32)), shape=(max(words, 1),))[:words]
        return addr, arr

    def free(self):
        self.L.gpuhash_host_free(self.base)


@pytest.mark.parametrize("layout", [mk.LAYOUT_PAIRS, mk.LAYOUT_REFERENCE], ids=["pairs", "reflayout"])
def test_ring_batches_match_oracle(gpu, layout):
    L = N.lib()
    mem_p, rings, slots, batches = 22, 3, 2, 7
    rng = np.random.default_rng(2026)
    table = mk.DeviceTable(mem_p, layout=layout)
    q = L.gpuhash_ring_create(C.byref(table.geom), table.ptr, rings, slots, 2, 400)
    assert q, "gpuhash_ring_create failed"
    pin = Pinned(64 << 20)
    o = po.Oracle(mem_p)
    try:
        universe = [H.random_requests(rng, 40000, loc_base=1 + 100000 * r) for r in range(rings)]   # disjoint keys per ring
        plan = []                                                  # (ring, ticket, out array, expected)
        state = [dict(inserted=0) for _ in range(rings)]
        for b in range(batches):
            for r in range(rings):
                u, st = universe[r], state[r]
                n_i = int(rng.integers(1, 4000)) if b != 3 else 0          # one batch without inserts
                n_s = int(rng.integers(1, 30000)) if b != 4 else 0         # one batch without searches
                n_d = int(rng.integers(0, 1500)) if st["inserted"] > 3000 else 0
                ins = u[st["inserted"]: st["inserted"] + n_i]
                pool = np.concatenate([u[: st["inserted"]], u[-2000:]])   # present keys (some deleted by now) + never-inserted ones
                sel = H.to_sel(pool[rng.integers(0, len(pool), n_s)]) if n_s else np.empty(0, mk.SEL_DT)
                dele = u[rng.integers(0, st["inserted"], n_d)] if n_d else np.empty(0, mk.IEL_DT)
                odd = (b + r) % 2 == 1
                a_s, v_s = pin.take(2 * n_s, odd); a_o, v_o = pin.take(2 * n_s, odd and b % 3 != 0)
                a_d, v_d = pin.take(3 * n_d); a_i, v_i = pin.take(3 * n_i)
                if n_s:
                    v_s[:] = sel.view(np.uint32)
                v_o[:] = 0xDEADBEEF
                if n_d:
                    v_d[:] = dele.view(np.uint32).reshape(-1)
                if n_i:
                    v_i[:] = ins.view(np.uint32).reshape(-1)
                want = o.search(sel) if n_s else np.empty(0, np.uint32)    # the oracle in the same per-ring order
                if n_d:
                    o.delete(dele)
                if n_i:
                    o.insert(ins)
                st["inserted"] += n_i
                t = L.gpuhash_ring_submit(q, r, a_s if n_s else None, n_s, a_o if n_s else None, a_d if n_d else None, n_d,
                                          a_i if n_i else None, n_i)
                assert t == b + 1, f"submit returned {t}"
                plan.append((r, t, v_o, want, b))
            if b == 2:
                time.sleep(1.0)                                    # > idle_ms: the kernel parks; the next submit relaunches it
            if b == 5:
                N.check(L.gpuhash_ring_park(q), "park")            # explicit park between cycles
        for r, t, v_o, want, b in plan:
            rc = L.gpuhash_ring_wait(q, r, t, 20000)
            assert rc == 0, f"wait(ring {r}, batch {t}) -> {rc}"
            assert np.array_equal(np.sort(v_o.reshape(-1, 2), 1), np.sort(want.reshape(-1, 2), 1)), f"ring {r} batch {b}"
        assert L.gpuhash_ring_drain(q, 20000) == 0
        N.check(L.gpuhash_ring_park(q), "park")
        assert o.digest(table=table.dump_reference()) == o.digest()
    finally:
        L.gpuhash_ring_destroy(q)
        pin.free()


def test_ring_park_with_pending_work_resumes(gpu):
    """batches whose doorbell rang while the kernel was leaving are not lost: the relaunch resumes at the first batch
    without a completion mark"""
    L = N.lib()
    mem_p = 20
    rng = np.random.default_rng(5)
    table = mk.DeviceTable(mem_p)
    q = L.gpuhash_ring_create(C.byref(table.geom), table.ptr, 1, 4, 1, 5000)
    assert q
    pin = Pinned(8 << 20)
    o = po.Oracle(mem_p)
    try:
        keys = H.random_requests(rng, 12000)
        # the SAME pinned buffers, refilled batch after batch with different requests: a kernel that never ends has no
        # launch boundary, so this is where a stale read of host memory would show
        a_s, v_s = pin.take(2 * 12000); a_o, v_o = pin.take(2 * 12000); a_i, v_i = pin.take(3 * 2000)
        for b in range(12):
            sel = H.to_sel(keys[rng.integers(0, 12000, 11000 + b)])
            ins = keys[(b % 6) * 2000:(b % 6 + 1) * 2000].copy()
            ins["loc"] += np.uint32(b * 100000)                      # later rounds update the same keys with new locations
            v_s[: 2 * len(sel)] = sel.view(np.uint32); v_i[:] = ins.view(np.uint32).reshape(-1); v_o[:] = 0xDEADBEEF
            want = o.search(sel); o.insert(ins)
            t = L.gpuhash_ring_submit(q, 0, a_s, len(sel), a_o, None, 0, a_i, len(ins))
            assert t == b + 1
            assert L.gpuhash_ring_wait(q, 0, t, 20000) == 0
            assert np.array_equal(v_o[: 2 * len(sel)], want), f"reused buffers, batch {t}"
            assert np.all(v_o[2 * len(sel):] == 0xDEADBEEF)
        N.check(L.gpuhash_ring_park(q), "park")
        L.gpuhash_ring_destroy(q)
        table.zero(); o = po.Oracle(mem_p)
        q = L.gpuhash_ring_create(C.byref(table.geom), table.ptr, 1, 4, 1, 5000)
        assert q
        outs = []
        for b in range(6):
            ins = keys[b * 2000:(b + 1) * 2000]
            sel = H.to_sel(keys[: (b + 1) * 2000])                 # includes this batch's own inserts: they must still miss
            a_s, v_s = pin.take(2 * len(sel)); a_o, v_o = pin.take(2 * len(sel)); a_i, v_i = pin.take(3 * len(ins))
            v_s[:] = sel.view(np.uint32); v_i[:] = ins.view(np.uint32).reshape(-1); v_o[:] = 0xDEADBEEF
            want = o.search(sel); o.insert(ins)
            t = L.gpuhash_ring_submit(q, 0, a_s, len(sel), a_o, None, 0, a_i, len(ins))
            assert t == b + 1
            outs.append((t, v_o, want))
            if b % 2 == 1:
                N.check(L.gpuhash_ring_park(q), "park")            # may leave the batch just submitted unprocessed
        for t, v_o, want in outs:
            assert L.gpuhash_ring_wait(q, 0, t, 20000) == 0
            assert np.array_equal(v_o, want), f"batch {t}"
    finally:
        L.gpuhash_ring_destroy(q)
        pin.free()
