"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Bar (BASELINE.json north_star): search results bit-exact per request; delete hit counts exact; table
contents exact -- slot for slot where the request order is defined (serial mode, conflict-free batches),
as a multiset where concurrent requests may legally take each other's slots.

Nothing here reads /root/reference.  The oracle is the checker, never the thing under test.
"""
import ctypes as C
import os

import numpy as np
import pytest

import megakv_b200 as mk
from megakv_b200 import _native as N
from oracle import pyoracle as po
from tests import helpers as H
from tests.golden import make_golden as MG

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def gpu_search(table, sel, prezero=True):
    sel = np.ascontiguousarray(sel, dtype=mk.SEL_DT)
    in_d = mk.DeviceBuffer.from_host(sel) if len(sel) else mk.DeviceBuffer(8)
    out_d = mk.DeviceBuffer(max(8 * len(sel), 8))
    if prezero:
        out_d.zero()
    else:                                                     # the kernel must define every word itself
        out_d.upload(np.full(2 * max(len(sel), 1), 0xDEADBEEF, dtype=np.uint32))
    mk.search_ex(table.geom, in_d, out_d, table, len(sel))
    mk.device_sync()
    return out_d.download(np.uint32, 2 * len(sel))


def gpu_insert(table, iel, flags=0, stats=None):
    iel = np.ascontiguousarray(iel, dtype=mk.IEL_DT)
    in_d = mk.DeviceBuffer.from_host(iel) if len(iel) else mk.DeviceBuffer(12)
    mk.insert_flat_ex(table.geom, table, in_d, len(iel), stats=stats, flags=flags)
    mk.device_sync()


def gpu_delete(table, iel, flags=0):
    iel = np.ascontiguousarray(iel, dtype=mk.IEL_DT)
    st = mk.DeviceStats()
    in_d = mk.DeviceBuffer.from_host(iel) if len(iel) else mk.DeviceBuffer(12)
    mk.delete_ex(table.geom, in_d, table, len(iel), stats=st, flags=flags)
    mk.device_sync()
    return st.read()["del_zeroed"]


LAYOUTS = [mk.LAYOUT_PAIRS, mk.LAYOUT_REFERENCE]


@pytest.fixture(params=LAYOUTS, ids=["pairs", "reflayout"])
def layout(request):
    return request.param


def table_from_oracle(o, layout=mk.LAYOUT_PAIRS):
    t = mk.DeviceTable(o.mem_p, o.algo, layout)
    t.load_reference(o.table)
    return t


# ----------------------------------------------------------------------------- search

@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
@pytest.mark.parametrize("mem_p,load", [(16, 0.5), (20, 0.9), (24, 0.3)])
def test_search_bit_exact_on_oracle_built_tables(gpu, layout, algo, mem_p, load, rng):
    o = po.Oracle(mem_p, algo)
    n = int(load * (1 << mem_p) / 8)
    iel = H.random_requests(rng, n)
    o.insert(iel)
    t = table_from_oracle(o, layout)
    miss = H.random_requests(rng, n // 4 + 1, loc_base=1)
    sel = np.concatenate([H.to_sel(iel), H.to_sel(miss)])
    rng.shuffle(sel)
    want = o.search(sel)
    assert np.array_equal(gpu_search(t, sel), want)
    assert np.array_equal(gpu_search(t, sel, prezero=False), want)      # misses are written as 0


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 255, 256, 257, 65535, 65536, 65537, 1000003])
def test_search_ragged_sizes(gpu, layout, n, rng):
    o = po.Oracle(18)
    iel = H.random_requests(rng, 20000)
    o.insert(iel)
    t = table_from_oracle(o, layout)
    sel = H.to_sel(iel)[rng.integers(0, len(iel), n)]
    assert np.array_equal(gpu_search(t, sel, prezero=False), o.search(sel))


@pytest.mark.parametrize("qpt", [1, 2, 4, -1, -4, -5, -6])
@pytest.mark.parametrize("split_mode", [1, 2])
def test_search_every_launch_variant(gpu, layout, qpt, split_mode, rng):
    o = po.Oracle(20)
    iel = H.random_requests(rng, 100000)
    o.insert(iel)
    t = table_from_oracle(o, layout)
    sel = np.concatenate([H.to_sel(iel), H.to_sel(H.random_requests(rng, 30001))])
    old = N.Tune(); N.lib().gpuhash_get_tuning(old)
    try:
        N.lib().gpuhash_set_tuning(N.Tune(qpt, split_mode, 4))
        assert np.array_equal(gpu_search(t, sel, prezero=False), o.search(sel))
    finally:
        N.lib().gpuhash_set_tuning(old)


@pytest.mark.parametrize("qpt", [-5, -6], ids=["staged", "warp"])
@pytest.mark.parametrize("in_off,out_off", [(0, 0), (1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 128, 129, 4097, 62259, 300001])
def test_search_staged_kernel_alignment_and_tails(gpu, layout, n, in_off, out_off, qpt, rng):
    """search_quad_staged_kernel moves the batch in 512 B bulk copies and search_warp_kernel in 512 B vector accesses
    per warp; both need 16 B-aligned tiles: request and
    result arrays that start on an odd 8 B boundary (every second batch of bench.py does: 62259 * 8 is not a multiple
    of 16), a lone head request, partial last tiles, more tiles than CTAs -- all bit-exact against the oracle, and
    nothing outside out[0 .. 2n) is written."""
    o = po.Oracle(18)
    iel = H.random_requests(rng, 20000)
    o.insert(iel)
    t = table_from_oracle(o, layout)
    sel = np.concatenate([H.to_sel(iel), H.to_sel(H.random_requests(rng, 5000))])[rng.integers(0, 25000, n)]
    in_d = mk.DeviceBuffer(8 * (n + 4)); out_d = mk.DeviceBuffer(8 * (n + 4))
    pad = np.zeros(n + 4, dtype=mk.SEL_DT); pad[in_off:in_off + n] = sel
    in_d.upload(pad)
    out_d.upload(np.full(2 * (n + 4), 0xDEADBEEF, dtype=np.uint32))
    old = N.Tune(); N.lib().gpuhash_get_tuning(old)
    try:
        N.lib().gpuhash_set_tuning(N.Tune(qpt, 0, 4))
        N.check(N.lib().gpuhash_search_ex(C.byref(t.geom), in_d.ptr + 8 * in_off, out_d.ptr + 8 * out_off, t.ptr, n, None, None))
        mk.device_sync()
    finally:
        N.lib().gpuhash_set_tuning(old)
    got = out_d.download(np.uint32)
    assert np.array_equal(got[2 * out_off:2 * (out_off + n)], o.search(sel))
    assert np.all(got[:2 * out_off] == 0xDEADBEEF) and np.all(got[2 * (out_off + n):] == 0xDEADBEEF)


@pytest.mark.parametrize("fused", [0, 1])
def test_index_zero_copy_submit_matches_oracle(gpu, layout, fused, rng):
    """gpuhash_index_set_zero_copy: the kernels read requests from and write results to PINNED HOST buffers themselves
    (bulk copies over the host link).  Same results as the staged-copy path and as the oracle."""
    L = N.lib()
    mem_p, graph_like_batches = 22, 3                             # load factor 0.13: no bucket overflows, so placement is order-free
    old = N.Tune(); L.gpuhash_get_tuning(old)
    L.gpuhash_set_tuning(N.Tune(0, 0, 4, fused))                  # 0: search/delete/insert launches (staged search kernel)
    o = po.Oracle(mem_p)
    ix = mk.GpuHashIndex(mem_p, workers=2, max_search=1 << 17, max_insert=1 << 16, max_delete=1 << 16, layout=layout)
    base = H.random_requests(rng, 60000)
    ix.insert(base); o.insert(base)
    L.gpuhash_index_set_zero_copy(ix.h, 1)
    n_s, n_i, n_d = 62259, 3277, 1000
    hs = L.gpuhash_host_alloc(8 * n_s * graph_like_batches); ho = L.gpuhash_host_alloc(8 * n_s * graph_like_batches)
    hi = L.gpuhash_host_alloc(12 * n_i * graph_like_batches); hd = L.gpuhash_host_alloc(12 * n_d * graph_like_batches)
    assert hs and ho and hi and hd
    try:
        s_np = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint32)), shape=(graph_like_batches, 2 * n_s))
        o_np = np.ctypeslib.as_array(C.cast(ho, C.POINTER(C.c_uint32)), shape=(graph_like_batches, 2 * n_s))
        i_np = np.ctypeslib.as_array(C.cast(hi, C.POINTER(C.c_uint32)), shape=(graph_like_batches, 3 * n_i))
        d_np = np.ctypeslib.as_array(C.cast(hd, C.POINTER(C.c_uint32)), shape=(graph_like_batches, 3 * n_d))
        want = []
        for b in range(graph_like_batches):                       # batch b: search, delete, insert -- in that order
            fresh = H.random_requests(rng, n_i, loc_base=1000000 + b * n_i)
            sel = np.concatenate([H.to_sel(base), H.to_sel(H.random_requests(rng, 5000))])[rng.integers(0, 65000, n_s)]
            dele = base[b * n_d:(b + 1) * n_d]
            s_np[b] = sel.view(np.uint32); i_np[b] = fresh.view(np.uint32).reshape(-1); d_np[b] = dele.view(np.uint32).reshape(-1)
            want.append(o.search(sel)); o.delete(dele); o.insert(fresh)
        o_np[:] = 0xDEADBEEF
        for b in range(graph_like_batches):                       # one worker stream: batches stay ordered
            N.check(L.gpuhash_index_submit(ix.h, 0, hs + 8 * n_s * b, n_s, ho + 8 * n_s * b, hd + 12 * n_d * b, n_d,
                                           hi + 12 * n_i * b, n_i))
        ix.sync()
        for b in range(graph_like_batches):
            assert np.array_equal(np.sort(o_np[b].reshape(-1, 2), 1), np.sort(want[b].reshape(-1, 2), 1)), f"batch {b}"
        assert o.digest(table=ix.dump()) == o.digest()
    finally:
        L.gpuhash_set_tuning(old)
        for p_ in (hs, ho, hi, hd):
            L.gpuhash_host_free(p_)
        ix.close()


def test_search_py_search_stream_fixture(gpu, layout, rng):
    """libgpuhash/test/back/py_search_stream.c:104-129: hit in both buckets for every query."""
    mem_p = 20
    o = po.Oracle(mem_p, table=H.fixture_every_bucket_1_to_8(mem_p))
    t = table_from_oracle(o, layout)
    sel = np.empty(200000, dtype=mk.SEL_DT)
    sel["hash"] = rng.integers(0, o.num_buckets, len(sel))
    sel["sig"] = rng.integers(1, 9, len(sel))
    got = gpu_search(t, sel)
    assert np.all(got == 1) and np.array_equal(got, o.search(sel))


def test_search_duplicate_signature_behaviour(gpu, layout):
    """key in both buckets -> both words; same signature twice in a bucket -> lowest slot; sig 0 -> empty slots match."""
    o = po.Oracle(16)
    tb = o.buckets()
    sig, h = 0x1234, 5
    b1, b2 = int(o.bucket1(h)), int(o.bucket2(h, sig))
    tb[b1, 0, :2] = [9, sig]; tb[b1, 1, :2] = [90, 111]
    tb[b2, 0, 0] = sig; tb[b2, 1, 0] = 222
    tb[4, 0, :3] = [0x99, 1, 0x99]; tb[4, 1, :3] = [10, 11, 12]
    tb[7, 1, :] = np.arange(50, 58)                                      # stale locs in an empty bucket
    sel = np.array([(sig, h), (sig, b2), (0x99, 4), (0, 7), (0x55 << 16, 4)], dtype=mk.SEL_DT)
    t = table_from_oracle(o, layout)
    want = o.search(sel)
    assert list(want) == [111, 222, 222, 111, 10, 0, 50, 50, 0, 0]
    assert np.array_equal(gpu_search(t, sel, prezero=False), want)


# ----------------------------------------------------------------------------- delete

@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
def test_delete_counts_and_table_bytes_exact(gpu, layout, algo, rng):
    o = po.Oracle(18, algo)
    iel = H.random_requests(rng, 28000)                                  # ~85 % load
    o.insert(iel)
    t = table_from_oracle(o, layout)
    dele = iel[rng.permutation(len(iel))[:9000]].copy()
    dele["loc"][::5] += 1                                                # wrong loc: must not delete
    dele = np.concatenate([dele, H.random_requests(rng, 2000)])          # absent keys
    want = o.delete(dele)
    assert gpu_delete(t, dele) == want
    assert np.array_equal(t.dump_reference(), o.table)                # stale locs included
    assert gpu_delete(t, dele) == o.delete(dele) == 0                    # idempotent


def test_delete_duplicate_requests_in_one_batch(gpu, layout, rng):
    """two identical delete requests: the sequential run zeroes once; with the key in both buckets the second
    request falls through to bucket 2 (gpu_hash.cu:465-468) -- the CAS-based kernel must agree on the count."""
    o = po.Oracle(16)
    sig, h = 0x4321, 17
    b1, b2 = int(o.bucket1(h)), int(o.bucket2(h, sig))
    tb = o.buckets()
    tb[b1, 0, 0] = sig; tb[b1, 1, 0] = 5; tb[b2, 0, 3] = sig; tb[b2, 1, 3] = 5
    iel = H.random_requests(rng, 3000); o.insert(iel)
    t = table_from_oracle(o, layout)
    dele = np.concatenate([iel[:500], iel[:500], np.array([(sig, h, 5)] * 2, dtype=mk.IEL_DT)])
    want = o.delete(dele)
    assert want == 502
    assert gpu_delete(t, dele) == want
    assert np.array_equal(t.dump_reference(), o.table)


# ----------------------------------------------------------------------------- insert, order defined

@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_golden_sequences_serial_mode_slot_exact(gpu, layout, name):
    """committed vectors: every search result, delete count and the final table bytes"""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    mem_p, algo, steps = MG.unpack_steps(g)
    t = mk.DeviceTable(mem_p, algo, layout)
    st = mk.DeviceStats()
    for i, (op, arr) in enumerate(steps):
        if op == MG.OP_INSERT:
            gpu_insert(t, arr, flags=mk.INSERT_SERIAL, stats=st)
        elif op == MG.OP_DELETE:
            assert gpu_delete(t, arr, flags=mk.INSERT_SERIAL) == int(g[f"res{i}"][0]), f"step {i}"
        else:
            assert np.array_equal(gpu_search(t, arr, prezero=False), g[f"res{i}"]), f"step {i}"
    assert np.array_equal(t.dump_reference(), g["table"])
    s = st.read()
    got = [s[k] for k in ("ins_skipped", "ins_updated", "ins_placed_b1", "ins_placed_b2", "ins_to_b2",
                          "ins_displaced", "ins_dropped", "ins_overwritten")]
    assert got == g["stats"].tolist()
    assert s["ins_gave_up"] == 0 and s["ins_cas_retry"] == 0


@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
@pytest.mark.parametrize("load", [0.5, 0.9, 1.05])
def test_insert_serial_mode_slot_exact_at_any_load(gpu, layout, algo, load, rng):
    mem_p = 17
    o = po.Oracle(mem_p, algo)
    iel = H.random_requests(rng, int(load * (1 << mem_p) / 8))
    t = mk.DeviceTable(mem_p, algo, layout)
    st = mk.DeviceStats()
    for part in np.array_split(iel, 3):
        o.insert(part)
        gpu_insert(t, part, flags=mk.INSERT_SERIAL, stats=st)
    assert np.array_equal(t.dump_reference(), o.table)
    s, w = st.read(), o.stats.as_dict()
    assert (s["ins_to_b2"], s["ins_displaced"], s["ins_dropped"], s["ins_overwritten"]) == \
           (w["to_b2"], w["displaced"], w["dropped"], w["overwritten"])
    if algo == po.CUCKOO:
        assert s["chain_hist"] == w["chain_hist"]
    sel = H.to_sel(iel)
    assert np.array_equal(gpu_search(t, sel), o.search(sel))


def test_insert_serial_segments_in_block_order(gpu, layout, rng):
    """legacy layout: 8 segments with device-side counts, some empty (mega_scheduler.c:486-489)"""
    mem_p = 16
    o = po.Oracle(mem_p)
    iel = H.random_requests(rng, 9000)                                   # > 100 % load
    blocks = mk.split_insert_blocks(iel, 8)
    blocks[2] = blocks[2][:0]; blocks[7] = blocks[7][:0]
    o.insert_blocks(blocks)
    t = mk.DeviceTable(mem_p, layout=layout)
    segs = mk.InsertSegments(blocks)
    mk.insert_ex(t.geom, t, segs, flags=mk.INSERT_SERIAL)
    mk.device_sync()
    assert np.array_equal(t.dump_reference(), o.table)


@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
def test_insert_concurrent_conflict_free_batches_slot_exact(gpu, layout, algo, rng):
    """no two requests of a batch share a candidate bucket => order cannot matter => identical bytes"""
    mem_p = 22
    o = po.Oracle(mem_p, algo)
    t = mk.DeviceTable(mem_p, algo, layout)
    total = 0
    for _ in range(12):
        batch = H.conflict_free(o, H.random_requests(rng, 6000, loc_base=total + 1))
        room = (o.buckets()[o.bucket1(batch["hash"]), 0, :] == 0).any(axis=1)     # bucket 1 still has a free slot
        batch = batch[room]
        total += len(batch)
        before = o.stats.to_b2
        o.insert(batch)
        assert o.stats.to_b2 == before                                   # stays inside its own buckets
        gpu_insert(t, batch)
        assert np.array_equal(t.dump_reference(), o.table)
    assert total > 40000


# ----------------------------------------------------------------------------- insert, concurrent

def test_insert_concurrent_multiset_exact_below_half_load(gpu, layout, rng):
    """cuckoo, unique keys, load 0.4: some buckets overflow and a few victims are re-homed, but nothing is
    dropped in any order, so the set of stored (sig, loc) pairs is order-independent; so is every search result
    as a set {o0, o1}."""
    mem_p, algo = 22, po.CUCKOO
    o = po.Oracle(mem_p, algo)
    t = mk.DeviceTable(mem_p, algo, layout)
    iel = H.random_requests(rng, int(0.4 * (1 << mem_p) / 8))
    st = mk.DeviceStats()
    for part in np.array_split(iel, 4):
        o.insert(part)
        gpu_insert(t, part, stats=st)
    assert o.stats.dropped == 0 and o.stats.updated == 0 and o.stats.to_b2 > 100
    got = t.dump_reference()
    s = st.read()
    if layout == mk.LAYOUT_PAIRS:
        assert o.digest(table=got) == o.digest()
        assert np.array_equal(H.occupied_pairs(o.buckets(got)), H.occupied_pairs(o.buckets()))
    else:
        # reference byte layout: sig and loc are published by two separate operations (as in the reference), so
        # an eviction that lands on a slot claimed microseconds earlier can pair a signature with a stale
        # location.  Bounded here; the pair layout above has no such window.
        a, b = H.occupied_pairs(o.buckets(got)), H.occupied_pairs(o.buckets())
        assert len(a) == len(b) and len(np.setdiff1d(a, b)) <= 4 * max(s["ins_displaced"], 1)
    assert s["ins_gave_up"] == 0 and s["ins_dropped"] == 0 and s["ins_updated"] == 0
    # a request that ends by evicting is not counted as placed, its victim is when it lands: the sum stays N
    assert s["ins_placed_b1"] + s["ins_placed_b2"] == len(iel)
    sel = H.to_sel(iel)
    g, w = gpu_search(t, sel).reshape(-1, 2), o.search(sel).reshape(-1, 2)
    assert int((g != 0).any(axis=1).sum()) > 0.99 * len(iel)
    # a key whose bucket 1 overflowed may sit in b1 or b2 depending on the order; orphaned victims depend on who
    # evicted them.  Everything else is identical word for word.
    same = (np.sort(g, axis=1) == np.sort(w, axis=1)).all(axis=1)
    assert (~same).sum() <= 4 * (o.stats.displaced + s["ins_displaced"]) + 8
    if layout == mk.LAYOUT_PAIRS:
        assert np.all((g == 0) | (g == iel["loc"][:, None]))


def test_insert_legacy_abi_then_search_then_delete(gpu, layout, rng):
    """the reference's own test, through the three legacy symbols with the reference's launch arguments
    (insert_test.c:145,173,207): inserted => found, deleted => gone."""
    mem_p, n = 22, 16384
    t = mk.DeviceTable(mem_p, layout=layout); t.make_default()
    o = po.Oracle(mem_p)
    nb = o.num_buckets
    out_d = mk.DeviceBuffer(8 * n)
    for it in range(6):
        iel = np.empty(n, dtype=mk.IEL_DT)
        iel["sig"] = rng.integers(1, 2**31, n)
        iel["hash"] = (np.arange(n) // (n // 8)) * (nb // 8) + rng.integers(0, nb // 8, n)
        iel["loc"] = rng.integers(1, 2**31, n)
        in_d = mk.DeviceBuffer.from_host(iel)
        ptrs = mk.DeviceBuffer.from_host(np.array([in_d.ptr + 12 * k * (n // 8) for k in range(8)], dtype=np.uint64))
        nums = mk.DeviceBuffer.from_host(np.full(8, n // 8, dtype=np.int32))
        mk.gpu_hash_insert(t, ptrs, nums, 8)
        mk.device_sync()
        sel_d = mk.DeviceBuffer.from_host(H.to_sel(iel))
        out_d.zero()
        mk.gpu_hash_search(sel_d, out_d, t, n, 16384, 128)
        mk.device_sync()
        out = out_d.download(np.uint32)
        assert np.all((out[0::2] == iel["loc"]) | (out[1::2] == iel["loc"]))
        o.insert(iel)
        assert o.digest(table=t.dump_reference()) == o.digest()
        if it % 2:
            mk.gpu_hash_delete(in_d, t, n, 16384, 128)
            mk.device_sync()
            out_d.zero()
            mk.gpu_hash_search(sel_d, out_d, t, n, 16384, 128)
            mk.device_sync()
            out = out_d.download(np.uint32)
            assert not np.any((out[0::2] == iel["loc"]) | (out[1::2] == iel["loc"]))
            o.delete(iel)
            assert o.digest(table=t.dump_reference()) == o.digest()


@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
def test_insert_concurrent_high_load_invariants(gpu, layout, algo, rng):
    """90 % load + churn, unconstrained batches: slot placement and which victim is dropped depend on the
    interleaving, so compare what cannot: pair integrity, conservation, and the oracle's statistics."""
    mem_p = 20
    slots = (1 << mem_p) // 8
    iel = H.random_requests(rng, int(0.9 * slots))
    o = po.Oracle(mem_p, algo); o.insert(iel)
    t = mk.DeviceTable(mem_p, algo, layout)
    st = mk.DeviceStats()
    for part in np.array_split(iel, 8):
        gpu_insert(t, part, stats=st)
    s, w = st.read(), o.stats.as_dict()
    got = o.buckets(t.dump_reference())
    pairs = H.occupied_pairs(got)
    legal = np.sort((iel["sig"].astype(np.uint64) << np.uint64(32)) | iel["loc"].astype(np.uint64))
    if algo == po.CUCKOO:
        torn = int((~np.isin(pairs, legal)).sum())
        if layout == mk.LAYOUT_PAIRS:
            assert torn == 0                                             # 64-bit CAS: a pair can never tear
            assert len(np.unique(pairs)) == len(pairs)                   # nothing duplicated
        else:
            assert torn <= 0.01 * len(pairs)                             # reference layout: its two-step publication window
        assert len(pairs) == len(iel) - s["ins_dropped"]                 # conservation
        # all requests of a launch run at once, so chains collide more often than in a sequential run
        assert 0.5 * w["dropped"] - 50 <= s["ins_dropped"] <= 3 * w["dropped"] + 100
        assert abs(s["ins_displaced"] - w["displaced"]) <= 0.25 * w["displaced"] + 50
    else:
        # 2-choice overwrites keep the old loc: the signature must be legal, the pair need not be
        assert np.all(np.isin(got[:, 0, :][got[:, 0, :] != 0], iel["sig"]))
        assert len(pairs) == len(iel) - s["ins_overwritten"]
        assert abs(s["ins_overwritten"] - w["overwritten"]) <= 0.1 * w["overwritten"] + 50
    assert s["ins_gave_up"] == 0
    assert abs(s["ins_to_b2"] - w["to_b2"]) <= 0.15 * w["to_b2"] + 50
    sel = H.to_sel(iel)
    found_gpu = (gpu_search(t, sel).reshape(-1, 2) == iel["loc"][:, None]).any(axis=1).sum()
    found_orc = (o.search(sel).reshape(-1, 2) == iel["loc"][:, None]).any(axis=1).sum()
    assert abs(int(found_gpu) - int(found_orc)) <= 0.03 * len(iel)


def test_insert_same_bucket_contention_keeps_pairs_intact(gpu, layout, rng):
    """adversarial: 40 000 requests aimed at 64 buckets (and their alternates) in one launch.  Almost
    everything is evicted or dropped; what survives must be real pairs and no slot may be claimed twice."""
    mem_p = 16
    t = mk.DeviceTable(mem_p, layout=layout)
    iel = H.random_requests(rng, 40000)
    iel["hash"] = (iel["hash"] & np.uint32(63))
    st = mk.DeviceStats()
    gpu_insert(t, iel, stats=st)
    o = po.Oracle(mem_p)
    got = o.buckets(t.dump_reference())
    pairs = H.occupied_pairs(got)
    legal = np.sort((iel["sig"].astype(np.uint64) << np.uint64(32)) | iel["loc"].astype(np.uint64))
    s = st.read()
    assert s["ins_gave_up"] == 0
    assert len(pairs) == len(iel) - s["ins_dropped"]                     # every request is stored or counted as dropped
    if layout == mk.LAYOUT_PAIRS:
        assert np.all(np.isin(pairs, legal))
        assert len(np.unique(pairs)) == len(pairs)
    else:
        assert np.isin(o.buckets(t.dump_reference())[:, 0, :].reshape(-1), np.concatenate([iel["sig"], [0]])).all()


def test_insert_duplicate_keys_in_one_batch(gpu, layout, rng):
    """zipf-like SET traffic: the same key many times in one launch.  One slot per key, loc = one of the
    batch's locs for that key (the sequential run keeps the last; a concurrent run keeps some request's)."""
    mem_p = 18
    base = H.random_requests(rng, 2000)
    rep = base[rng.integers(0, 50, 60000)].copy()                        # 50 hot keys
    rep["loc"] = np.arange(1, len(rep) + 1)
    batch = np.concatenate([base[50:], rep]); rng.shuffle(batch)
    t = mk.DeviceTable(mem_p, layout=layout)
    gpu_insert(t, batch)
    o = po.Oracle(mem_p); o.insert(batch)
    got = o.buckets(t.dump_reference())
    assert np.array_equal(np.sort(got[:, 0, :], axis=1), np.sort(o.buckets()[:, 0, :], axis=1))   # same sigs per bucket
    out = gpu_search(t, H.to_sel(base[:50])).reshape(-1, 2)
    for k in range(50):
        locs = rep["loc"][(rep["sig"] == base["sig"][k]) & (rep["hash"] == base["hash"][k])]
        assert out[k, 0] in locs and out[k, 1] in (0, out[k, 0])    # o1 == o0 iff alt bucket == bucket 1


# ----------------------------------------------------------------------------- scheduler-cycle object

@pytest.mark.parametrize("fused", [1, 0])
def test_index_cycle_matches_oracle_in_reference_order(gpu, layout, fused, rng):
    """search -> delete -> insert inside one cycle (mega_scheduler.c:392-502): searches of a cycle do not
    see that cycle's inserts."""
    mem_p = 20
    old = N.Tune(); N.lib().gpuhash_get_tuning(old)
    N.lib().gpuhash_set_tuning(N.Tune(0, 0, 4, fused))              # one launch per cycle, or one per operation kind
    ix = mk.GpuHashIndex(mem_p, workers=2, max_search=1 << 16, max_insert=1 << 15, max_delete=1 << 15, layout=layout)
    o = po.Oracle(mem_p)
    live = H.random_requests(rng, 30000)
    ix.insert(live[:15000]); ix.insert(live[15000:])
    o.insert(live)
    loc0 = len(live) + 1
    for cyc in range(5):
        fresh = H.random_requests(rng, 3000, loc_base=loc0); loc0 += 3000
        dele = live[rng.permutation(len(live))[:2000]]
        sel = np.concatenate([H.to_sel(live[::2]), H.to_sel(fresh)])
        want = o.search(sel)
        o.delete(dele); o.insert(fresh)
        got = ix.cycle(search=sel, delete=dele, insert=fresh, worker=cyc % 2)
        assert np.array_equal(np.sort(got.reshape(-1, 2), axis=1), np.sort(want.reshape(-1, 2), axis=1))
        assert not got.reshape(-1, 2)[-3000:].any()                      # this cycle's inserts are not visible yet
        assert o.digest(table=ix.dump()) == o.digest()
        live = np.concatenate([live[~np.isin(live["loc"], dele["loc"])], fresh])
    ix.close()
    N.lib().gpuhash_set_tuning(old)


def test_legacy_gpu_delete_insert_is_one_launch_with_both_effects(gpu, layout, rng):
    """libgpuhash.h:53-62 declares gpu_delete_insert; the reference never defines it.  Here: delete batch, then the
    8-segment insert batch (device-side counts), in one launch -- equal to gpu_hash_delete + gpu_hash_insert."""
    import ctypes as C
    mem_p = 20
    t = mk.DeviceTable(mem_p, layout=layout); t.make_default()
    o = po.Oracle(mem_p)
    base = H.random_requests(rng, 40000)
    gpu_insert(t, base); o.insert(base)
    dele = base[rng.permutation(len(base))[:7000]]
    fresh = H.random_requests(rng, 9000, loc_base=50001)
    fresh[:100] = dele[:100]; fresh["loc"][:100] += 7                 # re-insert of keys deleted in the same call
    blocks = mk.split_insert_blocks(fresh, 8); blocks[3] = blocks[3][:0]
    segs = mk.InsertSegments(blocks)
    del_d = mk.DeviceBuffer.from_host(dele)
    N.lib().gpu_delete_insert(t.ptr, del_d.ptr, len(dele), segs.ptrs.ptr, segs.nums.ptr, 8, 16384, 256, None)
    mk.device_sync()
    o.delete(dele); o.insert_blocks(blocks)
    assert o.digest(table=t.dump_reference()) == o.digest()
    sel = H.to_sel(np.concatenate([base, fresh]))
    g_, w_ = np.sort(gpu_search(t, sel).reshape(-1, 2), 1), np.sort(o.search(sel).reshape(-1, 2), 1)
    bad = np.nonzero((g_ != w_).any(axis=1))[0]
    # which key a full bucket pushes out (and orphans, gpu_hash.cu:334-335) depends on the order of the requests
    assert len(bad) <= 4 * o.stats.displaced + 4, f"{len(bad)} rows differ, e.g. rows {bad[:5]}: gpu {g_[bad[:5]].tolist()} oracle {w_[bad[:5]].tolist()}"


# ----------------------------------------------------------------------------- full size (BASELINE config 2)

def test_full_size_table_properties(gpu, layout, rng):
    """MEM_P 34 (16 GiB, 2^28 buckets): sizes the oracle cannot hold on the host, so use properties:
    inserted => found at its loc (both words otherwise 0 or the loc), absent => miss, deleted => gone,
    and the stats counters balance."""
    mem_p = 34
    free = np.zeros(1, dtype=np.uint64); total = np.zeros(1, dtype=np.uint64)
    import ctypes as C
    N.check(N.lib().gpuhash_device_info(0, None, None, C.cast(free.ctypes.data, C.POINTER(C.c_size_t)),
                                        C.cast(total.ctypes.data, C.POINTER(C.c_size_t))))
    if int(free[0]) < (20 << 30):
        pytest.skip("needs 20 GiB of free device memory")
    t = mk.DeviceTable(mem_p, layout=layout)
    from megakv_b200 import keystream as ks
    n = 1 << 24                                                          # 16 M keys
    st = mk.DeviceStats()
    iel, sel = ks.uniform_inserts(7, 0, n)
    gpu_insert(t, iel, stats=st)
    s = st.read()
    assert s["ins_placed_b1"] + s["ins_placed_b2"] + s["ins_updated"] == n and s["ins_gave_up"] == 0
    out = gpu_search(t, sel, prezero=False).reshape(-1, 2)
    assert np.all((out == iel["loc"][:, None]).any(axis=1))
    assert np.all((out == 0) | (out == iel["loc"][:, None]))
    _, absent = ks.uniform_inserts(8, 0, 1 << 20)
    assert not gpu_search(t, absent, prezero=False).any()
    assert gpu_delete(t, iel[: 1 << 20]) == (1 << 20)
    out = gpu_search(t, sel[: 1 << 21], prezero=False).reshape(-1, 2)
    assert not out[: 1 << 20].any() and np.all((out[1 << 20:] == iel["loc"][1 << 20: 1 << 21, None]).any(axis=1))
    # top-of-table buckets are reachable (64-bit addressing): hash = HASH_MASK
    hi = np.array([(0x7777, 0x0FFFFFFF, 99), (0x7778, 0xFFFFFFFF, 98)], dtype=mk.IEL_DT)
    gpu_insert(t, hi)
    assert list(gpu_search(t, H.to_sel(hi))[[0, 2]]) == [99, 98]


# ----------------------------------------------------------------------------- the steps either side of the path

@pytest.mark.parametrize("in_off,out_off", [(0, 0), (1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 4097, 62259, 300001])
def test_search_compact_is_the_senders_choice(gpu, layout, n, in_off, out_off, rng):
    """gpuhash_search_compact_ex: one word per request = search_out[2i] if non-zero else search_out[2i+1]
    (src/mega_send.c:411-414), including keys that sit in their alternate bucket and keys present in both."""
    o = po.Oracle(16)
    iel = H.random_requests(rng, 7000)                           # load 0.85: many keys in bucket 2, some orphans
    o.insert(iel)
    t = table_from_oracle(o, layout)
    sel = np.concatenate([H.to_sel(iel), H.to_sel(H.random_requests(rng, 2000))])[rng.integers(0, 9000, n)]
    in_d = mk.DeviceBuffer(8 * (n + 4)); out_d = mk.DeviceBuffer(4 * (n + 8))
    pad = np.zeros(n + 4, dtype=mk.SEL_DT); pad[in_off:in_off + n] = sel
    in_d.upload(pad)
    out_d.upload(np.full(n + 8, 0xDEADBEEF, dtype=np.uint32))
    N.check(N.lib().gpuhash_search_compact_ex(C.byref(t.geom), in_d.ptr + 8 * in_off, out_d.ptr + 4 * out_off, t.ptr, n, None, None))
    mk.device_sync()
    got = out_d.download(np.uint32)
    want2 = o.search(sel)
    assert (want2.reshape(-1, 2)[:, 0] == 0).any() or n < 64     # the case exercises bucket-2 hits
    assert np.array_equal(got[out_off:out_off + n], po.compact_results(want2))
    assert np.all(got[:out_off] == 0xDEADBEEF) and np.all(got[out_off + n:] == 0xDEADBEEF)


@pytest.mark.parametrize("fold", [0, 1])
@pytest.mark.parametrize("nkey,stride", [(8, 8), (16, 16), (20, 24), (31, 31), (128, 130)])
def test_fold_keys_matches_receiver_derivation(gpu, nkey, stride, fold, rng):
    """gpuhash_fold_keys_ex against the restatement of src/mega_recv.c:349-362 (plain and -DSIGNATURE builds)"""
    n = 50001
    raw = rng.integers(0, 256, size=(n, stride), dtype=np.uint8)
    keys_d = mk.DeviceBuffer.from_host(raw)
    out_d = mk.DeviceBuffer(8 * n)
    N.check(N.lib().gpuhash_fold_keys_ex(keys_d.ptr, stride, nkey, fold, n, out_d.ptr, None))
    mk.device_sync()
    got = out_d.download(np.uint32).view(mk.SEL_DT)
    want = po.fold_keys(raw[:, :nkey], bool(fold))
    assert np.array_equal(got["sig"], want["sig"]) and np.array_equal(got["hash"], want["hash"])


@pytest.mark.parametrize("mem_p", [26, 28], ids=["L2-resident", "beyond-L2"])
def test_config2_zipf_mixed_two_choice(gpu, layout, mem_p, rng):
    """BASELINE configs[2]: zipf(0.99) mixed insert/search, HASH_2CHOICE, a table that fits in L2 (64 MiB) and one that
    does not (256 MiB).  Rounds of {search batch, then insert batch} through the scheduler-cycle call; hot keys repeat
    many times inside one batch, so every SET of a key in a round carries the same new location (the outcome is then
    independent of which duplicate wins) and every search word of every round must equal the oracle's."""
    from megakv_b200 import keystream as ks
    pop = (1 << mem_p) // 8 // 4                                   # load factor 0.25
    o = po.Oracle(mem_p, po.TWO_CHOICE)
    ix = mk.GpuHashIndex(mem_p, algo=po.TWO_CHOICE, workers=1, max_search=1 << 17, max_insert=1 << 17, max_delete=8, layout=layout)
    try:
        chunk = 1 << 20
        for first in range(0, pop, chunk):
            iel, _ = ks.uniform_inserts(3, first, min(chunk, pop - first))
            o.insert_mt(iel, 8)
        ix.load(o.table)
        z = ks.Zipf(pop, 0.99, rng)
        for rnd in range(1, 6):
            sel = ks.keys_to_requests(ks._keys_at(3, z.ranks(60000)))
            idx = z.ranks(20000)
            fresh_idx = pop + (rnd - 1) * 3000 + np.arange(3000)   # 13 % of the SETs are new keys
            allidx = np.concatenate([idx, fresh_idx])
            rng.shuffle(allidx)
            iel, _ = ks.keys_to_requests(ks._keys_at(3, allidx), (allidx + 1 + rnd * (1 << 27)).astype(np.uint64))
            want = o.search(sel)
            got = ix.cycle(search=sel, insert=iel)
            o.insert(iel)
            assert np.array_equal(got, want), f"round {rnd}"
            assert (want.reshape(-1, 2) != 0).any(axis=1).mean() > 0.99
        assert o.digest(table=ix.dump()) == o.digest()
    finally:
        ix.close()


@pytest.mark.parametrize("zero_copy", [0, 1])
def test_index_compact_results_mode(gpu, layout, zero_copy, rng):
    """gpuhash_index_set_compact_results: search_out_h gets ONE word per request (mega_send.c:411-414's choice), through
    staging copies and through zero-copy, with deletes and inserts of the same cycle still applied in order."""
    L = N.lib()
    mem_p = 22
    o = po.Oracle(mem_p)
    ix = mk.GpuHashIndex(mem_p, workers=1, max_search=1 << 17, max_insert=1 << 16, max_delete=1 << 16, layout=layout)
    base = H.random_requests(rng, 60000)
    ix.insert(base); o.insert(base)
    L.gpuhash_index_set_zero_copy(ix.h, zero_copy); L.gpuhash_index_set_compact_results(ix.h, 1)
    n_s, n_i, n_d = 62259, 3277, 500
    hs = L.gpuhash_host_alloc(8 * n_s); ho = L.gpuhash_host_alloc(4 * n_s + 16)
    hi = L.gpuhash_host_alloc(12 * n_i); hd = L.gpuhash_host_alloc(12 * n_d)
    try:
        s_np = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint32)), shape=(2 * n_s,))
        o_np = np.ctypeslib.as_array(C.cast(ho, C.POINTER(C.c_uint32)), shape=(n_s + 4,))
        i_np = np.ctypeslib.as_array(C.cast(hi, C.POINTER(C.c_uint32)), shape=(3 * n_i,))
        d_np = np.ctypeslib.as_array(C.cast(hd, C.POINTER(C.c_uint32)), shape=(3 * n_d,))
        for rnd in range(2):
            fresh = H.random_requests(rng, n_i, loc_base=2000000 + rnd * n_i)
            sel = np.concatenate([H.to_sel(base), H.to_sel(H.random_requests(rng, 5000))])[rng.integers(0, 65000, n_s)]
            dele = base[rnd * n_d:(rnd + 1) * n_d]
            s_np[:] = sel.view(np.uint32); i_np[:] = fresh.view(np.uint32).reshape(-1); d_np[:] = dele.view(np.uint32).reshape(-1)
            o_np[:] = 0xDEADBEEF
            want = po.compact_results(o.search(sel)); o.delete(dele); o.insert(fresh)
            N.check(L.gpuhash_index_submit(ix.h, 0, hs, n_s, ho, hd, n_d, hi, n_i))
            ix.sync()
            assert np.array_equal(o_np[:n_s], want), f"round {rnd}"
            assert np.all(o_np[n_s:] == 0xDEADBEEF)
        L.gpuhash_index_set_compact_results(ix.h, 0)
        assert o.digest(table=ix.dump()) == o.digest()
    finally:
        for p_ in (hs, ho, hi, hd):
            L.gpuhash_host_free(p_)
        ix.close()


@pytest.mark.parametrize("mode", ["legacy", "submit", "ring"])
def test_c_host_example_runs_all_three_integration_levels(gpu, mode, tmp_path):
    """examples/scheduler_cycle.c on the GPU: the reference's own launch loop against the legacy ABI, the one-call-per-
    worker cycle, and the launch-free ring; every GET must come back with the location that was SET (the property
    libgpuhash/test/insert_test.c:178-195 checks).  The program exits non-zero on any wrong result."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "scheduler_cycle"
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(root, "examples", "scheduler_cycle.c"), os.path.join(root, "megakv_b200", "lib", "libgpuhash.a"),
                           "-L", cuda_lib, "-lcudart", "-lrt", "-lpthread", "-ldl", "-o", str(exe)])
    env = dict(os.environ); env["LD_LIBRARY_PATH"] = cuda_lib + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([str(exe), mode, "4", "8", "26"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 wrong results" in r.stdout
