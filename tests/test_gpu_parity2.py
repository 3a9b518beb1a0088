"""GPU parity, second file (round 2): the cases VERDICT r01 asked for on top of tests/test_gpu_parity.py.

  * the full-size geometry (MEM_P 34, 25-bit BLOCK_HASH_MASK) word for word against the oracle -- and against the
    reference's own search kernel where oracle/_ref holds its MEM_P 34 build;
  * the key derivation of the reference's LOCAL_TEST load generator (src/mega_recv.c:700-704), whose alternate bucket
    is degenerate: every key of a block shares ONE alternate bucket;
  * a search launch racing an insert launch on another stream of the same table (mega_scheduler.c runs one stream per
    worker with no order between them);
  * concurrent inserts at 90 % load against the envelope of sequential oracle runs over random permutations of the
    same requests (instead of the wide 0.5x .. 3x bands of test_insert_concurrent_high_load_invariants).
"""
import ctypes as C
import os

import numpy as np
import pytest

import megakv_b200 as mk
from megakv_b200 import _native as N
from oracle import pyoracle as po
from tests import helpers as H
from tests.test_gpu_parity import gpu_search, gpu_insert, gpu_delete

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mem_available():
    try:
        return int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        return 0


def test_full_size_table_word_for_word(gpu, rng):
    """MEM_P 34 (2^28 buckets, BLOCK_HASH_MASK 25 bits): 200 000 keys with full 32-bit hashes -- so the top buckets and
    the top bits of the alternate-bucket mask are exercised -- inserted, searched (hits, misses), a third deleted,
    searched again: every result word equals the oracle's.  The oracle's 16 GiB host table is calloc'ed: only the pages
    that hold keys are ever touched."""
    L = N.lib()
    free_, total_ = C.c_size_t(), C.c_size_t()
    N.check(L.gpuhash_device_info(0, None, None, C.byref(free_), C.byref(total_)))
    if free_.value < (20 << 30) or _mem_available() < (6 << 30):
        pytest.skip("needs 20 GiB of device memory and 6 GiB of host memory")
    mem_p = 34
    keys = H.random_requests(rng, 200000)
    keys["hash"][:64] |= np.uint32(0xFFFFFFC0)                       # the very last buckets of the table
    keys["hash"][64:128] &= np.uint32(0x3F)                          # and the very first
    o = po.Oracle(mem_p); o.insert(keys)
    t = mk.DeviceTable(mem_p)
    gpu_insert(t, keys)
    miss = H.random_requests(rng, 50000, loc_base=10**7)
    probe = np.concatenate([H.to_sel(keys), H.to_sel(miss)])
    want = o.search(probe)
    got = gpu_search(t, probe, prezero=False)
    assert np.array_equal(got, want)
    assert ((want.reshape(-1, 2)[:200000] == keys["loc"][:, None]).any(axis=1)).all()
    # the reference's own kernel on the same table bytes (its MEM_P 34 build, when this box has it)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libgpuhash_ref_cuckoo_34.so")
    if os.path.exists(ref_so):
        R = C.CDLL(ref_so)
        R.gpu_hash_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        R.gpu_hash_search.restype = None
        in_d = mk.DeviceBuffer.from_host(probe); out_d = mk.DeviceBuffer(8 * len(probe), zero=True)
        as_ref = N.Geom.from_buffer_copy(bytes(t.geom)); as_ref.layout = N.LAYOUT_REFERENCE
        N.check(L.gpuhash_table_convert(C.byref(t.geom), t.ptr, N.LAYOUT_REFERENCE, None)); mk.device_sync()
        try:
            R.gpu_hash_search(in_d.ptr, out_d.ptr, t.ptr, len(probe), 24576, 256, None); mk.device_sync()
            assert np.array_equal(out_d.download(np.uint32), want), "the reference's kernel disagrees with its restatement"
        finally:
            N.check(L.gpuhash_table_convert(C.byref(as_ref), t.ptr, t.geom.layout, None)); mk.device_sync()
    zeroed = gpu_delete(t, keys[::3])
    assert zeroed == o.delete(keys[::3]) == len(keys[::3])
    assert np.array_equal(gpu_search(t, probe, prezero=False), o.search(probe))
    t.free()


def local_test_keys(first, n, bits_insert_buf=3):
    """src/mega_recv.c:690-704 (LOCAL_TEST load generator): the 8 key bytes are {k, (bswap32(k & 0xff) << (8 - bits)) | k};
    the receiver takes sig = low word, hash = high word (mega_recv.c:350,361-362).  loc = k."""
    k = np.arange(first, first + n, dtype=np.uint64).astype(np.uint32)
    iel = np.empty(n, dtype=mk.IEL_DT)
    iel["sig"] = k
    iel["hash"] = (((k & np.uint32(0xFF)) << np.uint32(24)) << np.uint32(8 - bits_insert_buf)) | k
    iel["loc"] = k
    return iel


@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
@pytest.mark.parametrize("layout", [mk.LAYOUT_PAIRS, mk.LAYOUT_REFERENCE], ids=["pairs", "reflayout"])
def test_local_test_key_formula(gpu, layout, algo):
    """hash = ((k & 7) << 29) | k, sig = k: hash ^ sig has no low bits, so the alternate bucket of EVERY key of a block is
    the block's bucket 0 -- once first buckets fill, all overflow of a block meets in one bucket.  Serial mode: table bytes
    and search words equal the oracle's; concurrent mode: counters and search results as sets."""
    mem_p = 16
    nb = 1 << (mem_p - 6)
    keys = local_test_keys(1, 10 * nb)                               # ten keys per first bucket: two overflow each
    o = po.Oracle(mem_p, algo)
    b2 = o.bucket2(keys["hash"], keys["sig"])
    assert len(np.unique(b2)) <= 8 and not (b2 & np.uint32((nb >> 3) - 1)).any()      # the degenerate alternate
    o.insert(keys)
    t = mk.DeviceTable(mem_p, algo, layout)
    gpu_insert(t, keys, flags=mk.INSERT_SERIAL)
    assert np.array_equal(t.dump_reference(), o.table)
    sel = np.concatenate([H.to_sel(keys), H.to_sel(local_test_keys(10 * nb + 1, 500))])
    assert np.array_equal(gpu_search(t, sel, prezero=False), o.search(sel))
    # the same requests as one concurrent launch
    t2 = mk.DeviceTable(mem_p, algo, layout)
    st = mk.DeviceStats()
    gpu_insert(t2, keys, stats=st)
    s, w = st.read(), o.stats.as_dict()
    assert s["ins_gave_up"] == 0
    assert abs(s["ins_to_b2"] - w["to_b2"]) <= 0.05 * w["to_b2"] + 20    # (re-homed victims land in first buckets: mildly order-dependent)
    got = o.buckets(t2.dump_reference())
    legal_sig = np.isin(got[:, 0, :][got[:, 0, :] != 0], keys["sig"])
    assert legal_sig.all()
    if algo == po.CUCKOO and layout == mk.LAYOUT_PAIRS:
        pairs = H.occupied_pairs(got)
        legal = (keys["sig"].astype(np.uint64) << np.uint64(32)) | keys["loc"].astype(np.uint64)
        assert np.isin(pairs, legal).all() and len(np.unique(pairs)) == len(pairs)
        assert len(pairs) == len(keys) - s["ins_dropped"]


def test_search_races_insert_on_another_stream(gpu, rng):
    """Two streams on one table, no order between them (the reference's one stream per worker): searches for keys that were
    there before must always find them (load factor 0.2: nothing is ever evicted), searches for the keys being inserted
    return 0 or the key's own location -- never anything else."""
    L = N.lib()
    mem_p = 24
    t = mk.DeviceTable(mem_p)
    old = H.random_requests(rng, 200000)
    gpu_insert(t, old)
    new = H.random_requests(rng, 200000, loc_base=10**6)
    s1, s2 = L.gpuhash_stream_create(), L.gpuhash_stream_create()
    probe = np.concatenate([H.to_sel(old), H.to_sel(new)])
    expect = np.concatenate([old["loc"], new["loc"]])
    in_d = mk.DeviceBuffer.from_host(probe)
    new_d = mk.DeviceBuffer.from_host(new)
    rounds, parts = 12, 25                                           # 25 x 8000 = every new key; the last part after the last search
    outs = [mk.DeviceBuffer(8 * len(probe)) for _ in range(rounds)]
    step = len(new) // parts
    for r in range(rounds):                                          # searches keep coming while the inserts trickle in
        N.check(L.gpuhash_search_ex(C.byref(t.geom), in_d.ptr, outs[r].ptr, t.ptr, len(probe), None, s1))
        for p in (2 * r, 2 * r + 1):
            N.check(L.gpuhash_insert_flat_ex(C.byref(t.geom), t.ptr, new_d.ptr + 12 * step * p, step, None, 0, s2))
    N.check(L.gpuhash_insert_flat_ex(C.byref(t.geom), t.ptr, new_d.ptr + 12 * step * 24, len(new) - 24 * step, None, 0, s2))
    N.check(L.gpuhash_stream_sync(s1)); N.check(L.gpuhash_stream_sync(s2))
    seen_partial = False
    for r in range(rounds):
        got = outs[r].download(np.uint32).reshape(-1, 2)
        assert ((got == expect[:, None]) | (got == 0)).all(), "a search returned a location that is not its key's"
        assert (got[:len(old)] == old["loc"][:, None]).any(axis=1).all(), "a key that was there before the race was missed"
        found_new = (got[len(old):] == new["loc"][:, None]).any(axis=1).mean()
        seen_partial |= 0.0 < found_new < 1.0
    final = gpu_search(t, probe).reshape(-1, 2)
    assert (final == expect[:, None]).any(axis=1).all()
    L.gpuhash_stream_destroy(s1); L.gpuhash_stream_destroy(s2)


def test_concurrent_high_load_within_the_envelope_of_sequential_orders(gpu, rng):
    """90 % load, cuckoo, pair layout: the counters of concurrent launches against sequential oracle runs of the same requests
    in six random orders (measured first: tools/exp_concurrent_envelope.py, profiles/r02_concurrent_envelope.md).
      * 512 launches of ~1 800 requests: the run is nearly sequential and has to land INSIDE the envelope of the sequential
        orders, widened by the envelope's own width + 1.5 % of the mean (+ 20);
      * 8 launches of ~118 000 requests: every request of a launch tries its first bucket before the evictions of that
        launch have re-homed any victim into it, so fewer requests find bucket 1 full (to_b2 6-8 % lower), more of them
        meet in the alternates (displaced 5-7 % higher) and more chains reach MAX_CUCKOO_NUM (dropped 1.3-1.7 x): a
        systematic, one-directional shift, asserted as such (round 1 allowed 0.5 x .. 3 x either way)."""
    mem_p = 20
    slots = (1 << mem_p) // 8
    iel = H.random_requests(rng, int(0.9 * slots))
    env = {"to_b2": [], "displaced": [], "dropped": []}
    for k in range(6):
        o = po.Oracle(mem_p, po.CUCKOO)
        o.insert(iel[rng.permutation(len(iel))])
        w = o.stats.as_dict()
        for key in env:
            env[key].append(w[key])

    def concurrent(nparts):
        t = mk.DeviceTable(mem_p, po.CUCKOO, mk.LAYOUT_PAIRS)
        st = mk.DeviceStats()
        for part in np.array_split(iel, nparts):
            gpu_insert(t, part, stats=st)
        s = st.read()
        assert s["ins_gave_up"] == 0
        pairs = H.occupied_pairs(po.Oracle(mem_p).buckets(t.dump_reference()))
        assert len(pairs) == len(iel) - s["ins_dropped"] and len(np.unique(pairs)) == len(pairs)
        t.free()
        return {"to_b2": s["ins_to_b2"], "displaced": s["ins_displaced"], "dropped": s["ins_dropped"]}

    fine = concurrent(512)
    for key, vals in env.items():
        lo, hi, mean = min(vals), max(vals), float(np.mean(vals))
        slack = (hi - lo) + 0.015 * mean + 20
        assert lo - slack <= fine[key] <= hi + slack, f"{key}: {fine[key]} outside the sequential envelope [{lo}, {hi}] +- {slack:.0f}"
    coarse = concurrent(8)
    m = {k: float(np.mean(v)) for k, v in env.items()}
    assert 0.88 * m["to_b2"] <= coarse["to_b2"] <= 1.0 * m["to_b2"], (coarse, m)
    assert 1.0 * m["displaced"] <= coarse["displaced"] <= 1.15 * m["displaced"], (coarse, m)
    assert 1.0 * m["dropped"] <= coarse["dropped"] <= 2.2 * m["dropped"] + 30, (coarse, m)
