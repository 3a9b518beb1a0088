"""CPU suite: pins the oracle (oracle/gpuhash_oracle.c) to everything the reference offers for this path.

The reference ships no golden vectors.  What it does offer, and what is checked here:
  * the property test libgpuhash/test/insert_test.c:111-244 (inserted => findable, deleted => gone,
    block-aligned hashes, <= 10 % load) -- test_insert_test_property
  * the fixture shape of libgpuhash/test/back/py_search_stream.c:104-129 -- test_py_search_stream_fixture
  * the survey's sequential-model counts (seed 1) -- test_appendix_d_anchors
  * every quirk of SURVEY.md Appendix B, as hand-built micro-cases citing gpu_hash.cu lines
  * the committed golden vectors under tests/golden/ (made by tests/golden/make_golden.py)
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import helpers as H

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def one(sig, hash_, loc):
    return np.array([(sig, hash_, loc)], dtype=po.IEL_DT)


# ----------------------------------------------------------------------------- geometry

def test_geometry_matches_reference_macros():
    # gpu_hash.h:57-69 with MEM_P = 30: HASH_MASK = 2^24-1, BLOCK_HASH_MASK = 2^21-1
    o = po.Oracle(16)
    assert o.g.hash_mask == (1 << 10) - 1 and o.g.block_mask == (1 << 7) - 1
    g = po.Geom(); po.lib().orc_geom_init(g, 30, po.CUCKOO)
    assert g.hash_mask == (1 << 24) - 1 and g.block_mask == (1 << 21) - 1 and g.max_cuckoo == 5
    po.lib().orc_geom_init(g, 34, po.CUCKOO)
    assert g.hash_mask == 0x0FFFFFFF and g.block_mask == 0x01FFFFFF


def test_alt_bucket_keeps_top3_bits_and_is_an_involution(rng):
    o = po.Oracle(20)
    h = rng.integers(0, 2**32, 10000, dtype=np.uint64).astype(np.uint32)
    s = rng.integers(1, 2**32, 10000, dtype=np.uint64).astype(np.uint32)
    b1, b2 = o.bucket1(h), o.bucket2(h, s)
    shift = 20 - 6 - 3
    assert np.array_equal(b1 >> shift, b2 >> shift)            # gpu_hash.cu:66-67 keeps hash & ~BLOCK_HASH_MASK
    assert np.array_equal(o.bucket2(b2, s), b1)                 # alt(alt(b)) == b for the same sig
    # the C functions agree with the numpy restatement
    L = po.lib()
    for i in range(50):
        assert L.orc_bucket1(o.g, int(h[i])) == b1[i] and L.orc_bucket2(o.g, int(h[i]), int(s[i])) == b2[i]


# ----------------------------------------------------------------------------- reference's own checks

@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
def test_insert_test_property(algo, rng):
    """insert_test.c:111-244 at MEM_P 20: batches of 16384 with block-aligned hashes up to 10 % load;
    every inserted loc is returned in out[2i] or out[2i+1]; after deleting the batch none is."""
    mem_p, n = 20, 16384
    o = po.Oracle(mem_p, algo)
    nb = o.num_buckets
    per_blk, hash_blk = n // 8, nb // 8                        # BLOCK_ELEM_NUM, HASH_BLOCK_ELEM_NUM (insert_test.c:34-38)
    has = 0
    while has < 0.1 * (1 << mem_p) / 8:
        iel = np.empty(n, dtype=po.IEL_DT)
        iel["sig"] = rng.integers(1, 2**31, n)
        iel["hash"] = (np.arange(n) // per_blk) * hash_blk + rng.integers(0, hash_blk, n)
        iel["loc"] = rng.integers(1, 2**31, n)
        o.insert_blocks([iel[k * per_blk:(k + 1) * per_blk] for k in range(8)])
        out = o.search(H.to_sel(iel))
        assert np.all((out[0::2] == iel["loc"]) | (out[1::2] == iel["loc"]))       # :178-195
        o.delete(iel)
        out = o.search(H.to_sel(iel))
        assert not np.any((out[0::2] == iel["loc"]) | (out[1::2] == iel["loc"]))   # :237-244
        o.insert(iel)                                                            # keep filling
        has += n


def test_py_search_stream_fixture(rng):
    """py_search_stream.c:104-129: every bucket holds sigs 1..8 / loc 1, queries sig in 1..8 and any hash
    => a hit in BOTH probed buckets for every query."""
    mem_p = 18
    o = po.Oracle(mem_p, table=H.fixture_every_bucket_1_to_8(mem_p))
    n = 50000
    sel = np.empty(n, dtype=po.SEL_DT)
    sel["hash"] = rng.integers(0, o.num_buckets, n)
    sel["sig"] = rng.integers(1, 9, n)
    assert np.all(o.search(sel) == 1)


APPENDIX_D = [  # LF, N, to_b2, displaced, dropped, occupied, findable   (MEM_P 26, seed 1)
    (0.125, 1048576, 1, 0, 0, 1048576, 1048576),
    (0.5, 4194304, 36156, 1097, 0, 4194304, 4193207),
    (0.9, 7549747, 965361, 668217, 12880, 7536867, 6901368),
]


@pytest.mark.parametrize("lf,n,to_b2,disp,drop,occ,found", APPENDIX_D)
def test_appendix_d_anchors(lf, n, to_b2, disp, drop, occ, found):
    o = po.Oracle(26, po.CUCKOO)
    iel, sel = po.keys(1, 0, n)
    o.insert(iel)
    s = o.stats.as_dict()
    assert (s["to_b2"], s["displaced"], s["dropped"]) == (to_b2, disp, drop)
    assert o.occupied() == occ
    out = o.search(sel)
    hit = (out[0::2] == iel["loc"]) | (out[1::2] == iel["loc"])
    assert int(hit.sum()) == found
    assert int(((out[0::2] != 0) | (out[1::2] != 0)).sum()) == found
    assert sum(s["chain_hist"]) == n


# ----------------------------------------------------------------------------- quirks

def fill_bucket(o, b, sigs, locs):
    t = o.buckets()
    t[b, 0, :len(sigs)] = sigs
    t[b, 1, :len(locs)] = locs


def test_search_probes_both_buckets_and_reports_both():
    """B.4 (gpu_hash.cu:61-63 commented out): a signature present in both candidate buckets is reported twice."""
    o = po.Oracle(16)
    sig, h = 0x1234, 5
    b1, b2 = int(o.bucket1(h)), int(o.bucket2(h, sig))
    assert b1 != b2
    fill_bucket(o, b1, [9, sig], [90, 111])
    fill_bucket(o, b2, [sig], [222])
    out = o.search(np.array([(sig, h)], dtype=po.SEL_DT))
    assert list(out) == [111, 222]
    # a different key with the same signature whose bucket 1 is b2: false positive by design
    out = o.search(np.array([(sig, b2)], dtype=po.SEL_DT))
    assert out[0] == 222


def test_search_when_alt_equals_b1_reports_same_loc_twice():
    o = po.Oracle(16)
    sig = 0x55 << 16            # low BLOCK_HASH_MASK bits of sig are 0 -> alt == b1 (Appendix A.1)
    assert sig & o.g.block_mask == 0
    o.insert(one(sig, 77, 4242))
    assert list(o.search(np.array([(sig, 77)], dtype=po.SEL_DT))) == [4242, 4242]


def test_major_location_and_circular_first_empty():
    """gpu_hash.cu:301-309: first empty slot scanning from sig & 7 and wrapping."""
    o = po.Oracle(16)
    h = 3
    o.insert(one(0x10 | 5, h, 1))                     # empty bucket: slot 5
    assert o.buckets()[3, 0, 5] == 0x15
    o.insert(one(0x20 | 5, h, 2))                     # slot 5 taken -> 6
    o.insert(one(0x30 | 5, h, 3))                     # -> 7
    o.insert(one(0x40 | 5, h, 4))                     # wraps -> 0
    assert list(o.buckets()[3, 0, :]) == [0x45, 0, 0, 0, 0, 0x15, 0x25, 0x35]
    assert list(o.buckets()[3, 1, :]) == [4, 0, 0, 0, 0, 1, 2, 3]


def test_insert_existing_signature_updates_lowest_matching_slot():
    """gpu_hash.cu:277-287: __ffs(ballot)-1 = lowest lane; no second copy is made."""
    o = po.Oracle(16)
    o.insert(one(0xABCD, 9, 1)); o.insert(one(0xABCD, 9, 2))
    assert o.occupied() == 1 and o.stats.updated == 1
    assert list(o.search(np.array([(0xABCD, 9)], dtype=po.SEL_DT)))[0] == 2


def _full_bucket(o, b, base_sig):
    fill_bucket(o, b, [base_sig + l for l in range(8)], [1000 + base_sig + l for l in range(8)])


def test_cuckoo_victim_rehomed_with_requests_hash_not_its_own():
    """B.1 (gpu_hash.cu:334-335, 403-405): the victim goes to alt(request.hash, victim.sig)."""
    o = po.Oracle(16, po.CUCKOO)
    sig, h = 0x0A03, 40                                # major slot 3
    b1, b2 = int(o.bucket1(h)), int(o.bucket2(h, sig))
    _full_bucket(o, b1, 0x100); _full_bucket(o, b2, 0x208)
    victim_sig, victim_loc = int(o.buckets()[b2, 0, 3]), int(o.buckets()[b2, 1, 3])
    o.insert(one(sig, h, 7))
    t = o.buckets()
    assert t[b2, 0, 3] == sig and t[b2, 1, 3] == 7                               # :360 victim slot = sig0 & 7
    b3 = int(o.bucket2(h, victim_sig))                                           # NOT alt(victim's own bucket)
    assert b3 not in (b1, b2)
    assert t[b3, 0, 3] == victim_sig and t[b3, 1, 3] == victim_loc               # first empty from major 3 (:301)
    assert o.stats.displaced == 1 and o.stats.dropped == 0 and o.stats.chain_hist[1] == 1


def test_cuckoo_sixth_displacement_drops_the_victim():
    """gpu_hash.cu:361-365 vs :414-422: five victims are carried on, the sixth is overwritten and lost."""
    o = po.Oracle(16, po.CUCKOO)
    t = o.buckets()
    t[:, 0, :] = np.arange(1, 1 + t.shape[0] * 8, dtype=np.uint32).reshape(-1, 8) | 0x40000000   # every slot full
    t[:, 1, :] = 7
    before = H.occupied_pairs(o.buckets())
    o.insert(one(0x0BB1, 12, 99))
    after = H.occupied_pairs(o.buckets())
    assert o.stats.displaced == 5 and o.stats.dropped == 1 and o.stats.chain_hist[5] == 1
    assert len(after) == len(before)                                             # one in, one lost
    assert len(np.setdiff1d(before, after)) == 1 and len(np.setdiff1d(after, before)) == 1


def test_two_choice_full_full_overwrites_signature_only():
    """B.3 (gpu_hash.cu:197-209): the new signature inherits the victim's location."""
    o = po.Oracle(16, po.TWO_CHOICE)
    sig, h = 0x0C06, 21
    b1, b2 = int(o.bucket1(h)), int(o.bucket2(h, sig))
    _full_bucket(o, b1, 0x300); _full_bucket(o, b2, 0x400)
    old_loc = int(o.buckets()[b2, 1, 6])
    o.insert(one(sig, h, 555))
    assert o.buckets()[b2, 0, 6] == sig and o.buckets()[b2, 1, 6] == old_loc
    assert o.stats.overwritten == 1
    assert list(o.search(np.array([(sig, h)], dtype=po.SEL_DT))) == [0, old_loc]


def test_delete_needs_sig_and_loc_and_leaves_loc_stale():
    """B.5 (gpu_hash.cu:459-463, 474-476)."""
    o = po.Oracle(16)
    o.insert(one(0x77, 8, 31))
    assert o.delete(one(0x77, 8, 32)) == 0 and o.occupied() == 1                 # wrong loc: no-op
    assert o.delete(one(0x77, 8, 31)) == 1 and o.occupied() == 0
    b = int(o.bucket1(8))
    assert o.buckets()[b, 1, 7] == 31                                            # loc word untouched
    assert o.delete(one(0x77, 8, 31)) == 0


def test_delete_skips_bucket2_when_bucket1_matched():
    """gpu_hash.cu:465-468."""
    o = po.Oracle(16)
    sig, h = 0x4321, 17
    b1, b2 = int(o.bucket1(h)), int(o.bucket2(h, sig))
    fill_bucket(o, b1, [sig], [5]); fill_bucket(o, b2, [sig], [5])
    assert o.delete(one(sig, h, 5)) == 1
    assert o.buckets()[b1, 0, 0] == 0 and o.buckets()[b2, 0, 0] == sig
    assert o.delete(one(sig, h, 5)) == 1                                         # now bucket 1 misses -> bucket 2
    assert o.buckets()[b2, 0, 0] == 0


def test_all_zero_request_is_skipped_and_zero_sig_writes_loc_into_an_empty_slot():
    """gpu_hash.cu:259-262; Appendix A.2 note on sig0 == 0, loc0 != 0."""
    o = po.Oracle(16)
    o.insert(one(0, 3, 0))
    assert o.stats.skipped == 1 and not o.buckets().any()
    o.insert(one(0, 3, 9))
    assert o.occupied() == 0 and o.buckets()[3, 1, 0] == 9                       # "match" on the lowest empty slot
    assert o.search(np.array([(0, 3)], dtype=po.SEL_DT))[0] == 9                 # lowest empty slot now holds loc 9


def test_search_lowest_matching_lane_wins_on_duplicate_signatures():
    """what the reference's kernel does on a B200: tests/golden/ref_search_cuckoo_16.npz (dup_* arrays)"""
    o = po.Oracle(16)
    fill_bucket(o, 4, [0x99, 1, 0x99], [10, 11, 12])
    assert o.search(np.array([(0x99, 4)], dtype=po.SEL_DT))[0] == 10


# ----------------------------------------------------------------------------- drivers

def test_insert_blocks_equals_flat_insert_in_block_order(rng):
    iel = H.random_requests(rng, 40000)
    a, b = po.Oracle(18), po.Oracle(18)
    blocks = [iel[i::8] for i in range(8)]
    a.insert_blocks(blocks)
    b.insert(np.concatenate(blocks))
    assert np.array_equal(a.table, b.table)


@pytest.mark.parametrize("threads", [2, 4, 8])
def test_threaded_drivers_equal_sequential(threads, rng):
    iel = H.random_requests(rng, 120000)                    # ~92 % load at MEM_P 20: long chains
    a, b = po.Oracle(20), po.Oracle(20)
    a.insert(iel); b.insert_mt(iel, threads)
    assert np.array_equal(a.table, b.table)
    sel = H.to_sel(iel)
    assert np.array_equal(a.search(sel), b.search_mt(sel, threads))


def test_digest_is_order_independent_and_sensitive(rng):
    iel = H.random_requests(rng, 5000)
    a, b = po.Oracle(18), po.Oracle(18)
    a.insert(iel); b.insert(iel[::-1])
    assert a.digest() == b.digest() and a.digest(True) == b.digest(True)
    b.buckets()[int(b.bucket1(iel["hash"][0])), 1, :] ^= 1
    assert a.digest() != b.digest()


# ----------------------------------------------------------------------------- committed golden vectors

@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")) if os.path.isdir(GOLDEN) else [])
def test_golden_vectors(name):
    from tests.golden import make_golden as MG
    g = np.load(os.path.join(GOLDEN, name))
    if name.startswith("ref_"):
        pytest.skip("reference-kernel vectors are checked in test_ref_golden.py")
    if name.startswith("zipf_"):
        pytest.skip("the reference's zipf generator is checked in test_zipf.py")
    got = MG.run_case(str(g["case"]))
    for k in g.files:
        if k == "case":
            continue
        assert np.array_equal(got[k], g[k]), f"{name}: {k} differs"


def test_fold_keys_and_compact_results_restatements():
    """src/mega_recv.c:349-362 and src/mega_send.c:411-414 on hand-computed cases"""
    k = np.arange(20, dtype=np.uint8).reshape(1, 20)
    w0 = int.from_bytes(bytes(range(0, 8)), "little"); w1 = int.from_bytes(bytes(range(8, 16)), "little")
    tail = int.from_bytes(bytes(range(16, 20)), "little")
    s = po.fold_keys(k, True)
    assert (int(s["hash"][0]) << 32 | int(s["sig"][0])) == w0 ^ w1 ^ tail
    s = po.fold_keys(k, False)
    assert (int(s["hash"][0]) << 32 | int(s["sig"][0])) == w0
    s = po.fold_keys(k[:, :8], True)
    assert (int(s["hash"][0]) << 32 | int(s["sig"][0])) == w0
    assert po.compact_results(np.array([5, 0, 0, 7, 0, 0, 3, 4], dtype=np.uint32)).tolist() == [5, 7, 0, 3]


def test_local_test_key_formula_has_a_degenerate_alternate_bucket():
    """src/mega_recv.c:690-704 (LOCAL_TEST): key bytes {k, (bswap32(k & 0xff) << (8 - bits)) | k} -> sig = k, hash = ((k & 7) << 29) | k
    with 3 insert-buffer bits.  hash ^ sig keeps only the top bits, so gpu_hash.cu:66-67 sends EVERY key of a block to the
    block's bucket 0 as its alternate: after the first buckets fill, a block's overflow meets in one bucket."""
    mem_p = 20
    o = po.Oracle(mem_p)
    k = np.arange(1, 200001, dtype=np.uint64).astype(np.uint32)
    hash_ = (((k & np.uint32(0xFF)) << np.uint32(24)) << np.uint32(5)) | k
    assert np.array_equal(hash_, ((k & np.uint32(7)) << np.uint32(29)) | k)
    b1, b2 = o.bucket1(hash_), o.bucket2(hash_, k)
    nb = o.num_buckets
    assert np.array_equal(b2, b1 & np.uint32(~((nb >> 3) - 1) & (nb - 1)))       # bucket 0 of bucket 1's block
    assert len(np.unique(b2)) == 8
    iel = np.empty(len(k), dtype=po.IEL_DT); iel["sig"], iel["hash"], iel["loc"] = k, hash_, k
    o.insert(iel)                                                    # 200 000 keys, 131 072 slots: every first bucket fills
    st = o.stats.as_dict()
    assert st["to_b2"] > 0 and st["dropped"] > 0 and st["placed_b2"] <= 8 * 8 + st["displaced"]
    found = (o.search(H.to_sel(iel)).reshape(-1, 2) == k[:, None]).any(axis=1)
    assert found[: 8 * nb // 2].mean() > 0.8                          # most early keys still sit in their first buckets
