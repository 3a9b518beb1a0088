"""The reference's zipf key-rank generator (src/zipf.h:44-183, SURVEY.md 8(d) Config 3) -- the oracle's restatement and the
workload module's vectorised one against vectors recorded from zipf.h itself (tests/golden/zipf_ref.npz, made by
tests/golden/make_zipf_golden.py from oracle/_ref/libzipf_ref.so), and the device generator against both."""
import ctypes as C
import os

import numpy as np
import pytest

from megakv_b200 import keystream as ks
from oracle import pyoracle as po

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "zipf_ref.npz"))
CASES = [(int(n), float(t), int(s)) for n, t, s in GOLD["cases"]]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_oracle_zipf_equals_the_reference_generator(k):
    n, theta, seed = CASES[k]
    z = po.Zipf(n, theta, seed)
    assert np.array_equal(z.ranks(20000), GOLD[f"ranks_{k}"])
    if theta > 0:
        assert z.zetan == GOLD[f"zetan_{k}"][0]                   # same sum, same order, same bits


@pytest.mark.parametrize("k", range(len(CASES)))
def test_vectorised_zipf_equals_the_reference_generator(k):
    n, theta, seed = CASES[k]
    z = ks.RefZipf(n, theta, seed)
    got = np.concatenate([z.ranks(1), z.ranks(4999), z.ranks(15000)])   # the stream continues across calls
    assert np.array_equal(got, GOLD[f"ranks_{k}"])
    if theta > 0:
        assert z.zetan == GOLD[f"zetan_{k}"][0]


def test_zipf_with_known_zetan_skips_the_sum():
    n, theta, seed = CASES[1]
    zetan = float(GOLD["zetan_1"][0])
    assert np.array_equal(po.Zipf(n, theta, seed, zetan=zetan).ranks(1000), GOLD["ranks_1"][:1000])
    assert np.array_equal(ks.RefZipf(n, theta, seed, zetan=zetan).ranks(1000), GOLD["ranks_1"][:1000])


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 4])
def test_device_zipf_equals_the_reference_generator(gpu, k):
    """gpuhash_gen_requests_ref_zipf: every thread jumps to its own element of the LCG sequence and applies the same
    rounding as the host code -> the ranks (expect_loc - 1) equal the recorded sequence, from any starting element"""
    import megakv_b200 as mk
    from megakv_b200 import _native as N
    n, theta, seed = CASES[k]
    L = N.lib()
    want = GOLD[f"ranks_{k}"]
    zetan = float(GOLD[f"zetan_{k}"][0]) if theta > 0 else 0.0
    for first, count in ((0, 20000), (777, 5000)):
        sel = mk.DeviceBuffer(8 * count); exp = mk.DeviceBuffer(4 * count)
        N.check(L.gpuhash_gen_requests_ref_zipf(sel.ptr, exp.ptr, 1, n, count, seed, first, theta, zetan, 0, None))
        mk.device_sync()
        got = exp.download(np.uint32).astype(np.uint64) - 1
        assert np.array_equal(got, want[first:first + count])
        reqs = sel.download(np.uint32).reshape(-1, 2)
        ref_sel, _ = ks.ref_zipf_queries(1, n, 64, theta, seed, zetan if theta > 0 else None)
        if first == 0:
            assert np.array_equal(reqs[:64, 0], ref_sel["sig"]) and np.array_equal(reqs[:64, 1], ref_sel["hash"])


def test_committed_zetan_table_is_what_the_restatement_computes():
    """megakv_b200/zetan_table.json saves bench.py the 2^29-term sum; its small entries are re-derived here"""
    import json
    tab = json.load(open(os.path.join(os.path.dirname(ks.__file__), "zetan_table.json")))
    for n in (1 << 21, 1 << 23):
        assert float(tab["zetan"][str(n)]) == ks.ref_zetan(n, tab["theta"])
    assert float(tab["zetan"][str(1 << 21)]) == po.Zipf(1 << 21, tab["theta"], 0).zetan
