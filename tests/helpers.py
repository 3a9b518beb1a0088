"""Shared generators for the parity tests (CPU side: numpy + the oracle)."""
import numpy as np

from oracle import pyoracle as po

IEL_DT, SEL_DT = po.IEL_DT, po.SEL_DT


def random_requests(rng, n, mem_p=None, loc_base=1):
    """n unique random keys: sig != 0, arbitrary hash, loc = loc_base + i"""
    keys = rng.integers(1, 2**63, size=int(n * 1.1) + 16, dtype=np.int64).astype(np.uint64)
    keys = np.unique(keys)
    rng.shuffle(keys)
    keys = keys[:n]
    assert len(keys) == n
    iel = np.empty(n, dtype=IEL_DT)
    sig = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    sig[sig == 0] = 1
    iel["sig"], iel["hash"] = sig, (keys >> np.uint64(32)).astype(np.uint32)
    iel["loc"] = np.arange(loc_base, loc_base + n, dtype=np.uint64).astype(np.uint32)
    return iel


def to_sel(iel):
    sel = np.empty(len(iel), dtype=SEL_DT)
    sel["sig"], sel["hash"] = iel["sig"], iel["hash"]
    return sel


def conflict_free(orc, iel):
    """Subset of `iel` in which no two requests share a candidate bucket: the outcome of inserting it
    does not depend on the order of the requests (as long as no chain leaves those buckets)."""
    b1 = orc.bucket1(iel["hash"]); b2 = orc.bucket2(iel["hash"], iel["sig"])
    seen = set(); keep = []
    for i in range(len(iel)):
        a, b = int(b1[i]), int(b2[i])
        if a in seen or b in seen:
            continue
        seen.add(a); seen.add(b); keep.append(i)
    return iel[np.array(keep, dtype=np.int64)]


def occupied_pairs(buckets):
    """sorted array of (sig<<32 | loc) over slots with sig != 0; buckets = [nb, 2, 8] view"""
    sig = buckets[:, 0, :].reshape(-1).astype(np.uint64)
    loc = buckets[:, 1, :].reshape(-1).astype(np.uint64)
    m = sig != 0
    return np.sort((sig[m] << np.uint64(32)) | loc[m])


def per_bucket_sets(buckets):
    """canonical per-bucket form: slots sorted inside each bucket by (sig, loc), empties first"""
    sig = buckets[:, 0, :].astype(np.uint64); loc = buckets[:, 1, :].astype(np.uint64)
    key = np.where(sig != 0, (sig << np.uint64(32)) | loc, np.uint64(0))
    return np.sort(key, axis=1)


def fixture_every_bucket_1_to_8(mem_p):
    """The table libgpuhash/test/back/py_search_stream.c:104-114 builds: in every bucket slot l holds
    signature l+1 and location 1."""
    nb = 1 << (mem_p - 6)
    t = np.empty((nb, 2, 8), dtype=np.uint32)
    t[:, 0, :] = np.arange(1, 9, dtype=np.uint32)
    t[:, 1, :] = 1
    return t.reshape(-1)
