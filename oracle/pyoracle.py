"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  Nothing under megakv_b200/ does.

The oracle restates pzrq/megakv libgpuhash/gpu_hash.cu:28-480 one request at a time
(see gpuhash_oracle.h for the line-by-line map).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

CUCKOO, TWO_CHOICE = 0, 1

SEL_DT = np.dtype([("sig", "<u4"), ("hash", "<u4")])                       # selem_t
IEL_DT = np.dtype([("sig", "<u4"), ("hash", "<u4"), ("loc", "<u4")])       # ielem_t / delem_t


class Geom(C.Structure):
    _fields_ = [("hash_mask", C.c_uint32), ("block_mask", C.c_uint32),
                ("algo", C.c_int), ("max_cuckoo", C.c_int)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("skipped", "updated", "placed_b1", "placed_b2", "to_b2",
                 "displaced", "dropped", "overwritten")] + [("chain_hist", C.c_uint64 * 8)]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n, _ in self._fields_[:-1]}
        d["chain_hist"] = [int(v) for v in self.chain_hist]
        return d


def build():
    src = [os.path.join(_HERE, f) for f in ("gpuhash_oracle.c", "gpuhash_oracle.h", "Makefile")]
    if (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        gp = C.POINTER(Geom)
        L.orc_geom_init.argtypes = [gp, C.c_int, C.c_int]
        L.orc_table_bytes.argtypes = [gp]; L.orc_table_bytes.restype = sz
        L.orc_bucket1.argtypes = [gp, C.c_uint32]; L.orc_bucket1.restype = C.c_uint32
        L.orc_bucket2.argtypes = [gp, C.c_uint32, C.c_uint32]; L.orc_bucket2.restype = C.c_uint32
        L.orc_search.argtypes = [vp, gp, vp, sz, vp]
        L.orc_insert.argtypes = [vp, gp, vp, sz, C.POINTER(Stats)]
        L.orc_insert_blocks.argtypes = [vp, gp, C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(Stats)]
        L.orc_delete.argtypes = [vp, gp, vp, sz]; L.orc_delete.restype = u64
        L.orc_table_occupied.argtypes = [vp, gp]; L.orc_table_occupied.restype = u64
        L.orc_table_digest.argtypes = [vp, gp, C.c_int, C.POINTER(u64 * 2)]
        L.orc_keys_fill.argtypes = [u64, u64, sz, vp, vp]
        L.orc_search_mt.argtypes = [vp, gp, vp, sz, vp, C.c_int]
        L.orc_insert_mt.argtypes = [vp, gp, vp, sz, C.c_int]
        L.orc_now_sec.restype = C.c_double
        L.orc_pool_create.argtypes = [C.c_int]; L.orc_pool_create.restype = vp
        L.orc_pool_threads.argtypes = [vp]; L.orc_pool_threads.restype = C.c_int
        L.orc_pool_destroy.argtypes = [vp]
        L.orc_pool_cycle.argtypes = [vp, vp, gp, vp, sz, vp, vp, sz, C.c_int]
        L.orc_pool_insert.argtypes = [vp, vp, gp, vp, sz]
        L.orc_pool_preload.argtypes = [vp, vp, gp, u64, u64, u64]
        L.orc_gen_queries.argtypes = [u64, u64, sz, u64, vp]
        L.orc_zipf_init.argtypes = [vp, u64, C.c_double, u64]
        L.orc_zipf_init_zetan.argtypes = [vp, u64, C.c_double, u64, C.c_double]
        L.orc_zipf_fill.argtypes = [vp, sz, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """A table in the reference's byte layout (bucket_t[], gpu_hash.h:79-82) plus the
    sequential operations on it."""

    def __init__(self, mem_p, algo=CUCKOO, table=None):
        self.L = lib()
        self.mem_p, self.algo = mem_p, algo
        self.g = Geom()
        self.L.orc_geom_init(C.byref(self.g), mem_p, algo)
        nbytes = self.L.orc_table_bytes(C.byref(self.g))
        assert nbytes == 1 << mem_p
        if table is None:
            self.table = np.zeros(nbytes // 4, dtype=np.uint32)
        else:
            self.table = np.ascontiguousarray(table).view(np.uint32).copy()
            assert self.table.nbytes == nbytes
        self.stats = Stats()

    # --- geometry
    @property
    def num_buckets(self):
        return int(self.g.hash_mask) + 1

    def bucket1(self, hash_):
        return np.asarray(hash_, dtype=np.uint32) & np.uint32(self.g.hash_mask)

    def bucket2(self, hash_, sig):
        h = np.asarray(hash_, dtype=np.uint32); s = np.asarray(sig, dtype=np.uint32)
        bm = np.uint32(self.g.block_mask)
        return (((h ^ s) & bm) | (h & ~bm)) & np.uint32(self.g.hash_mask)

    # --- operations
    def search(self, sel):
        sel = np.ascontiguousarray(sel, dtype=SEL_DT)
        out = np.zeros(2 * len(sel), dtype=np.uint32)            # the caller's memset
        self.L.orc_search(_ptr(self.table), C.byref(self.g), _ptr(sel), len(sel), _ptr(out))
        return out

    def insert(self, iel):
        iel = np.ascontiguousarray(iel, dtype=IEL_DT)
        self.L.orc_insert(_ptr(self.table), C.byref(self.g), _ptr(iel), len(iel), C.byref(self.stats))

    def insert_blocks(self, blocks):
        blocks = [np.ascontiguousarray(b, dtype=IEL_DT) for b in blocks]
        ptrs = (C.c_void_p * len(blocks))(*[b.ctypes.data for b in blocks])
        nums = (C.c_int * len(blocks))(*[len(b) for b in blocks])
        self.L.orc_insert_blocks(_ptr(self.table), C.byref(self.g), ptrs, nums, len(blocks), C.byref(self.stats))

    def delete(self, iel):
        iel = np.ascontiguousarray(iel, dtype=IEL_DT)
        return int(self.L.orc_delete(_ptr(self.table), C.byref(self.g), _ptr(iel), len(iel)))

    def search_mt(self, sel, threads):
        sel = np.ascontiguousarray(sel, dtype=SEL_DT)
        out = np.zeros(2 * len(sel), dtype=np.uint32)
        self.L.orc_search_mt(_ptr(self.table), C.byref(self.g), _ptr(sel), len(sel), _ptr(out), threads)
        return out

    def insert_mt(self, iel, threads):
        iel = np.ascontiguousarray(iel, dtype=IEL_DT)
        self.L.orc_insert_mt(_ptr(self.table), C.byref(self.g), _ptr(iel), len(iel), threads)

    # --- inspection
    def occupied(self):
        return int(self.L.orc_table_occupied(_ptr(self.table), C.byref(self.g)))

    def digest(self, per_bucket=False, table=None):
        t = self.table if table is None else np.ascontiguousarray(table).view(np.uint32)
        d = (C.c_uint64 * 2)()
        self.L.orc_table_digest(_ptr(t), C.byref(self.g), int(per_bucket), C.byref(d))
        return (int(d[0]), int(d[1]))

    def buckets(self, table=None):
        """view as [num_buckets, 2, 8]: [:,0,:] signatures, [:,1,:] locations"""
        t = self.table if table is None else np.ascontiguousarray(table).view(np.uint32)
        return t.reshape(-1, 2, 8)


class Pool:
    """persistent host threads for the CPU arm of bench.py (orc_pool_*)"""

    def __init__(self, threads):
        self.L = lib()
        self.h = self.L.orc_pool_create(threads)
        self.threads = self.L.orc_pool_threads(self.h)

    def close(self):
        if self.h:
            self.L.orc_pool_destroy(self.h); self.h = None

    def cycle(self, orc, sel, out, iel, batches):
        """sel [batches * n_search], out uint32 [2 * batches * n_search], iel [batches * n_insert]"""
        self.L.orc_pool_cycle(self.h, _ptr(orc.table), C.byref(orc.g), _ptr(sel), len(sel) // batches, _ptr(out),
                              _ptr(iel), len(iel) // batches, batches)

    def insert(self, orc, iel):
        iel = np.ascontiguousarray(iel, dtype=IEL_DT)
        self.L.orc_pool_insert(self.h, _ptr(orc.table), C.byref(orc.g), _ptr(iel), len(iel))

    def preload(self, orc, seed, first, count):
        self.L.orc_pool_preload(self.h, _ptr(orc.table), C.byref(orc.g), seed, first, count)


class ZipfState(C.Structure):                 # orc_zipf_t
    _fields_ = [("n", C.c_uint64), ("theta", C.c_double), ("alpha", C.c_double), ("thres", C.c_double), ("dbl_n", C.c_double),
                ("zetan", C.c_double), ("eta", C.c_double), ("rand_state", C.c_uint64)]


class Zipf:
    """the reference's key-rank generator (src/zipf.h:73-183) as restated in gpuhash_oracle.c"""

    def __init__(self, n, theta, rand_seed, zetan=None):
        self.st = ZipfState()
        if zetan is None:
            lib().orc_zipf_init(C.byref(self.st), n, theta, rand_seed)
        else:
            lib().orc_zipf_init_zetan(C.byref(self.st), n, theta, rand_seed, zetan)

    @property
    def zetan(self):
        return self.st.zetan

    def ranks(self, m):
        out = np.empty(m, dtype=np.uint64)
        lib().orc_zipf_fill(C.byref(self.st), m, _ptr(out))
        return out


def gen_queries(seed, population, n, rng_seed):
    sel = np.empty(n, dtype=SEL_DT)
    lib().orc_gen_queries(seed, population, n, rng_seed, _ptr(sel))
    return sel


def keys(seed, first_index, n):
    """(ielem[n], selem[n]) of the SURVEY 8(d) splitmix64 key stream."""
    iel = np.empty(n, dtype=IEL_DT); sel = np.empty(n, dtype=SEL_DT)
    lib().orc_keys_fill(seed, first_index, n, _ptr(iel), _ptr(sel))
    return iel, sel


def now():
    return lib().orc_now_sec()


# ---- the steps either side of the path, restated for the tests ----

def fold_keys(keys, fold):
    """src/mega_recv.c:349-362.  keys: uint8 [n, nkey], nkey >= 8.  sig64 = the first 8 key bytes (little endian); with
    the reference's -DSIGNATURE (fold=True) every further full 8-byte word is XORed in (:352-354) and the last, partial
    word masked to the key's own bytes (:355-358).  hash = high 32 bits, sig = low 32 bits (:361-362)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint8)
    n, nkey = keys.shape
    assert nkey >= 8
    sig = keys[:, :8].copy().view("<u8").reshape(n)
    if fold:
        i = 8
        while i + 8 <= nkey:
            sig = sig ^ keys[:, i:i + 8].copy().view("<u8").reshape(n)
            i += 8
        if i < nkey:
            tail = np.zeros((n, 8), dtype=np.uint8)
            tail[:, : nkey - i] = keys[:, i:]
            sig = sig ^ tail.view("<u8").reshape(n)
    sel = np.empty(n, dtype=SEL_DT)
    sel["sig"] = (sig & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    sel["hash"] = (sig >> np.uint64(32)).astype(np.uint32)
    return sel


def compact_results(out):
    """src/mega_send.c:411-414: the sender takes search_out[2i], and search_out[2i+1] when that is 0."""
    o = np.asarray(out, dtype=np.uint32).reshape(-1, 2)
    return np.where(o[:, 0] != 0, o[:, 0], o[:, 1])
