"""Golden vectors produced by the REFERENCE's own kernels (GPU box only).

oracle/_ref/libgpuhash_ref_<policy>_<MEM_P>.so is pzrq/megakv libgpuhash/gpu_hash.cu compiled where it lies
(oracle/Makefile: compute_60 PTX -> sm_100 SASS, the only way its pre-Volta __ballot() assembles).  This script
drives those kernels through the reference's three entry points on seeded inputs and records what THEY return:

  ref_search_*    gpu_hash_search on an uploaded table (launch shape 24576 / 256, mega.c:163-165)
  ref_serial_*    gpu_hash_insert called with ONE request per launch, so the reference runs sequentially and its
                  result is deterministic even through eviction chains; final table bytes + search results
  ref_batch_*     the insert_test.c scenario: 8 block-aligned segments per launch at low load, then delete
  ref_delete_*    gpu_hash_delete on an uploaded table; final table bytes

Outputs go to gpurun_out/ref_golden/ (copied into tests/golden/ and committed by hand).  tests/test_ref_golden.py
then holds the CPU oracle -- and the CUDA path -- to these vectors.  Inputs are stored with the outputs, so the
replay needs neither this script's RNG nor the reference.

    timeout 600 python -m tests.golden.make_ref_golden
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import megakv_b200 as mk  # noqa: E402  (device memory plumbing only; no megakv_b200 kernel runs here)

IEL_DT, SEL_DT = mk.IEL_DT, mk.SEL_DT
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden")


class RefLib:
    def __init__(self, algo, mem_p):
        path = os.path.join(ROOT, "oracle", "_ref", f"libgpuhash_ref_{algo}_{mem_p}.so")
        self.L = C.CDLL(path)
        vp, i = C.c_void_p, C.c_int
        self.L.gpu_hash_search.argtypes = [vp, vp, vp, i, i, i, vp]; self.L.gpu_hash_search.restype = None
        self.L.gpu_hash_insert.argtypes = [vp, vp, vp, i, vp]; self.L.gpu_hash_insert.restype = None
        self.L.gpu_hash_delete.argtypes = [vp, vp, i, i, i, vp]; self.L.gpu_hash_delete.restype = None
        self.mem_p = mem_p
        self.table = mk.DeviceBuffer(1 << mem_p, zero=True)

    def load(self, words):
        self.table.upload(np.ascontiguousarray(words).view(np.uint32)); mk.device_sync()

    def dump(self):
        mk.device_sync()
        return self.table.download(np.uint32)

    def search(self, sel, num_thread=24576, tpb=256):
        sel = np.ascontiguousarray(sel, dtype=SEL_DT)
        in_d = mk.DeviceBuffer.from_host(sel)
        out_d = mk.DeviceBuffer(8 * len(sel), zero=True)                     # the caller's memset (mega_scheduler.c:406)
        self.L.gpu_hash_search(in_d.ptr, out_d.ptr, self.table.ptr, len(sel), num_thread, tpb, None)
        mk.device_sync()
        return out_d.download(np.uint32)

    def insert_blocks(self, blocks):
        segs = mk.InsertSegments(blocks)
        self.L.gpu_hash_insert(self.table.ptr, segs.ptrs.ptr, segs.nums.ptr, segs.num_blks, None)
        mk.device_sync()

    def insert_one_by_one(self, iel):
        """one request per launch: the reference's own code, executed in request order"""
        iel = np.ascontiguousarray(iel, dtype=IEL_DT)
        in_d = mk.DeviceBuffer.from_host(iel)
        ptr_d = mk.DeviceBuffer(8); num_d = mk.DeviceBuffer.from_host(np.array([1], dtype=np.int32))
        ptrs = np.zeros(1, dtype=np.uint64)
        for k in range(len(iel)):
            ptrs[0] = in_d.ptr + 12 * k
            ptr_d.upload(ptrs)
            self.L.gpu_hash_insert(self.table.ptr, ptr_d.ptr, num_d.ptr, 1, None)
        mk.device_sync()

    def delete(self, iel, num_thread=16384, tpb=256):
        iel = np.ascontiguousarray(iel, dtype=IEL_DT)
        in_d = mk.DeviceBuffer.from_host(iel)
        self.L.gpu_hash_delete(in_d.ptr, self.table.ptr, len(iel), num_thread, tpb, None)
        mk.device_sync()


def reqs(rng, n, loc0=1):
    iel = np.empty(n, dtype=IEL_DT)
    iel["sig"] = rng.integers(1, 2**32, n, dtype=np.uint64).astype(np.uint32)
    iel["hash"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    iel["loc"] = np.arange(loc0, loc0 + n, dtype=np.uint32)
    return iel


def to_sel(iel):
    s = np.empty(len(iel), dtype=SEL_DT); s["sig"], s["hash"] = iel["sig"], iel["hash"]
    return s


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()}, flush=True)


def case_serial(algo, mem_p, load, seed, tag):
    """sequential reference run through eviction chains, drops and 2-choice overwrites"""
    rng = np.random.default_rng(seed)
    r = RefLib(algo, mem_p)
    n = int(load * (1 << mem_p) / 8)
    iel = reqs(rng, n)
    iel[n // 3] = iel[n // 5]; iel["loc"][n // 3] = 424242                  # an in-place update
    r.insert_one_by_one(iel)
    probe = np.concatenate([to_sel(iel), to_sel(reqs(rng, 200))])
    save(f"ref_serial_{algo}_{mem_p}_{tag}", kind=np.array("serial"), algo=np.array(algo), mem_p=np.array(mem_p),
         iel=iel.view(np.uint32), sel=probe.view(np.uint32), table=r.dump(), out=r.search(probe))


def case_search(algo, mem_p, seed):
    """search on a hand-built table: random fill + the py_search_stream.c fixture rows + duplicate signatures"""
    rng = np.random.default_rng(seed)
    nb = 1 << (mem_p - 6)
    t = np.zeros((nb, 2, 8), dtype=np.uint32)
    fill = rng.random((nb, 8)) < 0.6
    t[:, 0, :] = np.where(fill, rng.integers(1, 2**32, (nb, 8), dtype=np.uint64).astype(np.uint32), 0)
    t[:, 1, :] = rng.integers(1, 2**32, (nb, 8), dtype=np.uint64).astype(np.uint32)       # stale locs under empty sigs too
    t[: nb // 4, 0, :] = np.arange(1, 9, dtype=np.uint32)                   # py_search_stream.c:104-114 rows
    t[: nb // 4, 1, :] = 1
    sig = t[:, 0, :]
    pick_b = rng.integers(0, nb, 4000); pick_l = rng.integers(0, 8, 4000)
    sel = np.empty(4000 + 2000 + 500, dtype=SEL_DT)
    sel["sig"][:4000] = sig[pick_b, pick_l]; sel["hash"][:4000] = pick_b | (rng.integers(0, 2**32 >> (mem_p - 6), 4000) << (mem_p - 6))
    sel["sig"][4000:6000] = rng.integers(1, 9, 2000); sel["hash"][4000:6000] = rng.integers(0, nb // 4, 2000)
    sel["sig"][6000:] = rng.integers(1, 2**32, 500, dtype=np.uint64); sel["hash"][6000:] = rng.integers(0, 2**32, 500, dtype=np.uint64)
    sel = sel[sel["sig"] != 0]
    r = RefLib(algo, mem_p); r.load(t.reshape(-1))
    out = r.search(sel)
    # same signature in two slots of one bucket: which lane's store survives is a hardware matter; record it separately
    d = np.zeros((nb, 2, 8), dtype=np.uint32)
    d[:, 0, :] = [7, 7, 3, 7, 0, 0, 3, 9]; d[:, 1, :] = np.arange(1, 9)
    dsel = np.array([(7, 5), (3, 9), (0, 11)], dtype=SEL_DT)
    r.load(d.reshape(-1))
    save(f"ref_search_{algo}_{mem_p}", kind=np.array("search"), algo=np.array(algo), mem_p=np.array(mem_p),
         table=t.reshape(-1), sel=sel.view(np.uint32), out=out,
         dup_table=d.reshape(-1), dup_sel=dsel.view(np.uint32), dup_out=r.search(dsel))


def case_batch(algo, mem_p, seed):
    """libgpuhash/test/insert_test.c:111-244: 8 block-aligned segments per launch, low load, then delete half"""
    rng = np.random.default_rng(seed)
    r = RefLib(algo, mem_p)
    nb = 1 << (mem_p - 6)
    n = 4096
    batches, outs, tables = [], [], []
    for it in range(4):
        iel = reqs(rng, n, loc0=1 + it * n)
        iel["hash"] = (np.arange(n) // (n // 8)) * (nb // 8) + rng.integers(0, nb // 8, n)
        r.insert_blocks([iel[k * (n // 8):(k + 1) * (n // 8)] for k in range(8)])
        outs.append(r.search(to_sel(iel), 16384, 128))
        if it % 2:
            r.delete(iel[: n // 2], 16384, 128)
        tables.append(r.dump()); batches.append(iel.view(np.uint32))
    save(f"ref_batch_{algo}_{mem_p}", kind=np.array("batch"), algo=np.array(algo), mem_p=np.array(mem_p),
         batches=np.stack(batches), outs=np.stack(outs), tables=np.stack(tables))


def case_delete(algo, mem_p, seed):
    rng = np.random.default_rng(seed)
    nb = 1 << (mem_p - 6)
    t = np.zeros((nb, 2, 8), dtype=np.uint32)
    fill = rng.random((nb, 8)) < 0.7
    t[:, 0, :] = np.where(fill, rng.integers(1, 2**20, (nb, 8)).astype(np.uint32), 0)
    t[:, 1, :] = rng.integers(1, 4, (nb, 8)).astype(np.uint32)
    b = rng.integers(0, nb, 3000); l = rng.integers(0, 8, 3000)
    dele = np.empty(3000, dtype=IEL_DT)
    dele["sig"], dele["hash"], dele["loc"] = t[b, 0, l], b, t[b, 1, l]
    dele["loc"][::4] += 1
    dele = dele[dele["sig"] != 0]
    r = RefLib(algo, mem_p); r.load(t.reshape(-1))
    r.delete(dele)
    save(f"ref_delete_{algo}_{mem_p}", kind=np.array("delete"), algo=np.array(algo), mem_p=np.array(mem_p),
         table_in=t.reshape(-1), dele=dele.view(np.uint32), table=r.dump())


def sparse(table_words):
    """(indices of the non-empty buckets, their 16 words): what a mostly empty table of 2^20 .. 2^26 bytes is stored as"""
    t = np.asarray(table_words, dtype=np.uint32).reshape(-1, 16)
    idx = np.nonzero(t.any(axis=1))[0].astype(np.uint32)
    return idx, t[idx].copy()


def bucket2_of(mem_p, hash_, sig):
    """gpu_hash.cu:66-67 with the geometry macros of gpu_hash.h:57-69 (only used to AIM probes; the answers come from the kernels)"""
    hm = np.uint32((1 << (mem_p - 6)) - 1); bm = np.uint32((1 << (mem_p - 9)) - 1); nbm = np.uint32(~int(bm) & 0xFFFFFFFF)
    return ((((hash_ ^ sig) & bm) | (hash_ & nbm)) & hm).astype(np.uint32)


def case_search_sparse(algo, mem_p, seed, nbuckets=20000):
    """gpu_hash_search at a larger geometry (BLOCK_HASH_MASK of 11 / 17 bits) on a sparse hand-built table: keys that sit in
    bucket 1 of their probe, keys that sit in the ALTERNATE bucket of their probe, duplicate signatures, misses"""
    rng = np.random.default_rng(seed)
    nb = 1 << (mem_p - 6)
    t = np.zeros((nb, 2, 8), dtype=np.uint32)
    n1 = nbuckets // 2
    b = rng.choice(nb, size=n1, replace=False)
    fill = rng.random((n1, 8)) < 0.6
    t[b, 0, :] = np.where(fill, rng.integers(1, 2**32, (n1, 8), dtype=np.uint64).astype(np.uint32), 0)
    t[b, 1, :] = rng.integers(1, 2**32, (n1, 8), dtype=np.uint64).astype(np.uint32)
    # probes whose key sits in bucket 1
    pl = rng.integers(0, 8, n1)
    sel1 = np.empty(n1, dtype=SEL_DT); sel1["sig"] = t[b, 0, pl]; sel1["hash"] = b | (rng.integers(0, 2**32 >> (mem_p - 6), n1) << (mem_p - 6)).astype(np.uint32)
    # keys placed into the alternate bucket of their probe: choose (hash, sig), put sig into slot sig & 7 of bucket2
    n2 = nbuckets - n1
    h2 = rng.integers(0, 2**32, n2, dtype=np.uint64).astype(np.uint32); s2 = rng.integers(1, 2**32, n2, dtype=np.uint64).astype(np.uint32)
    b2 = bucket2_of(mem_p, h2, s2)
    t[b2, 0, s2 & 7] = s2; t[b2, 1, s2 & 7] = rng.integers(1, 2**32, n2, dtype=np.uint64).astype(np.uint32)
    sel2 = np.empty(n2, dtype=SEL_DT); sel2["sig"], sel2["hash"] = s2, h2
    miss = np.empty(1000, dtype=SEL_DT)
    miss["sig"] = rng.integers(1, 2**32, 1000, dtype=np.uint64); miss["hash"] = rng.integers(0, 2**32, 1000, dtype=np.uint64)
    sel = np.concatenate([sel1, sel2, miss]); sel = sel[sel["sig"] != 0]
    r = RefLib(algo, mem_p); r.load(t.reshape(-1))
    idx, rows = sparse(t.reshape(-1))
    save(f"ref_search_sparse_{algo}_{mem_p}", kind=np.array("search_sparse"), algo=np.array(algo), mem_p=np.array(mem_p),
         bucket_idx=idx, bucket_rows=rows, sel=sel.view(np.uint32), out=r.search(sel))


def case_serial_sparse(algo, mem_p, seed, nkeys=6000):
    """gpu_hash_insert with ONE request per launch at a larger geometry: the keys' first buckets are 64 per block (512 buckets,
    4096 slots for 6000 keys), so bucket 1 overflows, alternates spread over the whole block, and the cuckoo chains / two-choice
    overwrites run under an 11 / 17-bit BLOCK_HASH_MASK; then a delete launch per key for a third of them"""
    rng = np.random.default_rng(seed)
    nb = 1 << (mem_p - 6)
    r = RefLib(algo, mem_p)
    iel = reqs(rng, nkeys)
    block = rng.integers(0, 8, nkeys).astype(np.uint32) * np.uint32(nb // 8)
    iel["hash"] = (block | rng.integers(0, 64, nkeys).astype(np.uint32)) | (rng.integers(0, 2**32 >> (mem_p - 6), nkeys) << (mem_p - 6)).astype(np.uint32)
    # alternates must collide too, or nothing is ever evicted: keep the low signature bits small so bucket 2 stays near bucket 1
    iel["sig"] = (iel["sig"] & np.uint32(~((1 << (mem_p - 9)) - 1) & 0xFFFFFFFF)) | rng.integers(1, 128, nkeys).astype(np.uint32)
    r.insert_one_by_one(iel)
    after_insert = r.dump()
    dele = np.ascontiguousarray(iel[::3])
    for k in range(len(dele)):
        r.delete(dele[k:k + 1])
    probe = np.concatenate([to_sel(iel), to_sel(reqs(rng, 200))])
    i1, r1 = sparse(after_insert); i2, r2 = sparse(r.dump())
    save(f"ref_serial_sparse_{algo}_{mem_p}", kind=np.array("serial_sparse"), algo=np.array(algo), mem_p=np.array(mem_p),
         iel=iel.view(np.uint32), dele=dele.view(np.uint32), sel=probe.view(np.uint32),
         insert_idx=i1, insert_rows=r1, final_idx=i2, final_rows=r2, out=r.search(probe))


if __name__ == "__main__":
    mk.require_gpu()
    only = sys.argv[1:] or ["search", "delete", "batch", "serial"]
    if "search" in only:
        case_search("cuckoo", 16, 1)
    if "delete" in only:
        case_delete("cuckoo", 16, 2)
    if "batch" in only:
        case_batch("cuckoo", 20, 3); case_batch("2choice", 20, 4)
    if "serial" in only:
        case_serial("cuckoo", 16, 0.5, 5, "half"); case_serial("cuckoo", 16, 0.97, 6, "full")
        case_serial("2choice", 16, 0.97, 7, "full")
    if "sparse" in only or not sys.argv[1:]:
        for mp in (20, 26):
            case_search_sparse("cuckoo", mp, 10 + mp); case_search_sparse("2choice", mp, 11 + mp)
            case_serial_sparse("cuckoo", mp, 12 + mp); case_serial_sparse("2choice", mp, 13 + mp)
    print("done")
