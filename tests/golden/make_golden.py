"""Golden vectors for the hash-index path.

The reference (pzrq/megakv) holds no golden vectors for libgpuhash, and its kernels
cannot run in the GPU-less authoring container.  These files therefore record, for small seeded
request sequences, the inputs and what the sequential restatement (oracle/gpuhash_oracle.c) returns:
search results, delete counts and the final table bytes in the reference's bucket_t layout.  They are
replayed (a) on the CPU against the oracle, so the oracle cannot drift silently, and (b) on the B200
through the C ABI (serial insert mode: slot-exact; see tests/test_gpu_parity.py).

Vectors named ref_*.npz are different: they are produced on a GPU box by the REFERENCE's own kernels
(oracle/_ref, compiled from /root/reference in legacy-warp mode) -- see tests/golden/make_ref_golden.py.

    python -m tests.golden.make_golden        # rewrites tests/golden/seq_*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402

OP_INSERT, OP_DELETE, OP_SEARCH = 0, 1, 2


def _reqs(rng, n, nb_bits, loc0):
    """requests whose hash is confined to nb_bits so that small tables fill up evenly"""
    iel = np.empty(n, dtype=po.IEL_DT)
    iel["sig"] = rng.integers(1, 2**32, n, dtype=np.uint64).astype(np.uint32)
    iel["hash"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    iel["loc"] = np.arange(loc0, loc0 + n, dtype=np.uint32)
    return iel


def _sel(iel):
    s = np.empty(len(iel), dtype=po.SEL_DT)
    s["sig"], s["hash"] = iel["sig"], iel["hash"]
    return s


CASES = {
    # name: (mem_p, algo, seed, script)   script = list of (op, count, source)
    "seq_cuckoo_fill95": (13, po.CUCKOO, 101),
    "seq_cuckoo_overfill": (12, po.CUCKOO, 102),
    "seq_2choice_fill95": (13, po.TWO_CHOICE, 103),
    "seq_2choice_overfill": (12, po.TWO_CHOICE, 104),
    "seq_cuckoo_churn": (13, po.CUCKOO, 105),
    "seq_2choice_churn": (13, po.TWO_CHOICE, 106),
    "seq_cuckoo_edge": (12, po.CUCKOO, 107),
}


def build_script(name):
    mem_p, algo, seed = CASES[name]
    rng = np.random.default_rng(seed)
    slots = (1 << mem_p) // 8
    steps = []
    if name.endswith("fill95"):
        a = _reqs(rng, int(slots * 0.95), mem_p, 1)
        steps += [(OP_INSERT, a), (OP_SEARCH, _sel(a))]
        miss = _reqs(rng, 300, mem_p, 1)
        steps += [(OP_SEARCH, _sel(miss))]
    elif name.endswith("overfill"):
        a = _reqs(rng, int(slots * 1.3), mem_p, 1)
        steps += [(OP_INSERT, a[: len(a) // 2]), (OP_INSERT, a[len(a) // 2:]), (OP_SEARCH, _sel(a))]
    elif name.endswith("churn"):
        live = _reqs(rng, int(slots * 0.9), mem_p, 1)
        steps += [(OP_INSERT, live)]
        loc0 = len(live) + 1
        for _ in range(6):
            k = slots // 10
            idx = rng.permutation(len(live))[:k]
            dele = live[idx].copy()
            dele["loc"][::7] ^= 0x8000                         # some deletes carry a stale loc: must not hit
            fresh = _reqs(rng, k, mem_p, loc0); loc0 += k
            upd = live[rng.permutation(len(live))[: k // 4]].copy()
            upd["loc"] += 1 << 20                              # re-SET of live keys: update in place
            ins = np.concatenate([fresh, upd]); rng.shuffle(ins)
            steps += [(OP_SEARCH, _sel(live[::3])), (OP_DELETE, dele), (OP_INSERT, ins)]
            keep = np.ones(len(live), bool); keep[idx] = False
            live = np.concatenate([live[keep], fresh])
        steps += [(OP_SEARCH, _sel(live))]
    elif name.endswith("edge"):
        a = _reqs(rng, 200, mem_p, 1)
        a["sig"][5] = 0; a["loc"][5] = 0                       # all-zero request: skipped (gpu_hash.cu:259)
        a["sig"][9] = 0                                        # sig 0, loc != 0: loc lands in an empty slot
        a[20] = a[10]; a["loc"][20] = 7777                     # duplicate key in one batch: last one wins
        a["sig"][30:40] = (a["sig"][30:40] & ~np.uint32(7)) | 3   # same major location
        a["hash"][30:40] = a["hash"][30]                       # ... in one bucket: circular slot claims
        a["sig"][50] &= ~np.uint32((1 << (mem_p - 9)) - 1)     # alt bucket == bucket 1
        steps += [(OP_INSERT, a), (OP_SEARCH, _sel(a)),
                  (OP_SEARCH, np.array([(0, int(a["hash"][9]))], dtype=po.SEL_DT)),
                  (OP_DELETE, a[::2]), (OP_DELETE, a[::4]), (OP_SEARCH, _sel(a))]
        big = _reqs(rng, int(slots * 1.1), mem_p, 10000)
        steps += [(OP_INSERT, big), (OP_SEARCH, _sel(big)), (OP_DELETE, big[100:400]), (OP_SEARCH, _sel(big))]
    return mem_p, algo, steps


def replay_oracle(mem_p, algo, steps):
    o = po.Oracle(mem_p, algo)
    res = []
    for op, arr in steps:
        if op == OP_INSERT:
            o.insert(arr); res.append(np.zeros(0, np.uint32))
        elif op == OP_DELETE:
            res.append(np.array([o.delete(arr)], dtype=np.uint64))
        else:
            res.append(o.search(arr))
    return res, o.table.copy(), o.stats.as_dict()


def pack(name):
    mem_p, algo, steps = build_script(name)
    res, table, stats = replay_oracle(mem_p, algo, steps)
    d = {"case": np.array(name), "mem_p": np.array(mem_p), "algo": np.array(algo),
         "ops": np.array([op for op, _ in steps], dtype=np.int32), "table": table,
         "stats": np.array([stats[k] for k in ("skipped", "updated", "placed_b1", "placed_b2", "to_b2",
                                                "displaced", "dropped", "overwritten")], dtype=np.uint64)}
    for i, ((_, arr), r) in enumerate(zip(steps, res)):
        d[f"in{i}"] = np.ascontiguousarray(arr).view(np.uint32)
        d[f"res{i}"] = r
    return d


def unpack_steps(g):
    steps = []
    for i, op in enumerate(g["ops"]):
        raw = g[f"in{i}"]
        arr = raw.view(po.SEL_DT) if op == OP_SEARCH else raw.view(po.IEL_DT)
        steps.append((int(op), arr))
    return int(g["mem_p"]), int(g["algo"]), steps


def run_case(name):
    """what the oracle produces NOW for the committed inputs of `name` (same keys as the .npz)"""
    g = np.load(os.path.join(HERE, name + ".npz"))
    mem_p, algo, steps = unpack_steps(g)
    res, table, stats = replay_oracle(mem_p, algo, steps)
    d = {"mem_p": np.array(mem_p), "algo": np.array(algo), "ops": g["ops"], "table": table,
         "stats": np.array([stats[k] for k in ("skipped", "updated", "placed_b1", "placed_b2", "to_b2",
                                                "displaced", "dropped", "overwritten")], dtype=np.uint64)}
    for i, ((_, arr), r) in enumerate(zip(steps, res)):
        d[f"in{i}"] = g[f"in{i}"]
        d[f"res{i}"] = r
    return d


if __name__ == "__main__":
    for name in CASES:
        d = pack(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "steps", len(d["ops"]), "table bytes", d["table"].nbytes, "stats", d["stats"].tolist())
