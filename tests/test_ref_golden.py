"""Vectors produced by the REFERENCE's own kernels on a B200 (tests/golden/ref_*.npz, made by
tests/golden/make_ref_golden.py from oracle/_ref = pzrq/megakv libgpuhash/gpu_hash.cu compiled where it lies).

  * CPU (not gpu): the oracle must reproduce them  -> the oracle is pinned to the reference itself;
  * GPU:           the CUDA path must reproduce them through the C ABI.

Nothing here needs /root/reference or oracle/_ref at run time: inputs and outputs are in the .npz files.
"""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
ALGO = {"cuckoo": po.CUCKOO, "2choice": po.TWO_CHOICE}


def load(path):
    g = np.load(path)
    return g, str(g["kind"]), ALGO[str(g["algo"])], int(g["mem_p"])


def dense(mem_p, idx, rows):
    """a sparse vector (non-empty buckets only: MEM_P 20 and 26 tables are mostly zeros) as the full table, reference layout"""
    t = np.zeros(((1 << mem_p) // 64, 16), dtype=np.uint32)
    t[idx] = rows
    return t.reshape(-1)


def test_reference_vectors_are_present():
    kinds = {str(np.load(f)["kind"]) for f in FILES}
    assert {"search", "delete"} <= kinds, "reference-kernel vectors missing (tests/golden/make_ref_golden.py)"


# ------------------------------------------------------------------ CPU: oracle vs reference

@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_reference_kernels(path):
    g, kind, algo, mem_p = load(path)
    if kind == "search":
        o = po.Oracle(mem_p, algo, table=g["table"])
        assert np.array_equal(o.search(g["sel"].view(po.SEL_DT)), g["out"])
        d = po.Oracle(mem_p, algo, table=g["dup_table"])                  # duplicate signatures in one bucket
        assert np.array_equal(d.search(g["dup_sel"].view(po.SEL_DT)), g["dup_out"])
    elif kind == "delete":
        o = po.Oracle(mem_p, algo, table=g["table_in"])
        o.delete(g["dele"].view(po.IEL_DT))
        assert np.array_equal(o.table, g["table"])
    elif kind == "serial":                                              # one request per reference launch
        o = po.Oracle(mem_p, algo)
        o.insert(g["iel"].view(po.IEL_DT))
        assert np.array_equal(o.table, g["table"]), "table bytes after a sequential reference run"
        assert np.array_equal(o.search(g["sel"].view(po.SEL_DT)), g["out"])
    elif kind == "batch":                                               # insert_test.c scenario, 8 segments per launch
        o = po.Oracle(mem_p, algo)
        for it in range(len(g["batches"])):
            iel = g["batches"][it].view(po.IEL_DT)
            n = len(iel)
            o.insert_blocks([iel[k * (n // 8):(k + 1) * (n // 8)] for k in range(8)])
            sel = np.empty(n, dtype=po.SEL_DT); sel["sig"], sel["hash"] = iel["sig"], iel["hash"]
            got_ref = g["outs"][it].reshape(-1, 2)
            mine = o.search(sel).reshape(-1, 2)
            assert np.array_equal(np.sort(mine, 1), np.sort(got_ref, 1))  # slot races inside a launch may swap b1/b2
            if it % 2:
                o.delete(iel[: n // 2])
            assert o.digest(table=g["tables"][it]) == o.digest()
    elif kind == "search_sparse":                                       # BLOCK_HASH_MASK of 11 / 17 bits through the reference's kernel
        o = po.Oracle(mem_p, algo, table=dense(mem_p, g["bucket_idx"], g["bucket_rows"]))
        assert np.array_equal(o.search(g["sel"].view(po.SEL_DT)), g["out"])
    elif kind == "serial_sparse":                                       # one request per reference launch, overflowing first buckets
        o = po.Oracle(mem_p, algo)
        o.insert(g["iel"].view(po.IEL_DT))
        assert np.array_equal(o.table, dense(mem_p, g["insert_idx"], g["insert_rows"])), "table bytes after the sequential inserts"
        o.delete(g["dele"].view(po.IEL_DT))
        assert np.array_equal(o.table, dense(mem_p, g["final_idx"], g["final_rows"])), "table bytes after the deletes"
        assert np.array_equal(o.search(g["sel"].view(po.SEL_DT)), g["out"])
    else:
        pytest.fail(f"unknown kind {kind}")


# ------------------------------------------------------------------ GPU: CUDA path vs reference

@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1], ids=["pairs", "reflayout"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_path_reproduces_reference_kernels(gpu, path, layout):
    import megakv_b200 as mk
    from tests.test_gpu_parity import gpu_search, gpu_insert, gpu_delete
    g, kind, algo, mem_p = load(path)
    t = mk.DeviceTable(mem_p, algo, layout)
    if kind == "search":
        t.load_reference(g["table"])
        assert np.array_equal(gpu_search(t, g["sel"].view(mk.SEL_DT), prezero=False), g["out"])
        t.load_reference(g["dup_table"])
        assert np.array_equal(gpu_search(t, g["dup_sel"].view(mk.SEL_DT)), g["dup_out"])
    elif kind == "delete":
        t.load_reference(g["table_in"])
        gpu_delete(t, g["dele"].view(mk.IEL_DT))
        assert np.array_equal(t.dump_reference(), g["table"])
    elif kind == "serial":
        gpu_insert(t, g["iel"].view(mk.IEL_DT), flags=mk.INSERT_SERIAL)
        assert np.array_equal(t.dump_reference(), g["table"])
        assert np.array_equal(gpu_search(t, g["sel"].view(mk.SEL_DT), prezero=False), g["out"])
    elif kind == "search_sparse":
        t.load_reference(dense(mem_p, g["bucket_idx"], g["bucket_rows"]))
        assert np.array_equal(gpu_search(t, g["sel"].view(mk.SEL_DT), prezero=False), g["out"])
    elif kind == "serial_sparse":
        gpu_insert(t, g["iel"].view(mk.IEL_DT), flags=mk.INSERT_SERIAL)
        assert np.array_equal(t.dump_reference(), dense(mem_p, g["insert_idx"], g["insert_rows"]))
        gpu_delete(t, g["dele"].view(mk.IEL_DT))
        assert np.array_equal(t.dump_reference(), dense(mem_p, g["final_idx"], g["final_rows"]))
        assert np.array_equal(gpu_search(t, g["sel"].view(mk.SEL_DT), prezero=False), g["out"])
    elif kind == "batch":
        o = po.Oracle(mem_p, algo)
        for it in range(len(g["batches"])):
            iel = g["batches"][it].view(mk.IEL_DT)
            n = len(iel)
            gpu_insert(t, iel)
            sel = np.empty(n, dtype=mk.SEL_DT); sel["sig"], sel["hash"] = iel["sig"], iel["hash"]
            assert np.array_equal(np.sort(gpu_search(t, sel).reshape(-1, 2), 1), np.sort(g["outs"][it].reshape(-1, 2), 1))
            if it % 2:
                gpu_delete(t, iel[: n // 2])
            assert o.digest(table=t.dump_reference()) == o.digest(table=g["tables"][it])
