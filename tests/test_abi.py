"""CPU suite: the drop-in boundary.  The library loads, exports exactly what include/*.h declares,
links the way the reference links libgpuhash (plain gcc, no libstdc++), and refuses to work without a GPU
instead of falling back to anything."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import megakv_b200 as mk
from megakv_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
CUDA_INC = "/usr/local/cuda/include"
CUDA_LIB = "/usr/local/cuda/lib64"
REF = "/root/reference"


def declared_functions(header):
    src = open(os.path.join(INC, header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*?(?<!\\)$", "", src, flags=re.M)          # drop preprocessor lines (incl. macros)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return {n for n in names if n.startswith(("gpu_", "gpuhash_"))}


def test_library_exports_every_declared_symbol(native):
    declared = declared_functions("libgpuhash.h") | declared_functions("gpuhash_ex.h")
    assert {"gpu_hash_search", "gpu_hash_insert", "gpu_hash_delete", "gpu_delete_insert"} <= declared
    assert len(declared) > 40
    raw = C.CDLL(N.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} is declared in include/ but not exported by libgpuhash.so"
    assert declared == set(N.SYMBOLS), f"binding table out of sync: {declared ^ set(N.SYMBOLS)}"


def test_static_archive_links_like_the_reference(native):
    """src/Makefile:26 links `gcc ... -lgpuhash -lcudart` with no libstdc++ / libm: the archive may only need
    libc and the CUDA runtime."""
    a = os.path.join(ROOT, "megakv_b200", "lib", "libgpuhash.a")
    assert os.path.exists(a)
    und = subprocess.check_output(["nm", "-u", a], text=True).split()
    und = {u for u in und if u != "U" and not u.endswith(":")}
    bad = {u for u in und if u.startswith(("_Z", "__cxa", "__gxx", "_Unwind")) or u in ("pow", "exp", "log")}
    assert not bad, f"C++ runtime / libm symbols needed: {bad}"


C_PROBE = r"""
#include <stdio.h>
#include "libgpuhash.h"
int main(void) {
    printf("%zu %zu %zu %zu %d %d %d %d %llu %d %d\n", sizeof(bucket_t), sizeof(selem_t), sizeof(ielem_t),
           sizeof(delem_t), (int)ELEM_NUM, (int)INSERT_BLOCK, (int)HASH_MASK, (int)BLOCK_HASH_MASK,
           (unsigned long long)HT_SIZE, (int)BUC_NUM,
#ifdef HASH_CUCKOO
           MAX_CUCKOO_NUM
#else
           -1
#endif
    );
    return 0;
}
"""


@pytest.mark.parametrize("flags,expect", [
    ([], "64 8 12 12 8 8 16777215 2097151 1073741824 16777216 5"),                       # reference defaults, MEM_P 30
    (["-DMEM_P=34", "-DHASH_2CHOICE"], "64 8 12 12 8 8 268435455 33554431 17179869184 268435456 -1"),
    (["-DMEM_P=26"], "64 8 12 12 8 8 1048575 131071 67108864 1048576 5"),
])
def test_headers_are_plain_c_with_reference_values(tmp_path, flags, expect):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("no CUDA headers")
    src = tmp_path / "probe.c"
    src.write_text(C_PROBE)
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=gnu99", "-Wall", "-Werror", "-I", INC, "-I", CUDA_INC, *flags, str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)], text=True).strip() == expect


def test_reference_test_program_links_unchanged(native, tmp_path):
    """The reference's own libgpuhash/test/insert_test.c, compiled against OUR headers and linked against OUR
    archive with the reference's link line (libgpuhash/test/Makefile:5,105).  Link only -- it needs a GPU to run."""
    srcf = os.path.join(REF, "libgpuhash", "test", "insert_test.c")
    if not os.path.exists(srcf) or not os.path.exists(os.path.join(CUDA_LIB, "libcudart.so")):
        pytest.skip("reference tree or CUDA runtime not present")
    obj, exe = tmp_path / "insert_test.o", tmp_path / "run"
    subprocess.check_call(["gcc", "-g", "-w", "-I", INC, "-I", CUDA_INC, "-c", srcf, "-o", str(obj)])
    subprocess.check_call(["gcc", "-g", str(obj), "-o", str(exe), "-lrt", "-lpthread",
                           "-L", os.path.join(ROOT, "megakv_b200", "lib"), "-lgpuhash", "-L", CUDA_LIB, "-lcudart"])
    assert os.path.getsize(exe) > 0


def test_geometry_api(native):
    g = mk.make_geom(34)
    assert (g.hash_mask, g.block_mask, g.algo, g.max_cuckoo) == (0x0FFFFFFF, 0x01FFFFFF, 0, 5)
    assert native.gpuhash_table_bytes(C.byref(g)) == 1 << 34
    s = mk.make_geom(30, mk.TWO_CHOICE, log2_shards=3)                         # 1/8 of a 1 GiB logical table
    assert (s.hash_mask, s.block_mask, s.algo) == ((1 << 21) - 1, (1 << 21) - 1, 1)
    assert native.gpuhash_table_bytes(C.byref(s)) == 1 << 27
    bad = N.Geom()
    assert native.gpuhash_geom_init_shard(C.byref(bad), 30, 4, 0) == -1          # > 8 shards are not closed
    assert native.gpuhash_geom_init(C.byref(bad), 8, 0) == -1
    d = N.Geom(); native.gpuhash_get_default_geom(C.byref(d))
    assert (d.hash_mask, d.block_mask, d.algo) == ((1 << 24) - 1, (1 << 21) - 1, 0)   # gpu_hash.h defaults


def test_no_cpu_fallback(native):
    if native.gpuhash_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(mk.GpuHashError):
        mk.require_gpu()
    with pytest.raises(mk.GpuHashError):
        mk.DeviceTable(16)
    with pytest.raises(mk.GpuHashError):
        mk.GpuHashIndex(16)
    g = mk.make_geom(16)
    assert native.gpuhash_search_ex(C.byref(g), 8, 8, 8, 1, None, None) != 0     # launch fails, no silent success


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under megakv_b200/ or include/ may mention it."""
    for base in ("megakv_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "pyoracle" not in text and "liboracle" not in text and "gpuhash_oracle.h" not in text, f
    so = subprocess.check_output(["nm", "-D", N.LIB_PATH], text=True)
    assert "orc_" not in so


def test_keystream_matches_oracle_generator():
    from megakv_b200 import keystream as ks
    from oracle import pyoracle as po
    for first, n in [(0, 5000), (123456789, 777)]:
        iel, sel = ks.uniform_inserts(1, first, n)
        oi, os_ = po.keys(1, first, n)
        assert np.array_equal(iel, oi) and np.array_equal(sel, os_)
    rng = np.random.default_rng(3)
    sel, idx = ks.zipf_queries(1, 100000, 50000, 0.99, rng)
    all_iel, _ = ks.uniform_inserts(1, 0, 100000)
    assert np.array_equal(sel["sig"], all_iel["sig"][idx]) and np.array_equal(sel["hash"], all_iel["hash"][idx])
    counts = np.bincount(idx, minlength=100000)
    assert counts[0] > counts[10] > counts[1000] and counts[0] > 0.05 * 50000      # skewed the right way


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver reads ONE JSON line from bench.py's stdout; library chatter (NCCL prints its version there) must not
    join it.  The reference arm runs on the host cores, so the contract can be checked here."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--mem-p", "24"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mops/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_zetan_closed_form_matches_the_sum():
    from megakv_b200 import keystream as ks
    n = 3_000_000
    exact = float(np.sum(1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), 0.99)))
    assert abs(ks.zetan(n, 0.99) - exact) / exact < 1e-10
    assert abs(ks.zetan(1000, 0.5) - float(np.sum(1.0 / np.sqrt(np.arange(1, 1001))))) < 1e-9


def test_c_host_example_compiles_and_links_statically(native, tmp_path):
    """examples/scheduler_cycle.c -- a C99 host using the legacy ABI, gpuhash_index_submit and gpuhash_ring_submit the way
    Mega-KV's scheduler would -- builds with plain gcc against the static archive and cudart only (the reference's link
    line, src/Makefile:26,67-75: no libstdc++)."""
    if not os.path.exists(os.path.join(CUDA_LIB, "libcudart.so")):
        pytest.skip("CUDA runtime not present")
    exe = tmp_path / "scheduler_cycle"
    subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-Wall", "-Werror", "-I", INC, "-I", CUDA_INC,
                           os.path.join(ROOT, "examples", "scheduler_cycle.c"),
                           os.path.join(ROOT, "megakv_b200", "lib", "libgpuhash.a"),
                           "-L", CUDA_LIB, "-lcudart", "-lrt", "-lpthread", "-ldl", "-o", str(exe)])
    needed = subprocess.check_output(["ldd", str(exe)], text=True)
    assert "libstdc++" not in needed and "libgpuhash" not in needed
