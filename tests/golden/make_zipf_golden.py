#!/usr/bin/env python
"""Records what the reference's own zipf generator returns (src/zipf.h compiled where it lies by `make -C oracle
_ref/libzipf_ref.so`) into tests/golden/zipf_ref.npz.  Run in the build container (needs /root/reference); the tests
only read the committed .npz."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
L = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libzipf_ref.so"))
L.ref_zipf_state_bytes.restype = C.c_size_t
L.ref_zipf_init.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64]
L.ref_zipf_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
L.ref_zipf_zetan.restype = C.c_double
L.ref_zipf_zetan.argtypes = [C.c_void_p]

cases = [(1000, 0.99, 0), (1 << 20, 0.99, 12345), (1 << 20, 0.5, 7), (1 << 24, 0.99, 99), (65536, 0.0, 3), (3, 0.9, 1)]
out = {"cases": np.array(cases, dtype=np.float64)}
for k, (n, theta, seed) in enumerate(cases):
    st = C.create_string_buffer(L.ref_zipf_state_bytes())
    L.ref_zipf_init(st, n, theta, seed)
    ranks = np.empty(20000, dtype=np.uint64)
    L.ref_zipf_fill(st, len(ranks), ranks.ctypes.data_as(C.c_void_p))
    out[f"ranks_{k}"] = ranks
    out[f"zetan_{k}"] = np.array([L.ref_zipf_zetan(st)])
np.savez_compressed(os.path.join(HERE, "zipf_ref.npz"), **out)
print("wrote zipf_ref.npz:", {k: (v.shape, v[:4]) for k, v in out.items() if k.startswith("ranks")})
