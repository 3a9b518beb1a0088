"""ctypes view of megakv_b200/lib/libgpuhash.so (the C ABI declared in include/*.h).

There is no fallback: if the shared library is missing, or a call is made without a CUDA
device, this module raises.  Nothing here imports oracle/.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPUHASH_LIB: another build of the SAME library (A/B runs of compile-time switches); never a different implementation
LIB_PATH = os.environ.get("GPUHASH_LIB") or os.path.join(_HERE, "lib", "libgpuhash.so")

CUCKOO, TWO_CHOICE = 0, 1
LAYOUT_PAIRS, LAYOUT_REFERENCE = 0, 1
INSERT_SERIAL = 1


class Geom(C.Structure):                      # gpuhash_geom_t
    _fields_ = [("hash_mask", C.c_uint32), ("block_mask", C.c_uint32),
                ("algo", C.c_uint32), ("max_cuckoo", C.c_uint32), ("layout", C.c_uint32)]


class Stats(C.Structure):                     # gpuhash_stats_t
    _names = ("ins_skipped", "ins_updated", "ins_placed_b1", "ins_placed_b2", "ins_to_b2",
              "ins_displaced", "ins_dropped", "ins_overwritten", "ins_cas_retry", "ins_gave_up")
    _tail = ("del_zeroed", "del_requests_hit", "search_hits_b1", "search_hits_b2")
    _fields_ = ([(n, C.c_ulonglong) for n in _names] + [("chain_hist", C.c_ulonglong * 8)]
                + [(n, C.c_ulonglong) for n in _tail])

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n in self._names + self._tail}
        d["chain_hist"] = [int(v) for v in self.chain_hist]
        return d


class Tune(C.Structure):                      # gpuhash_tune_t
    _fields_ = [("search_qpt", C.c_int), ("search_split_mode", C.c_int), ("insert_ctas_per_sm", C.c_int),
                ("fused_cycle", C.c_int)]

    def __init__(self, search_qpt=0, search_split_mode=0, insert_ctas_per_sm=4, fused_cycle=1):
        super().__init__(search_qpt, search_split_mode, insert_ctas_per_sm, fused_cycle)


class Batch(C.Structure):                     # gpuhash_batch_t
    _fields_ = [("search_in", C.c_void_p), ("search_out", C.c_void_p), ("delete_in", C.c_void_p), ("insert_in", C.c_void_p),
                ("n_search", C.c_uint32), ("n_delete", C.c_uint32), ("n_insert", C.c_uint32), ("reserved", C.c_uint32)]


class BenchResult(C.Structure):               # gpuhash_bench_result_t
    _fields_ = [("total_ms", C.c_float), ("search_ms", C.c_float), ("launches", C.c_ulonglong),
                ("search_ops", C.c_ulonglong), ("insert_ops", C.c_ulonglong), ("delete_ops", C.c_ulonglong),
                ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong)]


# every symbol include/*.h declares: name -> (restype, argtypes)
_vp, _sz, _i, _u = C.c_void_p, C.c_size_t, C.c_int, C.c_uint
_gp, _sp = C.POINTER(Geom), C.POINTER(Stats)
SYMBOLS = {
    # libgpuhash.h (legacy ABI, reference libgpuhash.h:29-62)
    "gpu_hash_search": (None, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "gpu_hash_insert": (None, [_vp, _vp, _vp, _i, _vp]),
    "gpu_hash_delete": (None, [_vp, _vp, _i, _i, _i, _vp]),
    "gpu_delete_insert": (None, [_vp, _vp, C.c_uint32, _vp, _vp, _i, C.c_uint32, C.c_uint32, _vp]),
    # gpuhash_ex.h
    "gpuhash_geom_init": (_i, [_gp, _i, _u]),
    "gpuhash_geom_init_shard": (_i, [_gp, _i, _i, _u]),
    "gpuhash_geom_init_auto": (_i, [_gp, _i, _u]),
    "gpuhash_table_bytes": (_sz, [_gp]),
    "gpuhash_table_convert": (_i, [_gp, _vp, _u, _vp]),
    "gpuhash_set_default_geom": (None, [_gp]),
    "gpuhash_get_default_geom": (None, [_gp]),
    "gpuhash_set_tuning": (None, [C.POINTER(Tune)]),
    "gpuhash_get_tuning": (None, [C.POINTER(Tune)]),
    "gpuhash_search_ex": (_i, [_gp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gpuhash_insert_ex": (_i, [_gp, _vp, _vp, _vp, _i, _vp, _u, _vp]),
    "gpuhash_insert_flat_ex": (_i, [_gp, _vp, _vp, _sz, _vp, _u, _vp]),
    "gpuhash_delete_ex": (_i, [_gp, _vp, _vp, _sz, _vp, _u, _vp]),
    "gpuhash_init_device": (_i, []),
    "gpuhash_cycle_ex": (_i, [_gp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _i, _vp, _vp]),
    "gpuhash_cycle_ws_ex": (_i, [_gp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "gpuhash_cycle_multi_ex": (_i, [_gp, _vp, C.POINTER(Batch), _vp, _i, _i, _vp, _vp, _vp]),
    "gpuhash_cycle_workspace_bytes": (_sz, [_i]),
    "gpuhash_cycle_error": (_i, [_i]),
    "gpuhash_set_cycle_ctas_per_sm": (None, [_i]),
    "gpuhash_device_count": (_i, []),
    "gpuhash_set_device": (_i, [_i]),
    "gpuhash_device_info": (_i, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz)]),
    "gpuhash_dev_alloc": (_vp, [_sz]),
    "gpuhash_dev_free": (_i, [_vp]),
    "gpuhash_dev_memset": (_i, [_vp, _i, _sz, _vp]),
    "gpuhash_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "gpuhash_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "gpuhash_host_alloc": (_vp, [_sz]),
    "gpuhash_host_free": (_i, [_vp]),
    "gpuhash_stream_create": (_vp, []),
    "gpuhash_stream_destroy": (_i, [_vp]),
    "gpuhash_stream_sync": (_i, [_vp]),
    "gpuhash_device_sync": (_i, []),
    "gpuhash_event_create": (_vp, []),
    "gpuhash_event_destroy": (_i, [_vp]),
    "gpuhash_event_record": (_i, [_vp, _vp]),
    "gpuhash_event_elapsed_ms": (_i, [_vp, _vp, C.POINTER(C.c_float)]),
    "gpuhash_set_l2_fetch_granularity": (_i, [_i]),
    "gpuhash_get_l2_fetch_granularity": (_i, []),
    "gpuhash_error_string": (C.c_char_p, [_i]),
    "gpuhash_build_info": (C.c_char_p, []),
    "gpuhash_roofline_gather": (_i, [_vp, _sz, _sz, _i, _i, _i, C.POINTER(C.c_float), _vp]),
    "gpuhash_index_create": (_vp, [_i, _u, _i, _sz, _sz, _sz]),
    "gpuhash_index_create_layout": (_vp, [_i, _u, _u, _i, _sz, _sz, _sz]),
    "gpuhash_index_destroy": (None, [_vp]),
    "gpuhash_index_table": (_vp, [_vp]),
    "gpuhash_index_geom": (_gp, [_vp]),
    "gpuhash_index_stream": (_vp, [_vp, _i]),
    "gpuhash_index_clear": (_i, [_vp]),
    "gpuhash_index_load": (_i, [_vp, _vp]),
    "gpuhash_index_dump": (_i, [_vp, _vp]),
    "gpuhash_index_stats": (_i, [_vp, _sp, _i]),
    "gpuhash_index_enable_stats": (_i, [_vp, _i]),
    "gpuhash_index_set_zero_copy": (_i, [_vp, _i]),
    "gpuhash_index_set_compact_results": (_i, [_vp, _i]),
    "gpuhash_index_submit": (_i, [_vp, _i, _vp, _sz, _vp, _vp, _sz, _vp, _sz]),
    "gpuhash_index_sync": (_i, [_vp]),
    "gpuhash_index_submit_all": (_i, [_vp, C.POINTER(Batch), _i]),
    "gpuhash_index_wait": (_i, [_vp, _i]),
    "gpuhash_index_set_unordered_cycles": (_i, [_vp, _i]),
    "gpuhash_route_scatter": (_i, [_vp, _sz, _i, C.c_uint32, _i, _vp, _vp, _vp, _sz, _vp]),
    "gpuhash_route_publish": (_i, [_vp, _i, _i, _vp, _vp, C.c_uint32, _vp]),
    "gpuhash_search_segments": (_i, [_gp, _vp, _i, _vp, _vp, _vp, _sz, _vp, C.c_uint32, _vp, _vp]),
    "gpuhash_results_publish": (_i, [_i, _i, _vp, C.c_uint32, _vp]),
    "gpuhash_route_gather": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _sz, _vp, C.c_uint32, _vp, _vp]),
    "gpuhash_delete_segments": (_i, [_gp, _vp, _i, _vp, _vp, _sz, _vp, _vp]),
    "gpuhash_bench_ring": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, C.POINTER(BenchResult), _i, C.POINTER(C.c_float)]),
    "gpuhash_search_compact_ex": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gpuhash_fold_keys_ex": (_i, [_vp, _sz, C.c_uint, _i, _sz, _vp, _vp]),
    "gpuhash_ring_create": (_vp, [_vp, _vp, _i, _i, _i, C.c_uint]),
    "gpuhash_ring_submit": (C.c_longlong, [_vp, _i, _vp, _sz, _vp, _vp, _sz, _vp, _sz]),
    "gpuhash_ring_wait": (_i, [_vp, _i, C.c_longlong, C.c_uint]),
    "gpuhash_ring_drain": (_i, [_vp, C.c_uint]),
    "gpuhash_ring_park": (_i, [_vp]),
    "gpuhash_ring_destroy": (None, [_vp]),
    "gpuhash_ring_ctas_per_ring": (_i, [_vp]),
    "gpuhash_ring_trace": (_i, [_vp, _i, C.POINTER(C.c_ulonglong)]),
    "gpuhash_route_map_bytes": (_sz, [_sz]),
    "gpuhash_route_scatter_tiles": (_i, [_vp, _sz, _i, C.c_uint32, _i, _vp, _vp, _vp, _sz, _i, _vp, _vp, _vp, C.c_uint32, _vp]),
    "gpuhash_route_gather_tiles": (_i, [_vp, _vp, _sz, _i, _vp, _sz, _vp]),
    "gpuhash_route_scatter_pub": (_i, [_vp, _sz, _i, C.c_uint32, _i, _vp, _vp, _vp, _sz, _i, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp]),
    "gpuhash_serve": (_i, [_gp, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp, _vp, _i, _vp, _vp, C.c_uint32, _vp, _vp]),
    "gpuhash_wait_flags": (_i, [_vp, _i, C.c_uint32, _vp, _vp]),
    "gpuhash_wait_mode": (_i, []),
    "gpuhash_ipc_export": (_i, [_vp, _vp]),
    "gpuhash_ipc_import": (_vp, [_vp]),
    "gpuhash_ipc_close": (_i, [_vp]),
    "gpuhash_xchg_create": (_vp, [_gp, _vp, C.c_uint32, _i, _i, _sz, _sz]),
    "gpuhash_xchg_arena": (_vp, [_vp, C.POINTER(_sz)]),
    "gpuhash_xchg_set_peers": (_i, [_vp, _vp]),
    "gpuhash_xchg_set_stats": (_i, [_vp, _vp]),
    "gpuhash_xchg_step": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _sz, _vp]),
    "gpuhash_xchg_flush": (_i, [_vp, _vp]),
    "gpuhash_xchg_seq": (C.c_uint, [_vp]),
    "gpuhash_xchg_error": (_i, [_vp]),
    "gpuhash_xchg_destroy": (None, [_vp]),
    "gpuhash_gen_inserts": (_i, [_vp, _vp, C.c_uint64, C.c_uint64, _sz, _vp]),
    "gpuhash_gen_queries": (_i, [_vp, _vp, C.c_uint64, C.c_uint64, _sz, C.c_uint64, C.c_double, C.c_double, _vp]),
    "gpuhash_gen_requests": (_i, [_vp, C.c_uint64, C.c_uint64, _sz, C.c_uint64, C.c_double, C.c_double, _vp]),
    "gpuhash_gen_requests_ref_zipf": (_i, [_vp, _vp, C.c_uint64, C.c_uint64, _sz, C.c_uint64, C.c_uint64, C.c_double, C.c_double, _i, _vp]),
    "gpuhash_bench_resident": (_i, [_gp, _vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _i, C.POINTER(BenchResult)]),
    "gpuhash_bench_e2e": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, _i, C.POINTER(BenchResult)]),
    "gpuhash_bench_cycles": (_i, [_gp, _vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _i, C.POINTER(BenchResult)]),
    "gpuhash_bench_e2e_cycles": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, _sz, _i, _i, C.POINTER(BenchResult)]),
}

_lib = None


class GpuHashError(RuntimeError):
    pass


def lib():
    """The loaded library with every prototype applied.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpuHashError(
                f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()). "
                "megakv_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)            # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().gpuhash_error_string(rc).decode()
        raise GpuHashError(f"{what or 'libgpuhash'} failed: {rc} ({msg})")


def require_gpu():
    n = lib().gpuhash_device_count()
    if n < 1:
        raise GpuHashError("no CUDA device visible; megakv_b200 has no CPU fallback")
    return n
