"""megakv_b200 -- B200-native replacement for Mega-KV's libgpuhash hot path.

The product is the C ABI in include/*.h, implemented by hand-written sm_100a CUDA under
megakv_b200/csrc/ and built into megakv_b200/lib/libgpuhash.{so,a}.  This package is the
thin host mirror used by the tests and the bench; it has no CPU fallback.
"""
from ._native import CUCKOO, TWO_CHOICE, INSERT_SERIAL, LAYOUT_PAIRS, LAYOUT_REFERENCE, GpuHashError, lib, require_gpu  # noqa: F401
from .hashindex import (  # noqa: F401
    SEL_DT, IEL_DT, INSERT_BLOCK, DeviceBuffer, DeviceTable, DeviceStats, GpuHashIndex, InsertSegments,
    gpu_hash_search, gpu_hash_insert, gpu_hash_delete, device_sync, make_geom, split_insert_blocks,
    search_ex, insert_flat_ex, insert_ex, delete_ex)
