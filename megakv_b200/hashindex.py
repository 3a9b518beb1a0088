"""Host-side mirror of the reference's libgpuhash interface, on top of the C ABI.

Two levels, both thin:

* ``DeviceTable`` + ``gpu_hash_search / gpu_hash_insert / gpu_hash_delete`` -- the legacy
  device-pointer calls with the reference's names and argument order
  (libgpuhash/libgpuhash.h:29-51), driven the way libgpuhash/test/insert_test.c drives them:
  caller-owned device buffers, explicit copies.
* ``GpuHashIndex`` -- the scheduler-cycle object (src/mega_scheduler.c:392-504): host batches
  in, host results out.

numpy arrays use the reference's record layouts: selem_t = (sig, hash) u32 pairs,
ielem_t/delem_t = (sig, hash, loc) u32 triples (gpu_hash.h:85-104).
"""
import ctypes as C

import numpy as np

from . import _native as N

SEL_DT = np.dtype([("sig", "<u4"), ("hash", "<u4")])
IEL_DT = np.dtype([("sig", "<u4"), ("hash", "<u4"), ("loc", "<u4")])
INSERT_BLOCK = 8      # gpu_hash.h:68


def _p(a):
    return C.c_void_p(a.ctypes.data)


def make_geom(mem_p, algo=N.CUCKOO, log2_shards=0, layout=N.LAYOUT_PAIRS):
    g = N.Geom()
    N.check(N.lib().gpuhash_geom_init_shard(C.byref(g), mem_p, log2_shards, algo), "gpuhash_geom_init")
    g.layout = layout
    return g


class DeviceBuffer:
    """cudaMalloc'd bytes (the reference's tests call cudaMalloc/cudaMemcpy inline)."""

    def __init__(self, nbytes, zero=False):
        N.require_gpu()
        self.nbytes = int(nbytes)
        self.ptr = N.lib().gpuhash_dev_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise N.GpuHashError(f"cudaMalloc({nbytes}) failed")
        if zero:
            N.check(N.lib().gpuhash_dev_memset(self.ptr, 0, self.nbytes, None))
            N.check(N.lib().gpuhash_device_sync())

    @classmethod
    def from_host(cls, arr):
        arr = np.ascontiguousarray(arr)
        b = cls(arr.nbytes)
        b.upload(arr)
        return b

    def upload(self, arr, stream=None):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        N.check(N.lib().gpuhash_h2d(self.ptr, _p(arr), arr.nbytes, stream))
        N.check(N.lib().gpuhash_stream_sync(stream))

    def download(self, dtype=np.uint32, count=None, stream=None):
        dtype = np.dtype(dtype)
        n = self.nbytes // dtype.itemsize if count is None else count
        out = np.empty(n, dtype=dtype)
        N.check(N.lib().gpuhash_d2h(_p(out), self.ptr, out.nbytes, stream))
        N.check(N.lib().gpuhash_stream_sync(stream))
        return out

    def zero(self):
        N.check(N.lib().gpuhash_dev_memset(self.ptr, 0, self.nbytes, None))
        N.check(N.lib().gpuhash_device_sync())

    def free(self):
        if self.ptr:
            N.lib().gpuhash_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceTable(DeviceBuffer):
    """One cudaMalloc(HT_SIZE) + cudaMemset(0), as mega_scheduler.c:273-274 / insert_test.c:80-81.
    `layout` is how the library lays slots out inside the 64 B buckets (gpuhash_ex.h); host images are
    always exchanged in the reference's bucket_t layout."""

    def __init__(self, mem_p, algo=N.CUCKOO, layout=N.LAYOUT_PAIRS, log2_shards=0):
        self.geom = make_geom(mem_p, algo, log2_shards, layout)
        super().__init__(N.lib().gpuhash_table_bytes(C.byref(self.geom)), zero=True)

    def load_reference(self, words):
        """upload a table image in the reference byte layout (bucket_t[], gpu_hash.h:79-82)"""
        self.upload(np.ascontiguousarray(words).view(np.uint32))
        if self.geom.layout != N.LAYOUT_REFERENCE:
            as_ref = N.Geom.from_buffer_copy(bytes(self.geom)); as_ref.layout = N.LAYOUT_REFERENCE
            N.check(N.lib().gpuhash_table_convert(C.byref(as_ref), self.ptr, self.geom.layout, None))
            N.check(N.lib().gpuhash_device_sync())

    def dump_reference(self):
        """the table as the reference would hold it (converted on the device, table left unchanged)"""
        L = N.lib()
        N.check(L.gpuhash_device_sync())
        if self.geom.layout == N.LAYOUT_REFERENCE:
            return self.download(np.uint32)
        N.check(L.gpuhash_table_convert(C.byref(self.geom), self.ptr, N.LAYOUT_REFERENCE, None))
        out = self.download(np.uint32)
        as_ref = N.Geom.from_buffer_copy(bytes(self.geom)); as_ref.layout = N.LAYOUT_REFERENCE
        N.check(L.gpuhash_table_convert(C.byref(as_ref), self.ptr, self.geom.layout, None))
        N.check(L.gpuhash_device_sync())
        return out

    def make_default(self):
        """Make this geometry the one the three legacy entry points use."""
        N.lib().gpuhash_set_default_geom(C.byref(self.geom))


# ---- legacy calls, reference names and argument order (device pointers, async) ----

def gpu_hash_search(in_d, out_d, hash_table, num_elem, num_thread=24576, threads_per_blk=256, stream=None):
    N.lib().gpu_hash_search(in_d.ptr, out_d.ptr, hash_table.ptr, num_elem, num_thread, threads_per_blk, stream)


def gpu_hash_insert(hash_table, blk_input_d, blk_elem_num_d, num_blks, stream=None):
    N.lib().gpu_hash_insert(hash_table.ptr, blk_input_d.ptr, blk_elem_num_d.ptr, num_blks, stream)


def gpu_hash_delete(in_d, hash_table, num_elem, num_thread=16384, threads_per_blk=256, stream=None):
    N.lib().gpu_hash_delete(in_d.ptr, hash_table.ptr, num_elem, num_thread, threads_per_blk, stream)


def device_sync():
    N.check(N.lib().gpuhash_device_sync(), "cudaDeviceSynchronize")


class InsertSegments:
    """Device copy of an insert batch the way the scheduler lays it out: num_blks sub-buffers, a device
    array of their device pointers and a device array of their lengths (mega_recv.c:132-150,
    mega_scheduler.c:484-494)."""

    def __init__(self, blocks):
        blocks = [np.ascontiguousarray(b, dtype=IEL_DT) for b in blocks]
        self.num_blks = len(blocks)
        self.bufs = [DeviceBuffer.from_host(b) if len(b) else DeviceBuffer(12) for b in blocks]
        self.ptrs = DeviceBuffer.from_host(np.array([b.ptr for b in self.bufs], dtype=np.uint64))
        self.nums = DeviceBuffer.from_host(np.array([len(b) for b in blocks], dtype=np.int32))


def split_insert_blocks(iel, num_blks=INSERT_BLOCK):
    """The receiver's partition of an insert batch: block id = top bits of hash (mega_recv.c:476-477)."""
    iel = np.ascontiguousarray(iel, dtype=IEL_DT)
    bits = int(num_blks).bit_length() - 1
    if bits == 0:
        return [iel]
    blk = iel["hash"] >> np.uint32(32 - bits)
    return [iel[blk == k] for k in range(num_blks)]


# ---- the scheduler-cycle object ----

class GpuHashIndex:
    """Table + per-worker streams and staging; ``cycle`` is one pass of mega_scheduler.c:392-504."""

    def __init__(self, mem_p, algo=N.CUCKOO, workers=1, max_search=1 << 16, max_insert=1 << 16, max_delete=1 << 16,
                 layout=N.LAYOUT_PAIRS):
        N.require_gpu()
        self.L = N.lib()
        self.mem_p, self.algo, self.workers = mem_p, algo, workers
        self.h = self.L.gpuhash_index_create_layout(mem_p, algo, layout, workers, max_search, max_insert, max_delete)
        if not self.h:
            raise N.GpuHashError("gpuhash_index_create failed (out of device memory?)")
        self.geom = self.L.gpuhash_index_geom(self.h).contents
        self.table_bytes = self.L.gpuhash_table_bytes(C.byref(self.geom))
        self._keep = []

    def close(self):
        if self.h:
            self.L.gpuhash_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def table_ptr(self):
        return self.L.gpuhash_index_table(self.h)

    def submit(self, worker=0, search=None, delete=None, insert=None):
        """Asynchronous; returns the (not yet valid) result array of the search part."""
        s = np.ascontiguousarray(search, dtype=SEL_DT) if search is not None else np.empty(0, SEL_DT)
        d = np.ascontiguousarray(delete, dtype=IEL_DT) if delete is not None else np.empty(0, IEL_DT)
        i = np.ascontiguousarray(insert, dtype=IEL_DT) if insert is not None else np.empty(0, IEL_DT)
        out = np.empty(2 * len(s), dtype=np.uint32)
        self._keep.append((s, d, i, out))
        N.check(self.L.gpuhash_index_submit(self.h, worker, _p(s), len(s), _p(out), _p(d), len(d), _p(i), len(i)),
                "gpuhash_index_submit")
        return out

    def sync(self):
        N.check(self.L.gpuhash_index_sync(self.h), "gpuhash_index_sync")
        self._keep.clear()

    def submit_all(self, batches):
        """One scheduler cycle of ALL workers in ONE launch (mega_scheduler.c:393-504): batches[w] = dict with optional
        'search', 'delete', 'insert' arrays of worker w.  Returns (ticket, [result array per worker]); results are valid
        after wait(ticket) or sync()."""
        descs = (N.Batch * len(batches))()
        outs = []
        for w, b in enumerate(batches):
            s = np.ascontiguousarray(b.get("search"), dtype=SEL_DT) if b.get("search") is not None else np.empty(0, SEL_DT)
            d = np.ascontiguousarray(b.get("delete"), dtype=IEL_DT) if b.get("delete") is not None else np.empty(0, IEL_DT)
            i = np.ascontiguousarray(b.get("insert"), dtype=IEL_DT) if b.get("insert") is not None else np.empty(0, IEL_DT)
            out = np.zeros(2 * len(s), dtype=np.uint32)
            self._keep.append((s, d, i, out))
            descs[w] = N.Batch(s.ctypes.data if len(s) else None, out.ctypes.data if len(s) else None,
                               d.ctypes.data if len(d) else None, i.ctypes.data if len(i) else None, len(s), len(d), len(i), 0)
            outs.append(out)
        ticket = self.L.gpuhash_index_submit_all(self.h, descs, len(batches))
        if ticket < 0:
            raise N.GpuHashError(f"gpuhash_index_submit_all failed: {ticket}")
        return ticket, outs

    def wait(self, ticket):
        N.check(self.L.gpuhash_index_wait(self.h, ticket), "gpuhash_index_wait")

    def cycle(self, search=None, delete=None, insert=None, worker=0):
        out = self.submit(worker, search, delete, insert)
        self.sync()
        return out

    def search(self, sel):
        return self.cycle(search=sel)

    def insert(self, iel):
        self.cycle(insert=iel)

    def delete(self, iel):
        self.cycle(delete=iel)

    def clear(self):
        N.check(self.L.gpuhash_index_clear(self.h))

    def load(self, table_words):
        t = np.ascontiguousarray(table_words).view(np.uint32)
        assert t.nbytes == self.table_bytes
        N.check(self.L.gpuhash_index_load(self.h, _p(t)))

    def dump(self):
        t = np.empty(self.table_bytes // 4, dtype=np.uint32)
        N.check(self.L.gpuhash_index_dump(self.h, _p(t)))
        return t

    def enable_stats(self, on=True):
        self.L.gpuhash_index_enable_stats(self.h, int(on))

    def stats(self, reset=False):
        st = N.Stats()
        N.check(self.L.gpuhash_index_stats(self.h, C.byref(st), int(reset)))
        return st.as_dict()


# ---- run-time-geometry launches on caller-owned device buffers (gpuhash_ex.h) ----

class DeviceStats(DeviceBuffer):
    def __init__(self):
        super().__init__(C.sizeof(N.Stats), zero=True)

    def read(self):
        raw = self.download(np.uint8)
        return N.Stats.from_buffer_copy(raw.tobytes()).as_dict()


def search_ex(geom, in_d, out_d, table, n, stats=None, stream=None):
    N.check(N.lib().gpuhash_search_ex(C.byref(geom), in_d.ptr, out_d.ptr, table.ptr, n,
                                      stats.ptr if stats else None, stream), "gpuhash_search_ex")


def insert_flat_ex(geom, table, in_d, n, stats=None, flags=0, stream=None):
    N.check(N.lib().gpuhash_insert_flat_ex(C.byref(geom), table.ptr, in_d.ptr, n,
                                           stats.ptr if stats else None, flags, stream), "gpuhash_insert_flat_ex")


def insert_ex(geom, table, segs, stats=None, flags=0, stream=None):
    N.check(N.lib().gpuhash_insert_ex(C.byref(geom), table.ptr, segs.ptrs.ptr, segs.nums.ptr, segs.num_blks,
                                      stats.ptr if stats else None, flags, stream), "gpuhash_insert_ex")


def delete_ex(geom, in_d, table, n, stats=None, flags=0, stream=None):
    N.check(N.lib().gpuhash_delete_ex(C.byref(geom), in_d.ptr, table.ptr, n,
                                      stats.ptr if stats else None, flags, stream), "gpuhash_delete_ex")
