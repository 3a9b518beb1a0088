#!/usr/bin/env python
"""One insert launch and one delete launch of 2^22 requests on a 2^32 B table at load factor 0.25 -- run under
ncu (lts__t_requests_srcunit_tex.sum, dram bytes, duration) with GPUHASH_UPDATE_PAIR=1 and =0 to count L2 requests per
update with two lanes per request and with one thread per request."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megakv_b200 as mk
from megakv_b200 import _native as N

L = mk.lib()
mem_p, n = 32, 1 << 22
t = mk.DeviceTable(mem_p)
pop = (1 << mem_p) // 8 // 4
buf = mk.DeviceBuffer(12 * (1 << 24))
for first in range(0, pop, 1 << 24):                       # preload: 8 launches
    N.check(L.gpuhash_gen_inserts(buf.ptr, None, 1, first, 1 << 24, None))
    N.check(L.gpuhash_insert_flat_ex(C.byref(t.geom), t.ptr, buf.ptr, 1 << 24, None, 0, None))
N.check(L.gpuhash_gen_inserts(buf.ptr, None, 1, pop, n, None))
N.check(L.gpuhash_device_sync())
N.check(L.gpuhash_insert_flat_ex(C.byref(t.geom), t.ptr, buf.ptr, n, None, 0, None))      # the measured insert launch (9th)
N.check(L.gpuhash_delete_ex(C.byref(t.geom), buf.ptr, t.ptr, n, None, 0, None))           # the measured delete launch
N.check(L.gpuhash_device_sync())
print("ok")
