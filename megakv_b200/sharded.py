"""Sharded hash index: one process per GPU, the logical table cut into G = world_size bucket ranges.

The reference is single-GPU (src/mega.c:410).  Sharding is exact because both candidate buckets of a key and
every bucket its eviction chain can reach share the top IBLOCK_P = 3 bits of the bucket index (gpu_hash.h:67-69,
gpu_hash.cu:66-67,334-335): rank g owns bucket range g as a local table with gpuhash_geom_init_shard geometry and
the results equal the single-table oracle's (tests/test_sharded_gloo.py, tests/test_gpu_sharded.py).

Per batch and rank:   scatter by owner -> exchange -> local kernel -> (searches) exchange back -> gather.

Two exchanges:
  "collective"  torch.distributed all_to_all_single on packed buffers (NCCL on GPUs, gloo in the CPU tests).
                Needs the split sizes on the host, i.e. one device->host sync per batch: the baseline.
  "p2p"         the fused path: the scatter kernel stores each request straight into its owner's inbox over
                NVLink, the lookup kernel stores each result straight into the origin's staging area (peer pointers
                from CUDA IPC), sequence-numbered flags replace every host sync and every collective.

torch is plumbing here (process group, NCCL, current stream); the kernels are megakv_b200/csrc/gpuhash_shard.cu.
The device work sits behind a small backend object so that the choreography below can be exercised on CPU with
a stand-in backend (tests only).
"""
import ctypes as C
import os

import numpy as np

MAX_SHARDS = 8


def log2_exact(n):
    l = int(n).bit_length() - 1
    if n < 1 or (1 << l) != n or l > 3:
        raise ValueError("world size must be 1, 2, 4 or 8 (the alternate bucket keeps 3 top bits)")
    return l


class ShardPlan:
    """Pure arithmetic of the partition (host side)."""

    def __init__(self, mem_p_total, world):
        self.mem_p_total, self.world, self.log2 = mem_p_total, world, log2_exact(world)
        self.bits = mem_p_total - 6                                 # bucket-index bits of the logical table
        self.hash_mask_total = (1 << self.bits) - 1
        self.shift = self.bits - self.log2
        self.mem_p_shard = mem_p_total - self.log2

    def owner(self, hash_):
        h = np.asarray(hash_, dtype=np.uint32)
        return ((h & np.uint32(self.hash_mask_total)) >> np.uint32(self.shift)).astype(np.int64)


class ShardedIndex:
    """search / insert / delete on the sharded table.  Requests and results are torch tensors on the backend's
    device: int32 [n, 2] (sig, hash) for searches, [n, 3] (sig, hash, loc) for inserts/deletes, results [n, 2]."""

    def __init__(self, backend, plan, group=None, exchange="collective"):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.be, self.plan = backend, plan
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world == plan.world
        self.exchange = exchange
        self.seq = 0
        if exchange == "p2p":
            backend.p2p_setup(self)

    # ---- collective exchange (baseline)
    def _a2a_counts(self, counts):
        recv = counts.new_empty(self.world)
        self.dist.all_to_all_single(recv, counts[: self.world].contiguous(), group=self.group)
        return recv

    def _route(self, req, words, want_perm):
        be = self.be
        send, counts, perm = be.scatter(req, words, want_perm)          # send: [G, cap, words]
        recv_counts = self._a2a_counts(counts)
        cs, rs = counts[: self.world].tolist(), recv_counts.tolist()    # host sync: the cost of this path
        packed = be.pack(send, cs)                                       # [sum(cs), words]
        inbox = be.empty(sum(rs), words)
        self.dist.all_to_all_single(inbox, packed, rs, cs, group=self.group)
        return inbox, cs, rs, perm

    def search(self, sel, out=None, after_serve=None):
        if self.exchange == "p2p":
            return self.be.p2p_search(self, sel, out, after_serve)
        inbox, cs, rs, perm = self._route(sel, 2, True)
        res = self.be.search_local(inbox, rs)                            # [sum(rs), 2]
        back = self.be.empty(sum(cs), 2)
        self.dist.all_to_all_single(back, res, cs, rs, group=self.group)
        return self.be.gather(back, cs, perm, sel.shape[0])

    def insert(self, iel, before_serve=None):
        if self.exchange == "p2p":
            return self.be.p2p_update(self, iel, insert=True, before_serve=before_serve)
        inbox, cs, rs, _ = self._route(iel, 3, False)
        self.be.insert_local(inbox, rs)

    def delete(self, iel):
        if self.exchange == "p2p":
            return self.be.p2p_update(self, iel, insert=False)
        inbox, cs, rs, _ = self._route(iel, 3, False)
        return self.be.delete_local(inbox, rs)


class CudaShardBackend:
    """The product backend: buffers are torch CUDA tensors (plumbing), work is libgpuhash kernels on torch's
    current stream.  `cap` = largest batch per rank."""

    def __init__(self, plan, rank, cap, algo=0, layout=0, device=None, table=None):
        import torch
        from . import _native as N
        from .hashindex import DeviceBuffer, make_geom
        self.torch, self.N, self.L = torch, N, N.lib()
        self.plan, self.rank, self.cap, self.G = plan, rank, (int(cap) + 7) & ~7, plan.world     # regions stay 32 B aligned
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.geom = make_geom(plan.mem_p_total, algo, plan.log2, layout)
        # several backends ("lanes": independent buffer sets for batches in flight) may share one shard table
        self.table = table if table is not None else DeviceBuffer(self.L.gpuhash_table_bytes(C.byref(self.geom)), zero=True)
        i32 = torch.int32
        self._send = None                                           # collective path only: allocated at first use
        self.counts = torch.zeros(MAX_SHARDS, dtype=i32, device=self.dev)
        # collective path: perm[d][slot] = request index; fused path: the tile map of gpuhash_route_scatter_tiles
        self.route_tiles = os.environ.get("GPUHASH_ROUTE_TILES", "1") != "0"
        perm_words = max(self.G * self.cap, (self.L.gpuhash_route_map_bytes(self.cap) + 3) // 4)
        self.perm = torch.empty(perm_words, dtype=i32, device=self.dev)
        self.seg_counts = torch.zeros(MAX_SHARDS, dtype=i32, device=self.dev)
        self.seg_ptrs_d = torch.zeros(MAX_SHARDS, dtype=torch.int64, device=self.dev)    # device array of region pointers
        self.p2p = None

    @property
    def send(self):
        if self._send is None:
            self._send = self.torch.empty((self.G, self.cap, 3), dtype=self.torch.int32, device=self.dev)   # widest element
        return self._send

    # -- helpers
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _ptrs(values):
        arr = (C.c_void_p * MAX_SHARDS)()
        for k, v in enumerate(values):
            arr[k] = v
        return arr

    def empty(self, n, words):
        return self.torch.empty((max(n, 0), words), dtype=self.torch.int32, device=self.dev)

    # -- collective path pieces
    def scatter(self, req, words, want_perm):
        n = req.shape[0]
        assert n <= self.cap and req.is_contiguous() and req.dtype == self.torch.int32
        base = self.send.data_ptr()
        dst = self._ptrs([base + d * self.cap * words * 4 for d in range(self.G)])
        self.N.check(self.L.gpuhash_route_scatter(req.data_ptr(), n, words, self.plan.hash_mask_total, self.plan.log2, dst,
                                                  self.counts.data_ptr(), self.perm.data_ptr() if want_perm else None,
                                                  self.cap, self._stream()), "gpuhash_route_scatter")
        view = self.send.view(-1)[: self.G * self.cap * words].view(self.G, self.cap, words)
        return view, self.counts, self.perm

    def pack(self, send, cs):
        return self.torch.cat([send[d, : cs[d]] for d in range(self.G)], dim=0).contiguous()

    def _segments(self, buf, sizes, words):
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        ptrs = [buf.data_ptr() + int(offs[s]) * words * 4 for s in range(self.G)]
        cnt = self.torch.tensor(list(sizes) + [0] * (MAX_SHARDS - self.G), dtype=self.torch.int32)
        self.seg_counts.copy_(cnt, non_blocking=False)
        return ptrs

    def search_local(self, inbox, rs):
        out = self.empty(sum(rs), 2)
        if sum(rs):
            self.N.check(self.L.gpuhash_search_segments(C.byref(self.geom), self.table.ptr, self.G,
                                                        self._ptrs(self._segments(inbox, rs, 2)), self.seg_counts.data_ptr(),
                                                        self._ptrs(self._segments(out, rs, 2)), sum(rs), None, 0, None,
                                                        self._stream()), "gpuhash_search_segments")
        return out

    def gather(self, back, cs, perm, n):
        out = self.empty(n, 2)
        if n:
            cnt = self.torch.tensor(list(cs) + [0] * (MAX_SHARDS - self.G), dtype=self.torch.int32, device=self.dev)
            self.N.check(self.L.gpuhash_route_gather(self._ptrs(self._segments(back, cs, 2)), perm.data_ptr(), cnt.data_ptr(),
                                                     self.cap, self.plan.log2, out.data_ptr(), n, None, 0, None,
                                                     self._stream()), "gpuhash_route_gather")
        return out

    def insert_local(self, inbox, rs):
        if sum(rs):
            ptrs = self._segments(inbox, rs, 3)
            self.seg_ptrs_d.copy_(self.torch.tensor(ptrs + [0] * (MAX_SHARDS - self.G), dtype=self.torch.int64))
            self.N.check(self.L.gpuhash_insert_ex(C.byref(self.geom), self.table.ptr, self.seg_ptrs_d.data_ptr(),
                                                  self.seg_counts.data_ptr(), self.G, None, 0, self._stream()), "gpuhash_insert_ex")

    def delete_local(self, inbox, rs):
        if sum(rs):
            self.N.check(self.L.gpuhash_delete_segments(C.byref(self.geom), self.table.ptr, self.G,
                                                        self._ptrs(self._segments(inbox, rs, 3)), self.seg_counts.data_ptr(),
                                                        sum(rs), None, self._stream()), "gpuhash_delete_segments")

    # -- fused path: symmetric arena per rank, exported over CUDA IPC
    #    layout: inbox [G][cap][3] i32 | stage [G][cap][2] i32 | inbox_count [8] | req_flag [8] | res_flag [8] | err [8]
    def _arena_alloc(self):
        from .hashindex import DeviceBuffer
        G, cap = self.G, self.cap
        self.off_inbox, self.off_stage = 0, G * cap * 12
        self.off_cnt = self.off_stage + G * cap * 8
        self.off_reqf, self.off_resf, self.off_err = self.off_cnt + 32, self.off_cnt + 64, self.off_cnt + 96
        self.off_cnt2, self.off_ticket = self.off_cnt + 128, self.off_cnt + 192     # local only: counters [2][8], tickets [2]
        self.arena = DeviceBuffer(self.off_cnt + 256, zero=True)

    def _set_peers(self, peer):
        """peer[r] = address of rank r's arena in THIS process (own, CUDA IPC import, or -- LocalCluster -- plain)"""
        t = self.torch
        self.peer = list(peer)
        G, cap, r = self.G, self.cap, self.rank
        self.pp_peer_inbox = self._ptrs([self.peer[d] + self.off_inbox + r * cap * 12 for d in range(G)])
        self.pp_peer_cnt = self._ptrs([self.peer[d] + self.off_cnt for d in range(G)])
        self.pp_peer_reqf = self._ptrs([self.peer[d] + self.off_reqf for d in range(G)])
        self.pp_peer_resf = self._ptrs([self.peer[d] + self.off_resf for d in range(G)])
        self.pp_my_inbox = self._ptrs([self.arena.ptr + self.off_inbox + s_ * cap * 12 for s_ in range(G)])
        self.pp_origin_stage = self._ptrs([self.peer[s_] + self.off_stage + r * cap * 8 for s_ in range(G)])
        self.pp_my_stage = self._ptrs([self.arena.ptr + self.off_stage + d * cap * 8 for d in range(G)])
        self.seg_ptrs_d.copy_(t.tensor([self.arena.ptr + self.off_inbox + s_ * cap * 12 for s_ in range(G)]
                                       + [0] * (MAX_SHARDS - G), dtype=t.int64))
        t.cuda.synchronize()
        self.p2p = True

    def p2p_setup(self, ix):
        t, L, N = self.torch, self.L, self.N
        G = self.G
        self._arena_alloc()
        handle = (C.c_ubyte * 64)()
        N.check(L.gpuhash_ipc_export(self.arena.ptr, handle), "cudaIpcGetMemHandle")
        mine = t.tensor(list(handle), dtype=t.uint8, device=self.dev)
        allh = [t.empty_like(mine) for _ in range(G)]
        ix.dist.all_gather(allh, mine, group=ix.group)
        peer = []
        for r in range(G):
            if r == self.rank:
                peer.append(self.arena.ptr)
            else:
                raw = (C.c_ubyte * 64)(*allh[r].cpu().tolist())
                p = L.gpuhash_ipc_import(raw)
                if not p:
                    raise N.GpuHashError(f"cudaIpcOpenMemHandle failed for rank {r}")
                peer.append(p)
        ix.dist.barrier(group=ix.group)
        self._set_peers(peer)

    # Waits are separate one-CTA launches (gpuhash_wait_flags): a CTA that waits inside a big kernel keeps an SM slot,
    # and with many lanes in flight two GPUs can fill up with CTAs waiting for each other's producers.
    def _wait(self, off, seq):
        A = self.arena.ptr
        self.N.check(self.L.gpuhash_wait_flags(A + off, self.G, seq, A + self.off_err, self._stream()), "gpuhash_wait_flags")

    def _p2p_scatter(self, ix, req, words, want_perm):
        """wait for the owners' ack of my previous batch; scatter into their inboxes; last CTA publishes counts + flag"""
        L, N, A = self.L, self.N, self.arena.ptr
        ix.seq += 1
        if ix.seq > 1:
            self._wait(self.off_resf, ix.seq - 1)
        if self.route_tiles:                                         # runs sorted in shared memory, written contiguously
            N.check(L.gpuhash_route_scatter_tiles(req.data_ptr() if req.shape[0] else None, req.shape[0], words,
                                                  self.plan.hash_mask_total, self.plan.log2, self.pp_peer_inbox, A + self.off_cnt2,
                                                  self.perm.data_ptr() if want_perm else None, self.cap, self.rank,
                                                  self.pp_peer_cnt, self.pp_peer_reqf, A + self.off_ticket, ix.seq,
                                                  self._stream()), "gpuhash_route_scatter_tiles")
            return
        N.check(L.gpuhash_route_scatter_pub(req.data_ptr() if req.shape[0] else None, req.shape[0], words,
                                            self.plan.hash_mask_total, self.plan.log2, self.pp_peer_inbox, A + self.off_cnt2,
                                            self.perm.data_ptr() if want_perm else None, self.cap, self.rank,
                                            self.pp_peer_cnt, self.pp_peer_reqf, A + self.off_ticket, ix.seq,
                                            None, None, self._stream()), "gpuhash_route_scatter_pub")

    def _p2p_serve(self, ix, op, n_hint=0):
        L, N, A = self.L, self.N, self.arena.ptr
        self._wait(self.off_reqf, ix.seq)
        N.check(L.gpuhash_serve(C.byref(self.geom), self.table.ptr, op, self.plan.log2, self.pp_my_inbox, A + self.off_cnt,
                                self.pp_origin_stage if op == 0 else None, n_hint or self.G * self.cap, None, None,
                                self.rank, self.pp_peer_resf, A + self.off_ticket + 4, ix.seq, None, self._stream()), "gpuhash_serve")

    def p2p_search(self, ix, sel, out=None, after_serve=None):
        """scatter+publish, serve (lookup + result flags), gather -- each behind a flag wait.  after_serve: called once
        this rank's serve kernel is enqueued (the update exchange of the same cycle, running on its own stream, orders
        its serve kernel behind it: the in-stream order search -> insert of the reference, mega_scheduler.c:392-502)"""
        L, N, A = self.L, self.N, self.arena.ptr
        n = sel.shape[0]
        self._p2p_scatter(ix, sel, 2, True)
        self._p2p_serve(ix, 0, 2 * max(n, 1))           # uniform keys: about n requests arrive; the grid strides if more do
        if after_serve:
            after_serve()
        if out is None:
            out = self.empty(n, 2)
        self._p2p_gather(ix, n, out)
        return out

    def _p2p_gather(self, ix, n, out):
        L, N, A = self.L, self.N, self.arena.ptr
        self._wait(self.off_resf, ix.seq)
        if self.route_tiles:
            N.check(L.gpuhash_route_gather_tiles(self.pp_my_stage, self.perm.data_ptr(), self.cap, self.plan.log2,
                                                 out.data_ptr() if n else None, n, self._stream()), "gpuhash_route_gather_tiles")
            return
        N.check(L.gpuhash_route_gather(self.pp_my_stage, self.perm.data_ptr(), A + self.off_cnt2 + 32 * (ix.seq & 1), self.cap,
                                       self.plan.log2, out.data_ptr() if n else None, n, None, 0, None, self._stream()),
                "gpuhash_route_gather")

    def p2p_update(self, ix, iel, insert, before_serve=None):
        """scatter+publish, serve (insert/delete + consumption ack).  The ack is awaited by the NEXT scatter.
        before_serve: called between the two (see p2p_search)."""
        self._p2p_scatter(ix, iel, 3, False)
        if before_serve:
            before_serve()
        self._p2p_serve(ix, 1 if insert else 2, 2 * max(iel.shape[0], 1))

    def p2p_error(self):
        """1 if a flag wait timed out (a peer died); checked by the callers after synchronising"""
        return int(self.arena.download(np.uint32)[self.off_err // 4])


class _Seq:
    """what the fused path needs from its ShardedIndex: the batch sequence number"""
    def __init__(self):
        self.seq = 0


class LocalCluster:
    """G virtual ranks on ONE GPU in one process: every rank has its own shard table and arena, "peer" pointers are
    plain device pointers.  The kernels, flags and buffer layout are exactly those of the multi-process fused path;
    what is missing is NVLink.  Used (a) by the GPU tests, so that a 1-GPU box exercises G = 2, 4, 8, and (b) to
    measure what the routed path costs per GPU without paying for eight of them.

    One call = one exchange of all ranks, issued phase by phase on the current stream (all scatters, all serves,
    all gathers): stream order alone satisfies every flag, so nothing can wait for work that is queued behind it."""

    def __init__(self, plan, cap, algo=0, layout=0, tables=None):
        self.plan, self.G = plan, plan.world
        self.be = [CudaShardBackend(plan, r, cap, algo, layout, table=tables[r] if tables else None) for r in range(self.G)]
        for b in self.be:
            b._arena_alloc()
        for b in self.be:
            b._set_peers([x.arena.ptr for x in self.be])
        self.ix = [_Seq() for _ in range(self.G)]

    @property
    def tables(self):
        return [b.table for b in self.be]

    def search(self, sels, outs=None, after_scatter=None):
        """sels[r]: rank r's requests, int32 [n_r, 2]; returns (or fills) outs[r] int32 [n_r, 2]"""
        be, ix, G = self.be, self.ix, self.G
        for r in range(G):
            be[r]._p2p_scatter(ix[r], sels[r], 2, True)
        if after_scatter:
            after_scatter()
        for r in range(G):
            be[r]._p2p_serve(ix[r], 0, 2 * max(max(s.shape[0] for s in sels), 1))
        if outs is None:
            outs = [be[r].empty(sels[r].shape[0], 2) for r in range(G)]
        for r in range(G):
            be[r]._p2p_gather(ix[r], sels[r].shape[0], outs[r])
        return outs

    def update(self, reqs, insert=True):
        be, ix, G = self.be, self.ix, self.G
        for r in range(G):
            be[r]._p2p_scatter(ix[r], reqs[r], 3, False)
        for r in range(G):
            be[r]._p2p_serve(ix[r], 1 if insert else 2, 2 * max(max(q.shape[0] for q in reqs), 1))

    def insert(self, reqs):
        self.update(reqs, True)

    def delete(self, reqs):
        self.update(reqs, False)

    def error(self):
        return sum(b.p2p_error() for b in self.be)


class ShardExchange:
    """The sharded index as ONE kernel per scheduler cycle (megakv_b200/csrc/gpuhash_xchg.cu): launch j scatters
    exchange j, serves exchange j-1 and gathers exchange j-2, tiles of the three interleaved by ticket; one stream
    memory operation per launch is all the inter-GPU synchronisation.  Mirrors gpuhash_xchg_*.

    step() is COLLECTIVE in the number of calls (every rank calls it equally often; any part may be empty).  The results
    of the step issued as call j are in its `out` tensor once two more steps (or flush()) have run on the stream.
    One exchange = one scheduler cycle of the reference: search -> delete -> insert (mega_scheduler.c:393-504)."""

    def __init__(self, plan, rank, cap_search, cap_update, algo=0, layout=0, table=None):
        import torch
        from . import _native as N
        from .hashindex import DeviceBuffer, make_geom
        self.torch, self.N, self.L = torch, N, N.lib()
        self.plan, self.rank, self.G = plan, rank, plan.world
        self.geom = make_geom(plan.mem_p_total, algo, plan.log2, layout)
        self.table = table if table is not None else DeviceBuffer(self.L.gpuhash_table_bytes(C.byref(self.geom)), zero=True)
        self.x = self.L.gpuhash_xchg_create(C.byref(self.geom), self.table.ptr, plan.hash_mask_total, plan.log2, rank,
                                            int(cap_search), int(cap_update))
        if not self.x:
            raise N.GpuHashError("gpuhash_xchg_create failed (arguments or device memory)")
        nbytes = C.c_size_t()
        self.arena = self.L.gpuhash_xchg_arena(self.x, C.byref(nbytes))
        self.arena_bytes = nbytes.value
        self._imports = []
        self._keep = []                                         # tensors of the exchanges still in flight

    def set_peers(self, arenas):
        arr = (C.c_void_p * MAX_SHARDS)()
        for k, v in enumerate(arenas):
            arr[k] = v
        self.N.check(self.L.gpuhash_xchg_set_peers(self.x, arr), "gpuhash_xchg_set_peers")

    def connect(self, dist, group=None):
        """exchange CUDA IPC handles of the arenas over the process group (one process per GPU)"""
        t, L, N = self.torch, self.L, self.N
        if self.G == 1:
            return
        handle = (C.c_ubyte * 64)()
        N.check(L.gpuhash_ipc_export(self.arena, handle), "cudaIpcGetMemHandle")
        dev = t.device("cuda", t.cuda.current_device())
        mine = t.tensor(list(handle), dtype=t.uint8, device=dev)
        allh = [t.empty_like(mine) for _ in range(self.G)]
        dist.all_gather(allh, mine, group=group)
        peers = []
        for r in range(self.G):
            if r == self.rank:
                peers.append(self.arena)
                continue
            p = L.gpuhash_ipc_import((C.c_ubyte * 64)(*allh[r].cpu().tolist()))
            if not p:
                raise N.GpuHashError(f"cudaIpcOpenMemHandle failed for rank {r}")
            self._imports.append(p)
            peers.append(p)
        dist.barrier(group=group)
        self.set_peers(peers)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def step(self, search=None, out=None, delete=None, insert=None):
        """search: int32 [n, 2] (sig, hash), out: int32 [n, 2]; delete / insert: int32 [n, 3] (sig, hash, loc); device or
        pinned host tensors.  Returns `out` (allocated if missing) -- filled two steps later."""
        ns = 0 if search is None else search.shape[0]
        if ns and out is None:
            out = self.torch.empty((ns, 2), dtype=self.torch.int32, device=search.device)
        nd = 0 if delete is None else delete.shape[0]
        ni = 0 if insert is None else insert.shape[0]
        for x_ in (search, out, delete, insert):
            assert x_ is None or (x_.is_contiguous() and x_.dtype == self.torch.int32)
        self.N.check(self.L.gpuhash_xchg_step(self.x, search.data_ptr() if ns else None, ns, out.data_ptr() if ns else None,
                                              delete.data_ptr() if nd else None, nd, insert.data_ptr() if ni else None, ni,
                                              self._stream()), "gpuhash_xchg_step")
        self._keep = self._keep[-2:] + [(search, out, delete, insert)]
        return out

    def flush(self):
        self.N.check(self.L.gpuhash_xchg_flush(self.x, self._stream()), "gpuhash_xchg_flush")
        self._keep = self._keep[-2:] + [None, None]

    def error(self):
        return self.L.gpuhash_xchg_error(self.x)

    def close(self):
        if self.x:
            self.torch.cuda.synchronize()
            for p in self._imports:
                self.L.gpuhash_ipc_close(p)
            self.L.gpuhash_xchg_destroy(self.x)
            self.x = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalExchangeCluster:
    """G virtual ranks of ShardExchange on ONE GPU in one process (plain device pointers as "peer" arenas): the kernel,
    flags, slots and arena layout of the multi-process run minus NVLink.  One step() = launch j of every rank, issued
    rank by rank on the current stream, which satisfies every flag by stream order."""

    def __init__(self, plan, cap_search, cap_update, algo=0, layout=0, tables=None):
        self.plan, self.G = plan, plan.world
        self.xs = [ShardExchange(plan, r, cap_search, cap_update, algo, layout, table=tables[r] if tables else None) for r in range(self.G)]
        for x in self.xs:
            x.set_peers([y.arena for y in self.xs])

    @property
    def tables(self):
        return [x.table for x in self.xs]

    def step(self, searches=None, outs=None, deletes=None, inserts=None):
        pick = lambda lst, r: None if lst is None else lst[r]
        return [self.xs[r].step(pick(searches, r), pick(outs, r), pick(deletes, r), pick(inserts, r)) for r in range(self.G)]

    def flush(self):
        for _ in range(2):
            self.step()

    def error(self):
        return sum(1 for x in self.xs if x.error() != 0)
