"""CPU stand-in for megakv_b200.sharded.CudaShardBackend -- TEST INFRASTRUCTURE.

Same interface, numpy + the oracle instead of kernels, CPU torch tensors instead of CUDA ones, so that the
routing choreography of ShardedIndex (owner computation, split sizes, the two all-to-alls, the un-permute) runs
under gloo with world_size 2 on a machine without GPUs.  The product never imports this file."""
import ctypes as C

import numpy as np
import torch

from oracle import pyoracle as po


class CpuShardBackend:
    def __init__(self, plan, rank, cap, algo=po.CUCKOO):
        self.plan, self.rank, self.cap, self.G = plan, rank, cap, plan.world
        self.o = po.Oracle(plan.mem_p_shard, algo)
        # shard-local geometry: local bucket count, BLOCK_HASH_MASK of the LOGICAL table (gpuhash_geom_init_shard)
        self.o.g.block_mask = (1 << (plan.mem_p_total - 6 - 3)) - 1
        self.p2p = None

    def empty(self, n, words):
        return torch.empty((max(n, 0), words), dtype=torch.int32)

    def scatter(self, req, words, want_perm):
        a = req.numpy().view(np.uint32)
        owner = self.plan.owner(a[:, 1])
        send = torch.zeros((self.G, self.cap, words), dtype=torch.int32)
        perm = torch.zeros((self.G, self.cap), dtype=torch.int32)
        counts = torch.zeros(8, dtype=torch.int32)
        for d in range(self.G):
            idx = np.nonzero(owner == d)[0]
            send[d, : len(idx)] = req[idx]
            perm[d, : len(idx)] = torch.from_numpy(idx.astype(np.int32))
            counts[d] = len(idx)
        return send, counts, perm

    def pack(self, send, cs):
        return torch.cat([send[d, : cs[d]] for d in range(self.G)], dim=0).contiguous()

    def search_local(self, inbox, rs):
        sel = np.ascontiguousarray(inbox.numpy()).view(np.uint32).reshape(-1, 2)
        q = np.empty(len(sel), dtype=po.SEL_DT); q["sig"], q["hash"] = sel[:, 0], sel[:, 1]
        out = self.o.search(q).view(np.int32).reshape(-1, 2)
        return torch.from_numpy(out.copy())

    def gather(self, back, cs, perm, n):
        out = torch.zeros((n, 2), dtype=torch.int32)
        off = 0
        for d in range(self.G):
            out[perm[d, : cs[d]].long()] = back[off: off + cs[d]]
            off += cs[d]
        return out

    def _iel(self, inbox):
        a = np.ascontiguousarray(inbox.numpy()).view(np.uint32).reshape(-1, 3)
        e = np.empty(len(a), dtype=po.IEL_DT); e["sig"], e["hash"], e["loc"] = a[:, 0], a[:, 1], a[:, 2]
        return e

    def insert_local(self, inbox, rs):
        self.o.insert(self._iel(inbox))

    def delete_local(self, inbox, rs):
        return self.o.delete(self._iel(inbox))
