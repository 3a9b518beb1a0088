"""CPU suite, world_size 2 over gloo: the N > 1 path of the sharded index (megakv_b200/sharded.py).

The routing logic is the product's; the per-shard table operations are the oracle (tests/sharded_cpu_backend.py).
Checked against the single-table oracle of the LOGICAL table: every search result as a set {o0, o1}, delete
counts, and the concatenation of the shard tables as a multiset of (sig, loc)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megakv_b200.sharded import ShardPlan, ShardedIndex
from oracle import pyoracle as po
from tests import helpers as H
from tests.sharded_cpu_backend import CpuShardBackend


def test_plan_owner_is_closed_under_alternate_bucket(rng):
    for world in (2, 4, 8):
        plan = ShardPlan(24, world)
        o = po.Oracle(24)
        h = rng.integers(0, 2**32, 20000, dtype=np.uint64).astype(np.uint32)
        s = rng.integers(1, 2**32, 20000, dtype=np.uint64).astype(np.uint32)
        own1 = plan.owner(h)
        own2 = plan.owner(o.bucket2(h, s))                      # the alternate bucket as a "hash": same owner
        assert np.array_equal(own1, own2)
        assert own1.min() == 0 and own1.max() == world - 1
        # local bucket index = global index minus the owner's base
        local_mask = (1 << (24 - 6 - plan.log2)) - 1
        assert np.array_equal(o.bucket1(h) & local_mask, o.bucket1(h) - (own1 << plan.shift))
    with pytest.raises(ValueError):
        ShardPlan(24, 16)
    with pytest.raises(ValueError):
        ShardPlan(24, 3)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, mem_p, algo, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(mem_p, world)
        n = 6000
        ix = ShardedIndex(CpuShardBackend(plan, rank, cap=2 * n, algo=algo), plan, exchange="collective")
        rng = np.random.default_rng(777)
        allk = H.random_requests(rng, world * n)                 # every rank derives the same global key set
        mine = allk[rank * n:(rank + 1) * n]
        ref = po.Oracle(mem_p, algo)                             # the logical table, built sequentially

        def t3(a):
            return torch.from_numpy(np.ascontiguousarray(a).view(np.uint32).reshape(-1, 3).view(np.int32).copy())

        def t2(a):
            return torch.from_numpy(np.ascontiguousarray(H.to_sel(a)).view(np.uint32).reshape(-1, 2).view(np.int32).copy())

        ix.insert(t3(mine))
        ref.insert(allk)
        # every rank searches a mix of all ranks' keys plus absent keys, ragged sizes per rank
        probe = np.concatenate([allk[rank::3], H.random_requests(rng, 500 + 100 * rank)])
        got = ix.search(t2(probe)).numpy().view(np.uint32)
        want = ref.search(H.to_sel(probe)).reshape(-1, 2)
        assert np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1)), "search mismatch vs single-table oracle"
        assert (got[: len(allk[rank::3])] == allk[rank::3]["loc"][:, None]).any(axis=1).all()
        # empty batch on one rank
        e = ix.search(t2(allk[:0] if rank == 0 else allk[:10]))
        assert e.shape[0] == (0 if rank == 0 else 10)
        # delete: each rank deletes a slice of ANOTHER rank's keys; total count equals the oracle's
        victim = allk[((rank + 1) % world) * n:((rank + 1) % world) * n + 1000]
        z = torch.tensor([ix.delete(t3(victim))], dtype=torch.int64)
        dist.all_reduce(z)
        zref = sum(ref.delete(allk[((r + 1) % world) * n:((r + 1) % world) * n + 1000]) for r in range(world))
        assert int(z) == zref == 1000 * world
        got = ix.search(t2(victim)).numpy()
        assert not got.any()
        # the shards, concatenated in rank order, hold the logical table's multiset of pairs
        shard = torch.from_numpy(ix.be.o.table.view(np.int32).copy())
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard)
        whole = np.concatenate([p.numpy().view(np.uint32) for p in parts])
        assert ref.digest(table=whole) == ref.digest()
        assert ref.digest(per_bucket=True, table=whole) == ref.digest(per_bucket=True)   # bucket g*B+b is shard g's bucket b
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("algo", [po.CUCKOO, po.TWO_CHOICE])
def test_sharded_index_matches_single_table_oracle(world, algo):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 22, algo, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, "a rank failed"
    assert sorted(ret.keys()) == list(range(world))
