"""GPU suite: the sharded index as one kernel per scheduler cycle (megakv_b200/csrc/gpuhash_xchg.cu) against the
single-table oracle of the logical table.

Semantics under test (what the kernel promises, and what the oracle model below replays): exchange e sees every
effect of the exchanges before it; inside an exchange every rank's searches run before any delete, every delete before
any insert (the reference's per-cycle order search -> delete -> insert, mega_scheduler.c:393-502, across all "workers")."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t3(torch, dev, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint32).reshape(-1, 3).view(np.int32).copy()).to(dev)


def _t2(torch, dev, H, a):
    return torch.from_numpy(np.ascontiguousarray(H.to_sel(a)).view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(dev)


def _scenario(rng, H, world, n):
    """exchanges[e][r] = (search, delete, insert) request arrays (ielem records; searches use sig/hash) of rank r"""
    allk = H.random_requests(rng, world * n)
    own = [allk[r * n:(r + 1) * n] for r in range(world)]
    thirds = [np.array_split(o, 3) for o in own]
    none = allk[:0]
    victims = [own[(r + 1) % world][:2000 + 3 * r] for r in range(world)]           # someone else's first-third keys
    ex = []
    ex.append([(none, none, thirds[r][0]) for r in range(world)])
    ex.append([(np.concatenate([allk[(r + 1)::7], H.random_requests(rng, 500 + 13 * r)]), none, thirds[r][1]) for r in range(world)])
    ex.append([(victims[r], victims[r], thirds[r][2]) for r in range(world)])        # searched BEFORE they are deleted
    e4 = [(np.concatenate([victims[r], allk[r::5]]), none, none) for r in range(world)]
    e4[0] = (none, none, none)                                                        # a rank with nothing to do still takes part
    ex.append(e4)
    ex.append([(allk[(3 * r)::11], none, victims[r][:100]) for r in range(world)])    # re-insert a few
    ex.append([(victims[r][:300], none, none) for r in range(world)])
    return allk, ex


def _replay(ref, H, ex):
    """expected search words per exchange and rank, replaying the semantics above on the single-table oracle"""
    want = []
    for per_rank in ex:
        want.append([ref.search(H.to_sel(s)).reshape(-1, 2) for (s, d, i) in per_rank])
        for (s, d, i) in per_rank:
            if len(d):
                ref.delete(d)
        for (s, d, i) in per_rank:
            if len(i):
                ref.insert(i)
    return want


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("layout", [0, 1], ids=["pairs", "reflayout"])
def test_exchange_virtual_ranks(gpu, world, layout):
    """G = 1, 2, 4, 8 virtual ranks on ONE GPU: same kernel, flags, slots and arena layout as one process per GPU."""
    import torch
    import megakv_b200 as mk
    from megakv_b200.sharded import ShardPlan, LocalExchangeCluster
    from oracle import pyoracle as po
    from tests import helpers as H
    mk.lib().gpuhash_set_device(0); torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    mem_p, n = 24, 30000
    plan = ShardPlan(mem_p, world)
    cl = LocalExchangeCluster(plan, cap_search=1 << 16, cap_update=1 << 15, layout=layout)
    rng = np.random.default_rng(4000 + world)
    allk, ex = _scenario(rng, H, world, n)
    ref = po.Oracle(mem_p)
    want = _replay(ref, H, ex)
    outs = []
    for per_rank in ex:
        outs.append(cl.step([_t2(torch, dev, H, s) for (s, d, i) in per_rank], None,
                            [_t3(torch, dev, d) for (s, d, i) in per_rank], [_t3(torch, dev, i) for (s, d, i) in per_rank]))
    cl.flush()
    torch.cuda.synchronize()
    assert cl.error() == 0
    for e, per_rank in enumerate(ex):
        for r in range(world):
            if len(per_rank[r][0]) == 0:
                continue
            got = outs[e][r].cpu().numpy().view(np.uint32)
            assert np.array_equal(np.sort(got, axis=1), np.sort(want[e][r], axis=1)), f"search mismatch: exchange {e}, rank {r}"
    parts = []
    for x in cl.xs:
        t = mk.DeviceTable.__new__(mk.DeviceTable); t.geom, t.ptr, t.nbytes = x.geom, x.table.ptr, x.table.nbytes
        parts.append(t.dump_reference()); t.ptr = None
    assert ref.digest(table=np.concatenate(parts)) == ref.digest()


def test_exchange_many_tiles_and_ragged_ends(gpu):
    """Exchanges of several hundred thousand requests per rank (every warp walks many interleaved tiles; regions end in the
    middle of a 64-request lookup tile and a 256-request routing tile), three in flight, odd sizes, a misaligned input."""
    import torch
    import megakv_b200 as mk
    from megakv_b200.sharded import ShardPlan, LocalExchangeCluster
    from oracle import pyoracle as po
    from tests import helpers as H
    mk.lib().gpuhash_set_device(0); torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    world, mem_p, n = 2, 26, 300001
    plan = ShardPlan(mem_p, world)
    cl = LocalExchangeCluster(plan, cap_search=1 << 19, cap_update=1 << 19)
    rng = np.random.default_rng(99)
    allk = H.random_requests(rng, world * n)
    ref = po.Oracle(mem_p); ref.insert(allk)
    cl.step(inserts=[_t3(torch, dev, allk[r * n:(r + 1) * n]) for r in range(world)])
    outs, wants = [], []
    for rep in range(4):
        probes = [np.concatenate([allk[rng.integers(0, world * n, n - 5000 - 77 * rep)], H.random_requests(rng, 5000 - 37 * r)]) for r in range(world)]
        sel = []
        for r, p in enumerate(probes):
            t = _t2(torch, dev, H, p)
            if rep == 1 and r == 0:                               # 8 B-aligned but not 16 B-aligned input and output
                pad = torch.empty((t.shape[0] + 1, 2), dtype=torch.int32, device=dev)
                pad[1:] = t
                t = pad[1:]
            sel.append(t)
        o = None
        if rep == 1:
            o = [torch.empty((s.shape[0] + 1, 2), dtype=torch.int32, device=dev)[1:] for s in sel]
        outs.append(cl.step(sel, o))
        wants.append([ref.search(H.to_sel(p)).reshape(-1, 2) for p in probes])
    cl.flush()
    torch.cuda.synchronize()
    assert cl.error() == 0
    for rep in range(4):
        for r in range(world):
            got = outs[rep][r].cpu().numpy().view(np.uint32)
            assert np.array_equal(np.sort(got, axis=1), np.sort(wants[rep][r], axis=1)), f"search mismatch rank {r} rep {rep}"


def test_exchange_search_precedes_insert_of_the_same_exchange(gpu):
    """A key searched and inserted in the SAME exchange is not found by that search (search -> insert order of a cycle),
    and is found by the next exchange; a key deleted and re-inserted in one exchange ends up present."""
    import torch
    import megakv_b200 as mk
    from megakv_b200.sharded import ShardPlan, LocalExchangeCluster
    from tests import helpers as H
    mk.lib().gpuhash_set_device(0); torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    world = 4
    plan = ShardPlan(22, world)
    cl = LocalExchangeCluster(plan, cap_search=1 << 14, cap_update=1 << 14)
    rng = np.random.default_rng(5)
    keys = [H.random_requests(rng, 5000, loc_base=1 + 10000 * r) for r in range(world)]
    o1 = cl.step([_t2(torch, dev, H, k) for k in keys], None, None, [_t3(torch, dev, k) for k in keys])
    o2 = cl.step([_t2(torch, dev, H, k) for k in keys], None, [_t3(torch, dev, k) for k in keys], [_t3(torch, dev, k) for k in keys])
    o3 = cl.step([_t2(torch, dev, H, k) for k in keys])
    cl.flush(); torch.cuda.synchronize()
    assert cl.error() == 0
    for r in range(world):
        assert not o1[r].cpu().numpy().any()
        for o in (o2[r], o3[r]):
            got = o.cpu().numpy().view(np.uint32)
            loc = keys[r]["loc"][:, None]                         # (both words carry it when bucket 1 == bucket 2: sig & BLOCK_HASH_MASK == 0)
            assert ((got == loc).sum(axis=1) >= 1).all() and ((got == loc) | (got == 0)).all()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    import megakv_b200 as mk
    from megakv_b200.sharded import ShardPlan, ShardExchange
    from oracle import pyoracle as po
    from tests import helpers as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank); mk.lib().gpuhash_set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = torch.device("cuda", rank)
        mem_p, n = 24, 30000
        plan = ShardPlan(mem_p, world)
        x = ShardExchange(plan, rank, cap_search=1 << 16, cap_update=1 << 15)
        x.connect(dist)
        rng = np.random.default_rng(4000 + world)                  # every rank builds the whole scenario and the oracle's replay
        allk, ex = _scenario(rng, H, world, n)
        ref = po.Oracle(mem_p)
        want = _replay(ref, H, ex)
        outs = [x.step(_t2(torch, dev, H, pr[rank][0]), None, _t3(torch, dev, pr[rank][1]), _t3(torch, dev, pr[rank][2])) for pr in ex]
        x.flush()
        torch.cuda.synchronize(); dist.barrier()
        assert x.error() == 0
        for e, pr in enumerate(ex):
            if len(pr[rank][0]) == 0:
                continue
            got = outs[e].cpu().numpy().view(np.uint32)
            assert np.array_equal(np.sort(got, axis=1), np.sort(want[e][rank], axis=1)), f"search mismatch: exchange {e}, rank {rank}"
        t = mk.DeviceTable.__new__(mk.DeviceTable); t.geom, t.ptr, t.nbytes = x.geom, x.table.ptr, x.table.nbytes
        shard = torch.from_numpy(t.dump_reference().view(np.int32).copy()).to(dev)
        t.ptr = None
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard)
        whole = np.concatenate([p.cpu().numpy().view(np.uint32) for p in parts])
        assert ref.digest(table=whole) == ref.digest()
        dist.barrier()
        x.close()
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_exchange_one_process_per_gpu(gpu, world):
    """one process per GPU, arenas exchanged over CUDA IPC, peer stores over NVLink (skipped on boxes with fewer GPUs;
    `gpurun --gpus N -- python -m pytest tests/test_gpu_xchg.py -m gpu` runs them)"""
    if gpu.lib().gpuhash_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, "a rank failed"
    assert sorted(ret.keys()) == list(range(world))
