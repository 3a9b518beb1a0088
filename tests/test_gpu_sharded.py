"""GPU suite, needs >= 2 devices: the sharded index with real kernels, both exchanges, against the single-table
oracle of the logical table.  Skipped on a 1-GPU box (the CPU suite covers the routing logic over gloo)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, exchange, ret):
    import torch
    import torch.distributed as dist
    import megakv_b200 as mk
    from megakv_b200.sharded import ShardPlan, ShardedIndex, CudaShardBackend
    from oracle import pyoracle as po
    from tests import helpers as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank); mk.lib().gpuhash_set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        mem_p, n = 24, 50000
        plan = ShardPlan(mem_p, world)
        be = CudaShardBackend(plan, rank, cap=1 << 17)
        ix = ShardedIndex(be, plan, exchange=exchange)
        rng = np.random.default_rng(4242)
        allk = H.random_requests(rng, world * n)
        mine = allk[rank * n:(rank + 1) * n]
        ref = po.Oracle(mem_p); ref.insert(allk)
        dev = torch.device("cuda", rank)

        def t3(a):
            return torch.from_numpy(np.ascontiguousarray(a).view(np.uint32).reshape(-1, 3).view(np.int32).copy()).to(dev)

        def t2(a):
            return torch.from_numpy(np.ascontiguousarray(H.to_sel(a)).view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(dev)

        for part in np.array_split(mine, 3):
            ix.insert(t3(part))
        for rep in range(3):                                      # several batches back to back: buffer reuse + flags
            probe = np.concatenate([allk[(rank + rep)::5], H.random_requests(rng, 1000 + 17 * rank)])
            got = ix.search(t2(probe)).cpu().numpy().view(np.uint32)
            want = ref.search(H.to_sel(probe)).reshape(-1, 2)
            assert np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1)), f"search mismatch ({exchange}, rep {rep})"
        assert ix.search(t2(allk[:0])).shape[0] == 0
        victim = allk[((rank + 1) % world) * n:((rank + 1) % world) * n + 3000]
        ix.delete(t3(victim))
        torch.cuda.synchronize(); dist.barrier()
        assert not ix.search(t2(victim)).cpu().numpy().any()
        for r in range(world):
            ref.delete(allk[((r + 1) % world) * n:((r + 1) % world) * n + 3000])
        # shard tables, concatenated in rank order, equal the logical table bucket for bucket (as multisets)
        t = mk.DeviceTable.__new__(mk.DeviceTable); t.geom, t.ptr, t.nbytes = be.geom, be.table.ptr, be.table.nbytes
        shard = torch.from_numpy(t.dump_reference().view(np.int32).copy()).to(dev)
        t.ptr = None
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard)
        whole = np.concatenate([p.cpu().numpy().view(np.uint32) for p in parts])
        assert ref.digest(table=whole) == ref.digest()
        if exchange == "p2p":
            assert be.p2p_error() == 0
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("exchange", ["collective", "p2p"])
def test_sharded_index_on_gpus(gpu, exchange, world):
    """world 1: the whole routed path (scatter, publish, wait, serve, gather) with this GPU as its own peer, so a
    1-GPU box still runs every kernel of the sharded index; world 2, 4, 8: one process per GPU across NVLink (skipped on
    boxes with fewer GPUs; `gpurun --gpus N -- python -m pytest tests/test_gpu_sharded.py -m gpu` runs them)."""
    if gpu.lib().gpuhash_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, "a rank failed"
    assert sorted(ret.keys()) == list(range(world))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("layout", [0, 1], ids=["pairs", "reflayout"])
def test_virtual_shards_on_one_gpu(gpu, world, layout):
    """G = 2, 4, 8 virtual ranks on ONE GPU (megakv_b200.sharded.LocalCluster): the kernels, flags, regions and peer
    pointer tables of the fused path exactly as the multi-process run uses them, minus NVLink -- so a 1-GPU box
    checks every shard count against the single-table oracle: search words, deletes, and the shard tables
    concatenated in rank order equal to the logical table as a multiset."""
    import torch
    import megakv_b200 as mk
    from megakv_b200.sharded import ShardPlan, LocalCluster
    from oracle import pyoracle as po
    from tests import helpers as H
    mk.lib().gpuhash_set_device(0); torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    mem_p, n = 24, 30000
    plan = ShardPlan(mem_p, world)
    cl = LocalCluster(plan, cap=1 << 16, layout=layout)
    rng = np.random.default_rng(777 + world)
    allk = H.random_requests(rng, world * n)
    ref = po.Oracle(mem_p); ref.insert(allk)

    def t3(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint32).reshape(-1, 3).view(np.int32).copy()).to(dev)

    def t2(a):
        return torch.from_numpy(np.ascontiguousarray(H.to_sel(a)).view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(dev)

    for part in range(3):                                         # three exchanges: buffer reuse, flags, acks
        cl.insert([t3(np.array_split(allk[r * n:(r + 1) * n], 3)[part]) for r in range(world)])
    for rep in range(3):
        probes = [np.concatenate([allk[(r + rep)::7], H.random_requests(rng, 500 + 13 * r)]) for r in range(world)]
        if rep == 2:
            probes[0] = probes[0][:0]                             # a rank with nothing to ask still takes part
        outs = cl.search([t2(p) for p in probes])
        for r in range(world):
            got = outs[r].cpu().numpy().view(np.uint32)
            want = ref.search(H.to_sel(probes[r])).reshape(-1, 2)
            assert np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1)), f"search mismatch rank {r} rep {rep}"
    victims = [allk[((r + 1) % world) * n:((r + 1) % world) * n + 2000] for r in range(world)]
    cl.delete([t3(v) for v in victims])
    for v in victims:
        ref.delete(v)
    outs = cl.search([t2(v) for v in victims])
    torch.cuda.synchronize()
    assert not any(o.cpu().numpy().any() for o in outs)
    assert cl.error() == 0
    parts = []
    for b in cl.be:
        t = mk.DeviceTable.__new__(mk.DeviceTable); t.geom, t.ptr, t.nbytes = b.geom, b.table.ptr, b.table.nbytes
        parts.append(t.dump_reference()); t.ptr = None
    assert ref.digest(table=np.concatenate(parts)) == ref.digest()


def test_virtual_shards_many_tiles_per_cta(gpu):
    """Exchanges large enough that every persistent CTA of the staged serve kernel walks several 64-request tiles,
    including the partial tile that ends a region in the middle of its walk (its barrier parity is per stage, not per
    iteration), and that scatter/gather stride their grids: 2 ranks x 300 001 searches, three times over."""
    import torch
    import megakv_b200 as mk
    from megakv_b200.sharded import ShardPlan, LocalCluster
    from oracle import pyoracle as po
    from tests import helpers as H
    mk.lib().gpuhash_set_device(0); torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    world, mem_p, n = 2, 26, 300001
    plan = ShardPlan(mem_p, world)
    cl = LocalCluster(plan, cap=1 << 19)
    rng = np.random.default_rng(99)
    allk = H.random_requests(rng, world * n)
    ref = po.Oracle(mem_p); ref.insert(allk)

    def t3(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint32).reshape(-1, 3).view(np.int32).copy()).to(dev)

    def t2(a):
        return torch.from_numpy(np.ascontiguousarray(H.to_sel(a)).view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(dev)

    cl.insert([t3(allk[r * n:(r + 1) * n]) for r in range(world)])
    for rep in range(3):
        probes = [np.concatenate([allk[rng.integers(0, world * n, n - 5000)], H.random_requests(rng, 5000 - 37 * r)]) for r in range(world)]
        outs = cl.search([t2(p) for p in probes])
        for r in range(world):
            got = outs[r].cpu().numpy().view(np.uint32)
            want = ref.search(H.to_sel(probes[r])).reshape(-1, 2)
            assert np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1)), f"search mismatch rank {r} rep {rep}"
    assert cl.error() == 0
