"""CPU tests of the multi-GPU host logic: pair sharding and the single all-gather, with a world_size-2
gloo process group (one process per rank, like the NCCL launch)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussiananything_b200 import sharding


@pytest.mark.parametrize("S,V,world", [(8, 8, 8), (8, 6, 4), (3, 5, 2), (1, 6, 4), (2, 1, 8)])
def test_shard_pairs_is_a_balanced_partition(S, V, world):
    allp = []
    sizes = []
    for r in range(world):
        p = sharding.shard_pairs(S, V, world, r)
        sizes.append(len(p))
        allp += p
    assert sorted(allp) == [(b, v) for b in range(S) for v in range(V)]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_pairs(S, V, world, world)
    g = sharding.group_pairs_by_sample(sharding.shard_pairs(S, V, world, 0))
    assert sum(len(v) for v in g.values()) == sizes[0]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = torch.full((1, 5, 13), float(rank + 1))
        local[0, :, 0] = torch.arange(5.0)
        allg = sharding.all_gather_surfels(local)
        ok = allg.shape == (world, 5, 13) and all(float(allg[r, 0, 1]) == r + 1 for r in range(world))

        class FakeRenderer:            # records which (sample, camera) it was asked to pair
            def render(self, g, cv, cvp, cp, tanfov, **kw):
                assert g.shape[0] == cv.shape[0]
                return {"image": g[:, 0:1, 1:2] * 100.0 + cv[:, :, 0, 0:1]}               # [B', V', 1]
        S, V = world, 3
        cams = torch.zeros(S, V, 4, 4)
        for b in range(S):
            for v in range(V):
                cams[b, v, 0, 0] = 10 * b + v
        res = sharding.render_sharded(FakeRenderer(), local, cams, cams, torch.zeros(S, V, 3), 0.36)
        mine = sharding.shard_pairs(S, V, world, rank)
        ok = ok and sorted(res.keys()) == sorted(mine)
        # sample b came from rank b (marker b + 1) and was rendered with ITS camera (b, v)
        ok = ok and all(float(res[(b, v)]["image"][0]) == (b + 1) * 100.0 + 10 * b + v for (b, v) in mine)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_all_gather_and_render_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]
