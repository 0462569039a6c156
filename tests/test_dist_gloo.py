"""CPU, world_size 2, gloo: the frame sharding and the {points, elapsed} aggregation that bench.py uses at N > 1."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cloudini_b200 import dist as cdist


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = cdist.shard_frames(11, rank, world)
    pts, ms = cdist.aggregate(len(mine) * 1000, [10.0 + rank, 5.0 - rank])
    sizes = cdist.gather_sizes([100 * rank + f for f in mine])
    q.put((rank, mine, pts, ms, sizes))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_aggregate_world2():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = sorted(f for _, mine, *_ in res for f in mine)
    assert frames == list(range(11))                       # disjoint cover
    assert abs(len(res[0][1]) - len(res[1][1])) <= 1       # balanced
    for rank, mine, pts, ms, sizes in res:
        assert pts == 11 * 1000                            # SUM over ranks
        assert ms == [11.0, 5.0]                           # MAX over ranks, per entry
        assert sizes == [[f for f in range(0, 11, 2)], [100 + f for f in range(1, 11, 2)]]


def test_single_process_passthrough():
    assert cdist.shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]
    assert cdist.aggregate(7, [1.5]) == (7, [1.5])
    assert cdist.gather_sizes([3, 4]) == [[3, 4]]
