"""Frame-sharded multi-GPU plumbing (BASELINE config C5). The codec path has no exchange step: frames are independent
(`PointcloudEncoder` keeps no cross-call state), so ranks only share a start barrier and a tiny reduction of
{points, elapsed} for the aggregate throughput. One process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Frame f goes to rank f mod world: disjoint, covering, balanced to within one frame."""
    return list(range(rank, n_frames, world))


def aggregate(points: int, elapsed_ms: Sequence[float], device=None):
    """All ranks: total points = SUM over ranks, elapsed = MAX over ranks (per entry). Returns (total_points, [ms...])."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(points), [float(x) for x in elapsed_ms]
    t = torch.tensor([float(x) for x in elapsed_ms], dtype=torch.float64, device=device)
    p = torch.tensor([float(points)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(p, op=dist.ReduceOp.SUM)
    return int(p.item()), [float(x) for x in t.tolist()]


def gather_sizes(sizes: Sequence[int], device=None) -> List[List[int]]:
    """Every rank learns the encoded size of every frame (what a consumer needs before pulling blobs to one GPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(sizes)]
    world = dist.get_world_size()
    n = torch.tensor([len(sizes)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    m = int(max(c.item() for c in counts))
    mine = torch.zeros(m, dtype=torch.int64, device=device)
    mine[:len(sizes)] = torch.tensor(list(sizes), dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    return [v[:int(c.item())].tolist() for v, c in zip(allv, counts)]
