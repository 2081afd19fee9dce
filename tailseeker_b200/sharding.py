"""Multi-GPU plumbing: one process per GPU, tiles / lanes sharded across ranks.

Tiles are independent on the whole per-cluster path, so there is no data-path
collective.  The only exchange is the sum the reference forms in Python when it merges per-tile
files: the pos/neg score histograms before cutoff estimation
(scripts/calculate-optimal-parameters.py:139-140), the ruler's positive histogram per round
(scripts/measure-polya-lengths.py:64-70) and the demultiplexing counters
(scripts/stats-merge-demultiplexing-counts.py:33-53).  Integer sums: order-independent, exact.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_tiles(tile_ids: Sequence, rank: int, world: int) -> List:
    """Round-robin tile -> rank assignment (lane-per-GPU when tiles are listed lane-major and
    world divides the lane count)."""
    return [t for i, t in enumerate(tile_ids) if i % world == rank]


def allreduce_stats(tensors: Sequence[torch.Tensor], group=None) -> None:
    """In-place SUM all-reduce of integer statistics tensors (int64) across ranks.
    NCCL over NVLink on GPUs, gloo in the CPU tests.  No-op without a process group."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
