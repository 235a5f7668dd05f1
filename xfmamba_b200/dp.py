"""Data-parallel plumbing for the scan path (one process per GPU, torch.distributed).

The path shards over the batch with no data-path exchange: every rank runs the fused kernels on its own
images.  The only collective is the reduction of PARAMETER gradients in training steps -- for the SS2D core those are
``dA``, ``dDs`` and ``ddelta_bias`` (the activation gradients ``dx, ddelta, dBs, dCs`` stay local).  They are packed into
one flat bucket so that a step costs a single all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> range:
    """contiguous, balanced split of ``global_batch`` images; the first ``global_batch % world`` ranks get one more"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class FlatBucket:
    """views of several tensors inside one contiguous buffer: fill, all-reduce once, read back"""

    def __init__(self, like: Sequence[torch.Tensor]):
        self.shapes = [t.shape for t in like]
        self.sizes = [t.numel() for t in like]
        self.flat = torch.empty(sum(self.sizes), dtype=like[0].dtype, device=like[0].device)
        self.views, off = [], 0
        for shp, n in zip(self.shapes, self.sizes):
            self.views.append(self.flat[off:off + n].view(shp))
            off += n

    def pack(self, tensors: Sequence[torch.Tensor]) -> None:
        for v, t in zip(self.views, tensors):
            v.copy_(t)

    def allreduce(self, average: bool = False, group=None) -> Sequence[torch.Tensor]:
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))
        return self.views


def allreduce_param_grads(tensors: Sequence[torch.Tensor], average: bool = False, group=None) -> None:
    """in-place sum (or mean) of the given gradient tensors over all ranks, as ONE collective"""
    bucket = FlatBucket(tensors)
    bucket.pack(tensors)
    for t, v in zip(tensors, bucket.allreduce(average, group)):
        t.copy_(v)
