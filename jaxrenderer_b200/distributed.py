"""Multi-GPU plumbing: one process per GPU, the batch axis is split
contiguously across ranks, images stay on their device, forward needs no
collective.  In differentiable mode the gradients of SHARED scene parameters
(light, camera, diffuse atlas, ...) are summed across ranks with one NCCL
all-reduce over NVLink; per-image parameters need no exchange.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ``[start, stop)`` of the global batch owned by ``rank``
    (remainder spread over the first ranks)."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's slice of a tensor whose leading axis is the global batch."""
    a, b = shard_range(t.shape[0], rank, world)
    return t[a:b]


def all_reduce_shared_grads(grads: Iterable[torch.Tensor], group=None) -> List[torch.Tensor]:
    """Sum the gradients of shared (un-batched) scene parameters over all ranks.

    The tensors are flattened into ONE bucket (light = 15 floats, camera <= 48,
    atlas up to a few MB: latency-bound, so a single collective) and all-reduced
    in place (NCCL on CUDA tensors, gloo on CPU tensors).  With a fixed world
    size and algorithm the result is reproducible run to run."""
    grads = [g for g in grads if g is not None]
    if not grads or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return grads
