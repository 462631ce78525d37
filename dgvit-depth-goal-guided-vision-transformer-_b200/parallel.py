"""Host-side data-parallel plumbing: one process per GPU, every rank holds a full
replica and updates on its own shard of the global minibatch; the two flat gradient arenas are
SUM-all-reduced (NCCL on GPUs, gloo in the CPU tests) between the phases of the update.  Losses
are means over the GLOBAL batch, so each rank's kernels scale by 1/global_batch and the reduced
gradients need no further division."""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_bounds(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [offset, offset+count) of rank ``rank``; remainders go to the low ranks."""
    base, rem = divmod(global_batch, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def shard_batch(batch: Dict[str, torch.Tensor], world: int, rank: int) -> Dict[str, torch.Tensor]:
    """Slice every [B, ...] tensor of a minibatch / noise dict to this rank's shard."""
    out = {}
    for k, v in batch.items():
        if v is None or not torch.is_tensor(v):
            out[k] = v
            continue
        o, c = shard_bounds(v.shape[0], world, rank)
        out[k] = v[o:o + c].contiguous()
    return out


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce of a flat arena (no-op when not distributed)."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    return t
