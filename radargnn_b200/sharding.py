"""Data-parallel sharding of a batch of frames over the GPUs of one box.

Frames are independent units -- edges never cross frames (the reference builds every frame's graph
separately, preprocessor/radarscenes/dataset_creation.py:651-660, and batches by disjoint union,
utils/data_handling.py:30) -- so each rank runs graph build + conv stack on its own contiguous block
of frames with rank-local node ids and no halo exchange.  The only collective of the path is the
all-reduce of the loss partials (sum, count).  Train-mode BatchNorm statistics are therefore
per-rank (DDP-without-SyncBN semantics)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def shard_frames(frame_sizes: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of frames per rank, balanced by point count.  Returns ``world_size``
    half-open ``(first_frame, last_frame)`` ranges covering all frames in order; a rank may be
    empty when there are fewer frames than ranks."""
    sizes = np.asarray(list(frame_sizes), dtype=np.int64)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if np.any(sizes < 0):
        raise ValueError("frame sizes must be non-negative")
    n_frames = len(sizes)
    prefix = np.concatenate([[0], np.cumsum(sizes)])
    total = int(prefix[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        # first boundary whose prefix is closest to the ideal split, never moving backwards
        b = int(np.searchsorted(prefix, target, side="left"))
        if b > 0 and b <= n_frames and abs(prefix[b - 1] - target) <= abs(prefix[min(b, n_frames)] - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), n_frames))
    bounds.append(n_frames)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def local_frame_ptr(frame_ptr: Sequence[int], first: int, last: int) -> np.ndarray:
    """Rank-local ``frame_ptr`` (starting at 0) of the frames ``[first, last)``."""
    fp = np.asarray(frame_ptr, dtype=np.int64)
    return (fp[first:last + 1] - fp[first]).copy()


def all_reduce_loss(loss_sum: torch.Tensor, count: torch.Tensor, group=None) -> torch.Tensor:
    """Global mean loss from the per-rank partials: one all-reduce of two scalars (NCCL over NVLink
    on the GPU box, gloo in the CPU tests).  ``loss_sum`` / ``count`` are 0-d or 1-element tensors."""
    import torch.distributed as dist
    packed = torch.stack([loss_sum.reshape(()).to(torch.float64), count.reshape(()).to(torch.float64)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[0] / packed[1].clamp(min=1.0)
