"""Data-parallel sharding of independent clips over ranks -- the only parallelism on the hot path.

The reference shards its batch_size-1 DataLoader round-robin with `accelerate` and maps a rank's k-th sample back to the
global clip index with get_original_index(machine_id, local_index, interval=num_devices) = machine_id + local_index*interval
(/root/reference/scripts/neuroclips_video_enhance.py:39-40, used at :324).  There is no collective on the denoising path;
north_star adds one gather of the decoded frames at the end (the reference gathers nothing, each rank writes its own GIFs).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def original_index(rank: int, local_index: int, world_size: int) -> int:
    """Global clip index of a rank's local sample (reference get_original_index)."""
    return rank + local_index * world_size


def shard_indices(num_clips: int, rank: int, world_size: int) -> List[int]:
    """Clips this rank processes, in order: rank, rank + world, rank + 2*world, ..."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_clips, world_size))


def gather_clips(local: torch.Tensor, num_clips: int, group=None) -> Optional[torch.Tensor]:
    """The single end-of-run collective: gather every rank's [n_local, ...] results to rank 0 and restore the global clip
    order.  Uneven shards (num_clips % world != 0) are padded to the longest shard for the all_gather.  Returns the
    [num_clips, ...] tensor on rank 0, None elsewhere.  Works with NCCL (GPU tensors) and gloo (CPU tensors)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[0] == num_clips
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per_rank = (num_clips + world - 1) // world
    n_local = len(shard_indices(num_clips, rank, world))
    assert local.shape[0] == n_local, (local.shape, n_local)
    padded = local.new_zeros((per_rank,) + tuple(local.shape[1:]))
    padded[:n_local] = local
    if padded.is_cuda:         # NCCL: one contiguous receive buffer, no per-rank staging copies
        flat = padded.new_empty((world,) + tuple(padded.shape))
        dist.all_gather_into_tensor(flat, padded.contiguous(), group=group)
        bucket = list(flat.unbind(0))
    else:
        bucket = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(bucket, padded.contiguous(), group=group)
    if rank != 0:
        return None
    out = local.new_empty((num_clips,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = shard_indices(num_clips, r, world)
        out[idx] = bucket[r][: len(idx)]
    return out
