"""Data-parallel plumbing for the feature front-end: frames are independent, so a batch (or a clip list) is cut
into contiguous per-rank shards and the transform path needs no collective.  The only
exchange on the path is the fingerprint's all-reduce (fingerprint.FingerprintAccumulator.all_reduce).

Mirrors what the reference gets from ``DistributedSampler`` (train_classifier.py:50-229): every rank sees a
disjoint, near-equal slice and the union is the whole set.
"""
from __future__ import annotations

import os

import torch


def dist_info() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n`` items owned by ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_shard(frames: torch.Tensor, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    """This rank's slice of a batch along dim 0 (a view; nothing is copied)."""
    r, _, w = dist_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    lo, hi = shard_bounds(frames.shape[0], rank, world)
    return frames[lo:hi]


def gather_features(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather per-rank feature batches back into one batch (evaluation bookkeeping only; the training loop
    never needs it).  Ragged shards are padded to the largest shard and trimmed."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    padded = local if local.shape[0] == cap else torch.cat(
        [local, local.new_zeros((cap - local.shape[0],) + tuple(local.shape[1:]))])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])
