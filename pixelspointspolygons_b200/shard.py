"""Batch-wise sharding of a jagged LiDAR batch over the GPUs of one box.

Tiles are independent units: rank g of G encodes tiles [g*B/G, (g+1)*B/G) (remainder spread over the first ranks) and
there is no collective on the inference path.  This is what the reference gets from `DistributedSampler`
(R:pixelspointspolygons/datasets/build_datasets.py:145,195) one level up, restated for a batch that already exists
(serving / bench): the jagged offsets are rebased per shard, the values are sliced without a copy.
`gather_tiles` is the optional result exchange (an all_gather of the per-rank outputs) used by tests and by callers
that want the full batch on every rank; the hot path never calls it.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_bounds(num_tiles: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the tiles owned by `rank`: contiguous, sizes differ by at most one, earlier ranks get the extras."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(num_tiles), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_jagged(values: torch.Tensor, offsets: torch.Tensor, rank: int, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(values (sumN, c), offsets (B+1)) -> this rank's (values view, rebased offsets (b+1))."""
    B = offsets.numel() - 1
    lo, hi = shard_bounds(B, rank, world)
    offs = offsets[lo:hi + 1]
    p0, p1 = (int(offs[0]), int(offs[-1])) if hi > lo else (0, 0)
    return values[p0:p1], (offs - offs[0]) if hi > lo else offsets.new_zeros(1)


def shard_lidar(x_lidar, rank: int, world: int):
    """Jagged NestedTensor, dense (B, N, 3) tensor or list of (N_b, 3) tensors -> the same kind, this rank's tiles."""
    if isinstance(x_lidar, (list, tuple)):
        lo, hi = shard_bounds(len(x_lidar), rank, world)
        return list(x_lidar[lo:hi])
    if x_lidar.is_nested:
        v, o = shard_jagged(x_lidar.values(), x_lidar.offsets(), rank, world)
        return torch.nested.nested_tensor_from_jagged(v, o.to(x_lidar.offsets().dtype))
    lo, hi = shard_bounds(x_lidar.shape[0], rank, world)
    return x_lidar[lo:hi]


def shard_dense(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_tiles(local_out: torch.Tensor, num_tiles: int, group=None) -> torch.Tensor:
    """Concatenate the per-rank outputs (tile-major) in tile order on every rank.  Not part of the hot path."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_bounds(num_tiles, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = local_out.new_zeros((cap,) + tuple(local_out.shape[1:]))
    pad[: local_out.shape[0]] = local_out
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)
