"""Frame sharding across GPUs (one process per GPU) and the only collective of the path: the
final gather of the per-frame quantities table.  Frames are independent
(/root/reference/src/predict.py:85-100 has no cross-frame state), so there is no data-path
collective; masks stay on the rank that produced them."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the SORTED frame list owned by ``rank``:
    lo = ceil(n*rank/world), hi = ceil(n*(rank+1)/world)."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    lo = -(-n_frames * rank // world)
    hi = -(-n_frames * (rank + 1) // world)
    return lo, hi


def gather_table(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather a per-frame table [n_local, K] (int32) into [n_total, K] in frame order.
    Slices are padded to the longest slice so one all_gather_into_tensor / all_gather suffices."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == n_total
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros(longest, local.shape[1], dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
