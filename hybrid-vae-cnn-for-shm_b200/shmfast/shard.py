"""Window-range sharding across the GPUs of one box.

Windows are independent, so the scoring path has NO data-path collective: rank r owns the contiguous
window range [lo_r, hi_r); weights and normalisation stats are replicated; when gathering from a series
each rank reads its start-index range plus a (T-1)-row halo.  Results are laid out by global window
index, so concatenating the per-rank slices in rank order reproduces the single-GPU outputs
(compacted indices are offset by the slice base, which preserves np.where's ascending order).
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to collect results and timings.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of [0, n): the first n % world ranks get one extra window."""
    if world <= 0 or not 0 <= rank < world or n < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def series_rows_for(lo: int, hi: int, T: int, stride: int) -> Tuple[int, int]:
    """Row range of the series that windows [lo, hi) touch (halo of T-1 rows at the end)."""
    if hi <= lo:
        return lo * stride, lo * stride
    return lo * stride, (hi - 1) * stride + T


def gather_by_rank(local: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate per-rank result slices (possibly of different lengths) in rank order."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def gather_flagged(idx_local: torch.Tensor, count_local: int, lo: int, group=None) -> torch.Tensor:
    """Global ascending flagged-window list from per-rank compacted lists (indices local to [lo, hi))."""
    return gather_by_rank(idx_local[:count_local].to(torch.int64) + lo, group)


def max_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    if not dist.is_available() or not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    if not dist.is_available() or not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
