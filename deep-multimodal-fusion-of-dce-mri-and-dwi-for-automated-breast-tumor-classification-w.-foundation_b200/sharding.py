"""Case sharding across GPUs.

Every case is independent at inference (BatchNorm uses running statistics; nothing in the path
mixes cases), so a batch is cut into contiguous per-rank shards, weights are replicated and the
only exchange is one all_gather of the [B/n, K] fp32 logits.  The reference has no distributed
code; this is the B200 arrangement of its single-device loop.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_cases: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank `rank`; the first n % world ranks take one extra case."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_cases, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_logits(local_logits: torch.Tensor, n_cases: int):
    """All ranks call this with their shard's logits; returns the [n_cases, K] logits in case order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    sizes = [shard_bounds(n_cases, r, world)[1] - shard_bounds(n_cases, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local_logits.new_zeros((pad, local_logits.shape[1]))
    buf[: local_logits.shape[0]] = local_logits
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def max_over_ranks(value: float, device) -> float:
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
