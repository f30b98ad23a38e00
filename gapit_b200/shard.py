"""Marker sharding across ranks.

Markers split into contiguous column blocks, rank r owns [lo, hi); the output
order is the concatenation, so the marker order of the reference is preserved.
Integer partial cross-products are summed across ranks (exact for any rank
count); the scan needs no collective, rank 0 gathers the per-shard results.
"""
from __future__ import annotations

import numpy as np


def marker_shard(m: int, rank: int, world: int):
    """Contiguous, balanced block of markers owned by `rank`."""
    lo = (m * rank) // world
    hi = (m * (rank + 1)) // world
    return lo, hi


def concat_shards(parts):
    """Rank-ordered concatenation of per-shard result arrays (last axis = markers)."""
    return np.concatenate(list(parts), axis=-1)


def vanraden_from_sums(G, sums, n):
    """Host restatement of the kinship epilogue from all-reduced integers (used to check the
    distributed plumbing on CPU): G int64 (n x n, all markers), sums = [sum c, sum c^2, #c==2n]."""
    sum_c, sum_c2, n2 = (int(v) for v in sums)
    c2 = sum_c2 - n2 * 4 * n * n
    c1 = sum_c - n2 * 2 * n
    D = 2 * n * c1 - c2
    Gp = G.astype(object) - 4 * n2
    s = Gp.sum(axis=1)
    num = n * n * Gp - n * (s[:, None] + s[None, :]) + c2
    return (2 * num / D).astype(np.float64)


def gather_markers(local, m, rank, world, dst=0, device=None):
    """Gather per-shard marker arrays (1-D, float64) to `dst` in marker order.

    Shards may differ by one marker, so each is padded to the largest shard before the
    collective and trimmed after it.  Returns the length-m array on `dst`, None elsewhere.
    Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    sizes = [marker_shard(m, r, world)[1] - marker_shard(m, r, world)[0] for r in range(world)]
    mx = max(sizes)
    buf = torch.zeros(mx, dtype=torch.float64, device=device)
    t = torch.as_tensor(local, dtype=torch.float64)
    buf[: t.shape[0]] = t.to(buf.device)
    if rank == dst:
        outs = [torch.empty(mx, dtype=torch.float64, device=device) for _ in range(world)]
        dist.gather(buf, outs, dst=dst)
        return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)])
    dist.gather(buf, None, dst=dst)
    return None
