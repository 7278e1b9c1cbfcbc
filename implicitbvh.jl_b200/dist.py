"""Multi-GPU plumbing for the two paths that shard: one process per GPU,
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) for the collectives.

  * build: does not shard (global sort + reduction). It runs on rank 0 and the built tree (leaves,
    nodes) is broadcast — or, with `replicate=True`, every rank builds the same deterministic tree.
  * traverse(bvh) / traverse(bvh1, bvh2): contiguous ranges of query leaves per rank; concatenating
    the shards in rank order reproduces the single-GPU contact list exactly.
  * traverse_rays: contiguous ray ranges per rank, same concatenation property.

The gather of variable-length shards is an all_gather of the per-rank counts followed by one padded
all_gather_into_tensor of the contact bytes. Only host logic lives here; it is covered by the
world_size-2 gloo tests on CPU (tests/test_dist_gloo.py) with fake shards.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, weights: Optional[np.ndarray] = None) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) query ranges, one per rank. With `weights` (per-query cost, e.g. the
    per-leaf contact counts of the previous step) the ranges carry near-equal total weight."""
    if world <= 0:
        raise ValueError("world must be positive")
    if weights is None:
        q, r = divmod(n, world)
        out, b = [], 0
        for k in range(world):
            e = b + q + (1 if k < r else 0)
            out.append((b, e))
            b = e
        return out
    w = np.asarray(weights, np.float64)
    assert len(w) == n
    c = np.concatenate([[0.0], np.cumsum(w + 1e-9)])
    cuts = [0]
    for k in range(1, world):
        cuts.append(int(np.searchsorted(c, c[-1] * k / world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(cuts[k]), int(cuts[k + 1])) for k in range(world)]


def broadcast_tensor_(t: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    dist.broadcast(t, src=src, group=group)
    return t


def gather_shards(shard: torch.Tensor, count: int, itemsize: int, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather variable-length byte shards. `shard` is a uint8 tensor holding at least count*itemsize
    bytes. Returns (concatenated uint8 tensor in rank order, per-rank counts)."""
    world = dist.get_world_size(group)
    dev = shard.device
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts_h = [int(c) for c in counts.cpu().tolist()]
    maxc = max(counts_h) if counts_h else 0
    if maxc == 0:
        return torch.empty(0, dtype=torch.uint8, device=dev), counts_h
    pad = torch.empty(maxc * itemsize, dtype=torch.uint8, device=dev)
    nb = int(count) * itemsize
    if nb:
        pad[:nb] = shard[:nb]
    allb = torch.empty(world * maxc * itemsize, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allb, pad, group=group)
    parts = [allb[r * maxc * itemsize: r * maxc * itemsize + counts_h[r] * itemsize] for r in range(world)]
    return torch.cat(parts), counts_h


def concat_in_rank_order(parts: Sequence[np.ndarray]) -> np.ndarray:
    return np.concatenate(list(parts)) if len(parts) else np.empty(0)
