"""Multi-GPU plumbing for the two paths that shard: one process per GPU,
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) for the collectives.

  * build: does not shard (global sort + reduction). It runs on rank 0 and the built tree (leaves,
    nodes) is broadcast — or, with `replicate=True`, every rank builds the same deterministic tree.
  * traverse(bvh) / traverse(bvh1, bvh2): contiguous ranges of query leaves per rank; concatenating
    the shards in rank order reproduces the single-GPU contact list exactly.
  * traverse_rays: contiguous ray ranges per rank, same concatenation property.

Two gathers of the variable-length shards:
  * `PeerGather` (the product path on NVLink boxes): one kernel of libibvh_b200.so per rank
    (ibvh_allgather_pairs) exchanges the counts and writes the shard into every rank's list through
    peer / NVSwitch-multicast memory; torch's symmetric memory only provides the mapped buffers.
  * `gather_shards`: NCCL / gloo collectives (all_gather of the counts, then one padded
    all_gather_into_tensor); the portable path, covered by the world_size-2 gloo tests on CPU
    (tests/test_dist_gloo.py) with fake shards.
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


class PeerGather:
    """All-gather of (index, index) pair shards over NVLink peer memory, rank order preserved.

    Every rank constructs it with the same capacity. `gather(shard, count)` is a collective: it returns
    (uint8 view of the gathered list, total pairs, offset of this rank's shard). The view aliases the
    symmetric buffer and is overwritten by the next gather()."""

    HEADER = 4096

    def __init__(self, capacity_pairs: int, pair_bytes: int = 8, device=None, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _capi, api
        self._C, self._capi = C, _capi
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > _capi.MAX_PEERS:
            raise ValueError(f"at most {_capi.MAX_PEERS} ranks")
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.pair_bytes = int(pair_bytes)
        self.capacity_bytes = (int(capacity_pairs) * self.pair_bytes + 255) & ~255
        self.buf = symm.empty(self.HEADER + self.capacity_bytes, dtype=torch.uint8, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.buf[: self.HEADER].zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        self.handle = api.get_handle(self.device)
        self.peer = _capi.Peer()
        self.peer.rank, self.peer.world = self.rank, self.world
        for r in range(self.world):
            self.peer.buffers[r] = int(self.hdl.buffer_ptrs[r])
        mc = getattr(self.hdl, "multicast_ptr", 0) or 0
        self.peer.multicast = int(mc)
        self.peer.header_bytes, self.peer.capacity_bytes = self.HEADER, self.capacity_bytes
        self.epoch = 0
        self.fused_seq = 0

    def next_fused(self):
        """ctypes reference to the peer descriptor stamped for the next fused traversal (api.traverse(peer=...))."""
        self.epoch += 1
        self.fused_seq += 1
        self.peer.epoch, self.peer.fused_seq = self.epoch, self.fused_seq
        return self._C.pointer(self.peer)

    # -- regions of the fused traversals: rank r appends to [region_begin[r], region_begin[r + 1]) -----------------
    def last_counts(self) -> List[int]:
        """Per-rank counts of the last fused traversal (identical on every rank)."""
        C = self._C
        out = (C.c_int64 * self.world)()
        self._capi.lib().ibvh_peer_last_counts(self.handle, out, self.world)
        return [int(v) for v in out]

    def set_regions(self, counts: Optional[Sequence[int]] = None, slack: float = 0.06, pad: int = 4096) -> bool:
        """Size the ranks' regions from per-rank counts (+ slack, + `pad` entries each). Every rank must pass the same
        counts (e.g. last_counts()). None = equal split. Returns False if the list area cannot hold them."""
        cap = self.capacity_bytes // self.pair_bytes
        rb = self.peer.region_begin
        if counts is None:
            for r in range(self._capi.MAX_PEERS + 1):
                rb[r] = 0
            return True
        sizes = [(int(c * (1.0 + slack)) + pad + 1) & ~1 for c in counts]
        if sum(sizes) > cap:
            return False
        b = 0                                                    # tight regions (the slack is what the gap filling has to move);
        for r in range(self.world):                              # the unused room stays behind the last one
            rb[r] = b
            b += sizes[r]
        rb[self.world] = b
        return True

    def list_area(self) -> torch.Tensor:
        return self.buf[self.HEADER: self.HEADER + self.capacity_bytes]

    def gather(self, shard: torch.Tensor, count: int) -> Tuple[torch.Tensor, int, int]:
        C = self._C
        self.epoch += 1
        self.peer.epoch = self.epoch
        total, offset = C.c_int64(0), C.c_int64(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._capi.lib().ibvh_allgather_pairs(self.handle, C.byref(self.peer), shard.data_ptr() if count else None, int(count),
                                                   self.pair_bytes, C.byref(total), C.byref(offset), stream)
        if rc != 0:
            msg = self._capi.lib().ibvh_last_error(self.handle).decode()
            raise RuntimeError(f"ibvh_allgather_pairs: {self._capi.status_string(rc)}: {msg} (need {total.value} pairs)")
        return self.buf[self.HEADER: self.HEADER + total.value * self.pair_bytes], int(total.value), int(offset.value)


def concat_in_rank_order(parts: Sequence[np.ndarray]) -> np.ndarray:
    return np.concatenate(list(parts)) if len(parts) else np.empty(0)
