"""Host-side multi-GPU logic on CPU: world_size-2 gloo. Shards are fake contact lists; the product's
gather / shard-bound code is what runs (no CUDA kernels involved)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ibvh_b200 import dist as ibdist
    try:
        # 1. contiguous query shards cover [0, n) exactly once, in rank order
        n = 1003
        bounds = ibdist.shard_bounds(n, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == n and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        # 2. every rank holds the contacts of its query range (fake: pair = (q, q + k)); gather them
        b, e = bounds[rank]
        per_query = [(q % 4) for q in range(n)]
        mine = np.array([(q + 1, q + 1 + k) for q in range(b, e) for k in range(1, per_query[q] + 1)], np.int32).reshape(-1, 2)
        shard = torch.from_numpy(mine.copy().reshape(-1).view(np.uint8))
        pad = torch.zeros(shard.numel() + 64, dtype=torch.uint8)        # buffer larger than the valid part, as cache1 is
        pad[: shard.numel()] = shard
        full, counts = ibdist.gather_shards(pad, len(mine), 8)
        want = np.array([(q + 1, q + 1 + k) for q in range(n) for k in range(1, per_query[q] + 1)], np.int32).reshape(-1, 2)
        got = full.numpy().view(np.int32).reshape(-1, 2)
        assert sum(counts) == len(want) and (got == want).all(), "concatenation in rank order must equal the single-rank list"
        # 3. an empty shard on one rank
        full, counts = ibdist.gather_shards(pad, len(mine) if rank == 0 else 0, 8)
        assert counts[1] == 0 and full.numel() == counts[0] * 8
        # 4. everything empty
        full, counts = ibdist.gather_shards(pad, 0, 8)
        assert full.numel() == 0 and counts == [0, 0]
        # 5. broadcast of the "tree" from rank 0
        t = torch.arange(100, dtype=torch.uint8) if rank == 0 else torch.zeros(100, dtype=torch.uint8)
        ibdist.broadcast_tensor_(t, src=0)
        assert (t == torch.arange(100, dtype=torch.uint8)).all()
        q.put((rank, "ok"))
    except Exception as ex:      # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_gather_and_shards_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_weighted_shard_bounds():
    import ibvh_b200  # noqa: F401
    from ibvh_b200 import dist as ibdist
    rng = np.random.default_rng(0)
    w = rng.integers(0, 10, 10_000).astype(np.float64)
    for world in (1, 2, 4, 8):
        b = ibdist.shard_bounds(len(w), world, w)
        assert b[0][0] == 0 and b[-1][1] == len(w) and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        loads = [w[s:e].sum() for s, e in b]
        assert max(loads) <= 1.05 * w.sum() / world + 10
    assert ibdist.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert ibdist.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
