"""torchrun -N: checks ibvh_allgather_pairs against the NCCL gather on random shards, then times both."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ibvh_b200 as ib
from ibvh_b200 import dist as ibdist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
total_pairs = int(os.environ.get("PAIRS", 40_000_000))
pg = ibdist.PeerGather(int(total_pairs * 1.1) + 1024, 8, dev)
if rank == 0:
    print("multicast", hex(pg.peer.multicast), flush=True)
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
for trial, scale in enumerate([0, 1, 1000, total_pairs // world]):
    cnt = 0 if scale == 0 else int(scale + (rank * 7919 + trial * 13) % max(1, scale // 3 + 1))
    if trial == 1 and rank == world - 1:
        cnt = 0
    shard = torch.randint(0, 2**31 - 1, (max(cnt, 1) * 2,), dtype=torch.int32, device=dev, generator=g).view(torch.uint8)
    want, counts = ibdist.gather_shards(shard, cnt, 8)
    got, tot, off = pg.gather(shard, cnt)
    torch.cuda.synchronize()
    assert tot == sum(counts) and off == sum(counts[:rank]), (tot, off, counts)
    assert torch.equal(got, want), f"trial {trial} rank {rank}: payload mismatch"
    dist.barrier()
    if rank == 0:
        print("trial", trial, "counts", counts, "ok", flush=True)
# timing
cnt = total_pairs // world
shard = torch.randint(0, 2**31 - 1, (cnt * 2,), dtype=torch.int32, device=dev, generator=g).view(torch.uint8)
for name, fn in (("peer", lambda: pg.gather(shard, cnt)), ("nccl", lambda: ibdist.gather_shards(shard, cnt, 8))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name}: {t.item():.3f} ms per gather of {total_pairs * 8 / 1e6:.0f} MB over {world} ranks", flush=True)
dist.barrier()
dist.destroy_process_group()
