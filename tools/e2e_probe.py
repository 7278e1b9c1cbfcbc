"""torchrun -N: per-phase timing of the end-to-end step at N > 1 (h2d of the rank's volume slice, NCCL all-gather of the
volumes, replicated build, fused traversal, d2h of the rank's slice of the gathered list) — diagnosis of bench.py's e2e."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ibvh_b200 as ib
from ibvh_b200 import dist as ibdist, synth
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 10_000_000
host = ib.DeviceArray.from_numpy(synth.random_spheres_np(n, seed=42), pin=True)
qb, qe = ibdist.shard_bounds(n, world)[rank]
lo, hi = qb * 16, qe * 16
d_in = ib.DeviceArray.empty(n, host.dtype, dev)
d_in.tensor.copy_(host.tensor)
bvh = ib.BVH(d_in, ib.BBox())
full = ib.traverse(bvh, ordered=False)
cap = int(full.num_contacts * 1.1) + 8192 * world
pg = ibdist.PeerGather(cap, 8, dev)
out = torch.empty(cap * 8, dtype=torch.uint8).pin_memory()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = np.zeros(5); wall = 0.0
for it in range(8):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev[0].record()
    d_in.tensor[lo:hi].copy_(host.tensor[lo:hi], non_blocking=True)
    ev[1].record()
    dist.all_gather_into_tensor(d_in.tensor, d_in.tensor[lo:hi])
    ev[2].record()
    bvh = ib.BVH(d_in, ib.BBox(), cache=bvh)
    ev[3].record()
    tr = ib.traverse(bvh, ordered=False, query_range=(qb, qe - qb), peer=pg)
    ev[4].record()
    tot = tr.num_contacts
    a, b = tot * rank // world, tot * (rank + 1) // world
    out[: (b - a) * 8].copy_(pg.list_area()[a * 8: b * 8], non_blocking=True)
    ev[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(5)]); wall += time.perf_counter() - t0
acc /= 5; wall /= 5
print(f"rank {rank}: h2d {acc[0]:.2f}  allgather {acc[1]:.2f}  build {acc[2]:.2f}  fused traverse {acc[3]:.2f}  d2h {acc[4]:.2f}  sum {acc.sum():.2f} ms  wall {wall * 1e3:.2f} ms", flush=True)
dist.barrier()
dist.destroy_process_group()
