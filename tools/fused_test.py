"""torchrun -N: fused traversal + all-gather (api.traverse(peer=...)) against the single-GPU contact list, then timing.
Exit code 0 = every check passed on every rank."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ibvh_b200 as ib
from ibvh_b200 import dist as ibdist, synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(os.environ.get("LEAVES", 2_000_000))
timing = os.environ.get("TIMING", "1") == "1"


def sorted_pairs(t, count):
    a = t[: count * 8].view(torch.int64)          # (a, b) int32 pair as one int64: sort order is irrelevant, equality is not
    return torch.sort(a).values


vols = synth.random_spheres_torch(n, dev, seed=7)
src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
bvh = ib.BVH(src, ib.BBox())
full = ib.traverse(bvh, ordered=False)
want = sorted_pairs(full.cache1.tensor, full.num_contacts)
pg = ibdist.PeerGather(int(full.num_contacts * 1.05) + 1024, 8, dev)
if rank == 0:
    print("world", world, "multicast", hex(pg.peer.multicast), "contacts", full.num_contacts, flush=True)
bounds = ibdist.shard_bounds(n, world)
qb, qe = bounds[rank]
ok = True
for rep in range(4):                                  # several epochs: exercises the counter rotation
    if pg.peer.multicast:
        tr = ib.traverse(bvh, ordered=False, query_range=(qb, qe - qb), peer=pg)
        got = sorted_pairs(tr.cache1.tensor, tr.num_contacts)
        ok &= tr.num_contacts == full.num_contacts and bool(torch.equal(got, want))
    # two-step path on the same buffer (shared epoch counter)
    tr2 = ib.traverse(bvh, ordered=True, query_range=(qb, qe - qb))
    lst, tot, off = pg.gather(tr2.cache1.tensor, tr2.num_contacts)
    ref = ib.traverse(bvh, ordered=True)
    ok &= tot == ref.num_contacts and bool(torch.equal(lst, ref.cache1.tensor[: tot * 8]))   # rank order == single-GPU order
# pair traversal, fused
vols2 = synth.random_spheres_torch(n // 2, dev, seed=8)
bvh2 = ib.BVH(ib.DeviceArray(vols2.view(torch.uint8).reshape(-1), ib.BSphere().dtype), ib.BBox())
fullp = ib.traverse(bvh, bvh2, ordered=False)
if pg.peer.multicast and fullp.num_contacts * 8 <= pg.capacity_bytes:
    trp = ib.traverse(bvh, bvh2, ordered=False, query_range=(qb, qe - qb), peer=pg)
    ok &= trp.num_contacts == fullp.num_contacts and bool(torch.equal(sorted_pairs(trp.cache1.tensor, trp.num_contacts), sorted_pairs(fullp.cache1.tensor, fullp.num_contacts)))
# rays: fused traversal + all-gather of the hits against the single-GPU hit list
R = 400_000
rp, rd = synth.random_rays_torch(R, dev, seed=3)
rp = rp * 0.3 + 0.5                                   # origins inside the unit cube of the sphere scene
full_r = ib.traverse_rays(bvh, rp, rd, ordered=False)
pg_r = ibdist.PeerGather(full_r.num_contacts + 1024, 8, dev)      # (tight on purpose: exercises the region re-sizing)
rb = ibdist.shard_bounds(R, world)[rank]
if pg_r.peer.multicast:
    for rep in range(2):
        fr = ib.traverse_rays(bvh, rp[rb[0]:rb[1]], rd[rb[0]:rb[1]], ordered=False, id_base=rb[0], peer=pg_r)
        ok &= fr.num_contacts == full_r.num_contacts and bool(torch.equal(sorted_pairs(fr.cache1.tensor, fr.num_contacts), sorted_pairs(full_r.cache1.tensor, full_r.num_contacts)))
    if rank == 0:
        print("rays fused:", fr.num_contacts, "hits", "ok" if ok else "MISMATCH", flush=True)

# Int64 indices / UInt64 Morton codes: 16-byte pairs through the same exchange
opt64 = ib.BVHOptions(index=np.int64, morton=ib.DefaultMortonAlgorithm(np.uint64))
n64 = max(1000, n // 4)
v64 = synth.random_spheres_torch(n64, dev, seed=9)
bvh64 = ib.BVH(ib.DeviceArray(v64.view(torch.uint8).reshape(-1), ib.BSphere().dtype), ib.BBox(), options=opt64)
full64 = ib.traverse(bvh64, ordered=True)
pg64 = ibdist.PeerGather(full64.num_contacts + 1024, 16, dev)
b64 = ibdist.shard_bounds(n64, world)[rank]
t64 = ib.traverse(bvh64, ordered=True, query_range=(b64[0], b64[1] - b64[0]))
lst, tot, off = pg64.gather(t64.cache1.tensor, t64.num_contacts)
ok &= tot == full64.num_contacts and bool(torch.equal(lst, full64.cache1.tensor[: tot * 16]))
if pg64.peer.multicast:
    f64 = ib.traverse(bvh64, ordered=False, query_range=(b64[0], b64[1] - b64[0]), peer=pg64)
    a = f64.cache1.tensor[: f64.num_contacts * 16].view(torch.int64).reshape(-1, 2)
    b = full64.cache1.tensor[: tot * 16].view(torch.int64).reshape(-1, 2)
    ok &= f64.num_contacts == tot and bool(torch.equal(torch.sort(a[:, 0] * (n64 + 1) + a[:, 1]).values, torch.sort(b[:, 0] * (n64 + 1) + b[:, 1]).values))
# capacity error is reported on every rank
small = ibdist.PeerGather(1000, 8, dev)
if small.peer.multicast:
    try:
        ib.traverse(bvh, ordered=False, query_range=(qb, qe - qb), peer=small)
        ok = False
    except Exception as ex:
        ok &= "too small" in str(ex) or "capacity" in str(ex).lower()
        if rank == 0:
            print("capacity error text:", str(ex)[:120], flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PARITY", "ok" if t.item() == 1 else "FAILED", flush=True)
if timing and t.item() == 1:
    def fused():
        ib.traverse(bvh, ordered=False, query_range=(qb, qe - qb), peer=pg)
    cache = {"tr": None}
    def twostep():
        cache["tr"] = ib.traverse(bvh, ordered=False, query_range=(qb, qe - qb), cache=cache["tr"])
        pg.gather(cache["tr"].cache1.tensor, cache["tr"].num_contacts)
    for name, fn in (("fused", fused), ("two-step", twostep)):
        if name == "fused" and not pg.peer.multicast:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name}: {tt.item():.3f} ms per sharded traversal + gather ({n} leaves, {world} ranks)", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
