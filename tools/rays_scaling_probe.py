"""Ray traversal time against the number of rays on ONE GPU (the per-rank shard sizes of the 1/2/4/8-GPU bench), to tell a
fixed per-call cost from a multi-GPU effect:  python tools/rays_scaling_probe.py"""
import sys, torch
import numpy as np
sys.path.insert(0, ".")
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
shell = synth.shell_spheres_np(1000, 1000)
bvh = ib.BVH(shell, ib.BBox(), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for R, start in [(100_000_000, 0), (50_000_000, 50_000_000), (25_000_000, 75_000_000), (12_500_000, 0), (12_500_000, 87_500_000), (6_250_000, 0), (1_000_000, 0), (100_000, 0), (10_000, 0), (1_000, 0)]:
    rp, rd = synth.random_rays_torch(R, dev, seed=7, start=start)
    t = ib.traverse_rays(bvh, rp, rd, ordered=False, id_base=start)
    cache = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(int(t.num_contacts * 1.02) + 1024, ib.pair_dtype(), dev), t.cache2)
    for _ in range(2):
        ib.traverse_rays(bvh, rp, rd, cache=cache, ordered=False, id_base=start)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0.record()
        t = ib.traverse_rays(bvh, rp, rd, cache=cache, ordered=False, id_base=start)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ids = t.contacts.tensor.view(torch.int32).reshape(-1, 2)[:, 1].to(torch.int64) - 1 - start
    per_ray = torch.bincount(ids, minlength=R)
    top = torch.topk(per_ray, 5).values.tolist()
    print(f"   hits per ray: max {top}  rays with > 100 hits: {int((per_ray > 100).sum())}  > 1000: {int((per_ray > 1000).sum())}")
    print(f"R={R:>11} start={start:>10}: {np.median(ms):8.3f} ms (min {min(ms):.3f})  {R / np.median(ms) / 1e6:.3f} G rays/s  hits {t.num_contacts}", flush=True)
    if R <= 25_000_000:
        to = ib.traverse_rays(bvh, rp, rd, ordered=True, id_base=start)
        oc = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(int(to.num_contacts) + 1024, ib.pair_dtype(), dev), to.cache2)
        ms = []
        for _ in range(4):
            e0.record()
            to = ib.traverse_rays(bvh, rp, rd, cache=oc, ordered=True, id_base=start)
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        print(f"   ordered (count + scan + write): {np.median(ms[1:]):8.3f} ms", flush=True)
        del to, oc
    del rp, rd, cache, t
