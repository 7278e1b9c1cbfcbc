import sys, torch
sys.path.insert(0, ".")
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
for n in (1_000_000, 10_000_000):
    vols = synth.random_spheres_torch(n, dev, seed=42)
    src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
    bvh = ib.BVH(src, ib.BBox())
    tr = ib.traverse(bvh)
    print(n, tr.num_contacts, flush=True)
