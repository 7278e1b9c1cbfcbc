import sys, torch
sys.path.insert(0, ".")
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
vols = synth.random_spheres_torch(n, dev, seed=42)
src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
bvh = ib.BVH(src, ib.BBox())
tr = ib.traverse(bvh)
big = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(tr.num_contacts + 1024, ib.pair_dtype(), dev), tr.cache2)
ib.traverse(bvh, cache=big, ordered=False)
torch.cuda.synchronize()
