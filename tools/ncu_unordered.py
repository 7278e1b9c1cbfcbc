import sys, torch
sys.path.insert(0, ".")
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
vols = synth.random_spheres_torch(n, dev, seed=42)
src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
bvh = ib.BVH(src, ib.BBox())
big = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(int(n * 4.2) + 1024, ib.pair_dtype(), dev), ib.DeviceArray.empty(n, "int32", dev))
for _ in range(2):                      # first call learns the pair-list sizes, second is the steady state
    tr = ib.traverse(bvh, cache=big, ordered=False)
torch.cuda.synchronize()
print(tr.num_contacts)
