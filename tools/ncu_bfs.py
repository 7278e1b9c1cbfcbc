"""One BFS contact traversal at 10 M leaves for  ncu --set full -k regex:bfs_ ...  (the first call sizes the lists, the second is profiled)."""
import sys, torch
sys.path.insert(0, ".")
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
vols = synth.random_spheres_torch(n, dev, seed=42)
src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
bvh = ib.BVH(src, ib.BBox())
tr = ib.traverse(bvh, ib.BFSTraversal())
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr = ib.traverse(bvh, ib.BFSTraversal(), cache=tr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(tr.num_contacts, tr.num_checks)
