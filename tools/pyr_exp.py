import sys, torch, ctypes as C
sys.path.insert(0, ".")
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
n = 10_000_000
vols = synth.random_spheres_torch(n, dev, seed=42)
src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
bvh = ib.BVH(src, ib.BBox())
tr = ib.traverse(bvh)
big = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(tr.num_contacts + 1024, ib.pair_dtype(), dev), tr.cache2)
lib = ib.capi.lib(); h = bvh._handle
for it in range(2):
    ib.traverse(bvh, cache=big, ordered=False)
torch.cuda.synchronize()
lib.ibvh_profile_enable(h, 1)
for it in range(3):
    ib.traverse(bvh, cache=big, ordered=False)
torch.cuda.synchronize()
name = C.create_string_buffer(64); ms = C.c_float(); agg = {}
for i in range(lib.ibvh_profile_count(h)):
    lib.ibvh_profile_get(h, i, name, 64, C.byref(ms)); agg.setdefault(name.value.decode(), []).append(ms.value)
print({k: [round(x, 3) for x in v[-5:]] for k, v in agg.items() if "refine" in k or "tile" in k})
