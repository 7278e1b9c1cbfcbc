import sys, torch, ctypes as C, time
sys.path.insert(0, ".")
import numpy as np
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
s = synth.shell_spheres_np(1000, 1000)
bvh = ib.BVH(s, ib.BBox())
p, d = synth.random_rays_torch(R, dev, seed=7)
t = ib.traverse_rays(bvh, p, d)
print("rays", R, "hits", t.num_contacts, "hits/ray", t.num_contacts / R, flush=True)
big = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(t.num_contacts + 1024, ib.pair_dtype(), dev), t.cache2)
lib = ib.capi.lib(); h = bvh._handle
for label, kw in (("ordered", dict()), ("unordered", dict(ordered=False))):
    ib.traverse_rays(bvh, p, d, cache=big, **kw); torch.cuda.synchronize()
    lib.ibvh_profile_enable(h, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ib.traverse_rays(bvh, p, d, cache=big, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    name = C.create_string_buffer(64); v = C.c_float(); agg = {}
    for i in range(lib.ibvh_profile_count(h)):
        lib.ibvh_profile_get(h, i, name, 64, C.byref(v)); agg[name.value.decode()] = agg.get(name.value.decode(), 0) + v.value / 3
    lib.ibvh_profile_enable(h, 0)
    print(f"{label}: {ms:.3f} ms -> {R / ms / 1e3:.1f} M rays/s", {k: round(x, 3) for k, x in agg.items()}, flush=True)
