"""One build + contact traversal (unordered and ordered) + ray traversal, bracketed by cudaProfilerStart/Stop, for
    ncu --set full --profile-from-start off --clock-control none --import-source on -o out python tools/ncu_step.py [n]
(every kernel of the hot path once, after a warm-up pass outside the profiled range)."""
import sys, torch
sys.path.insert(0, ".")
import ibvh_b200 as ib
from ibvh_b200 import synth
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
vols = synth.random_spheres_torch(n, dev, seed=42)
src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
shell = synth.shell_spheres_np(1000, 1000)
rp, rd = synth.random_rays_torch(4_000_000, dev, seed=7)


def step(cache_u=None, cache_o=None, cache_r=None):
    bvh = ib.BVH(src, ib.BBox())
    un = ib.traverse(bvh, ordered=False, cache=cache_u)
    od = ib.traverse(bvh, cache=cache_o)
    rbvh = ib.BVH(shell, ib.BBox(), device=dev)
    rt = ib.traverse_rays(rbvh, rp, rd, ordered=False, cache=cache_r)
    torch.cuda.synchronize()
    return un, od, rt


un, od, rt = step()
un, od, rt = step(un, od, rt)
torch.cuda.profiler.start()
step(un, od, rt)
torch.cuda.profiler.stop()
print("contacts", un.num_contacts, od.num_contacts, "ray hits", rt.num_contacts)
