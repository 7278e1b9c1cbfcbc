import sys, torch, ctypes as C, subprocess, time
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
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
def prof(label, fn, reps=3):
    lib.ibvh_profile_enable(h, 1)
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    name = C.create_string_buffer(64); ms = C.c_float(); agg = {}
    for i in range(lib.ibvh_profile_count(h)):
        lib.ibvh_profile_get(h, i, name, 64, C.byref(ms)); agg[name.value.decode()] = agg.get(name.value.decode(), 0) + ms.value / reps
    lib.ibvh_profile_enable(h, 0)
    print(label, {k: round(v, 3) for k, v in agg.items() if "refine" in k or "tile" in k or "gather" in k or "onesweep" in k}, clocks(), flush=True)
f_un = lambda: ib.traverse(bvh, cache=big, ordered=False)
prof("cold unordered", f_un)
t0 = time.time()
while time.time() - t0 < 3.0: f_un()
prof("after 3s of unordered", f_un)
f_build = lambda: ib.BVH(src, ib.BBox(), cache=bvh)
prof("build", f_build)
prof("unordered after build", f_un)
f_or = lambda: ib.traverse(bvh, cache=big)
prof("ordered", f_or)
prof("unordered after ordered", f_un)
