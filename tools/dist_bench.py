"""Robustness of the traversal schedules to the leaf distribution (1 GPU): build + contact time per scene,
default (pyramid) vs packet schedule, unordered and ordered, plus the pair-list growth the pyramid learned."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ibvh_b200 as ib
from ibvh_b200 import synth

dev = torch.device("cuda", 0)
n = int(os.environ.get("LEAVES", 4_000_000))
g = torch.Generator(device=dev); g.manual_seed(5)


def spheres(c, r):
    return ib.DeviceArray(torch.cat([c, r[:, None]], 1).contiguous().view(torch.uint8).reshape(-1), ib.BSphere().dtype)


def scene(name):
    s = synth.sphere_radius_scale(n)
    if name == "uniform":
        c = torch.rand(n, 3, device=dev, generator=g); r = s * (0.5 + 0.5 * torch.rand(n, device=dev, generator=g))
    elif name == "shell":          # 2-D manifold: points on a unit sphere surface, radius for ~8 neighbours in 2-D
        v = torch.randn(n, 3, device=dev, generator=g); c = v / v.norm(dim=1, keepdim=True)
        r = (0.9 * (4 * np.pi / n) ** 0.5) * (0.5 + 0.5 * torch.rand(n, device=dev, generator=g))
    elif name == "blobs":          # 64 Gaussian clusters of very different density
        k = 64
        cen = torch.rand(k, 3, device=dev, generator=g)
        sig = 0.002 + 0.05 * torch.rand(k, device=dev, generator=g) ** 2
        a = torch.randint(0, k, (n,), device=dev, generator=g)
        c = cen[a] + sig[a][:, None] * torch.randn(n, 3, device=dev, generator=g)
        r = 0.6 * sig[a] * (n / k) ** (-1 / 3) * (0.5 + torch.rand(n, device=dev, generator=g))
    elif name == "mixed_radii":    # log-uniform radii over two decades
        c = torch.rand(n, 3, device=dev, generator=g)
        r = 0.25 * s * 10 ** (2 * torch.rand(n, device=dev, generator=g) - 1)
    elif name == "line":           # 1-D manifold with jitter
        t = torch.rand(n, device=dev, generator=g)
        c = torch.stack([t, 0.5 + 0.1 * torch.sin(20 * t), 0.5 + 0.1 * torch.cos(20 * t)], 1) + 1e-4 * torch.randn(n, 3, device=dev, generator=g)
        r = (2.0 / n) * (0.5 + torch.rand(n, device=dev, generator=g))
    return spheres(c.float(), r.float())


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


import ctypes as C
def prof_build(bvh_fn, handle):
    lib = ib.capi.lib()
    lib.ibvh_profile_reset(handle); lib.ibvh_profile_enable(handle, 1)
    for _ in range(3):
        bvh_fn()
    torch.cuda.synchronize()
    name = C.create_string_buffer(64); ms = C.c_float(); acc = {}
    for i in range(lib.ibvh_profile_count(handle)):
        lib.ibvh_profile_get(handle, i, name, 64, C.byref(ms))
        acc[name.value.decode()] = acc.get(name.value.decode(), 0.0) + ms.value / 3
    lib.ibvh_profile_enable(handle, 0)
    return {k: round(v, 3) for k, v in acc.items()}


for name in os.environ.get("SCENES", "uniform,shell,blobs,mixed_radii,line").split(","):
    src = scene(name)
    st = {"bvh": None}
    def build():
        st["bvh"] = ib.BVH(src, ib.BBox(), cache=st["bvh"]); return st["bvh"]
    tb, bvh = timed(build)
    print("   build kernels:", prof_build(build, bvh._handle), flush=True)
    res = {}
    for label, kw in (("pyr_unordered", dict(ordered=False)), ("pyr_ordered", dict(ordered=True)), ("packet_unordered", dict(ordered=False, packet=True))):
        c = {"tr": None}
        def trav():
            c["tr"] = ib.traverse(bvh, cache=c["tr"], **kw); return c["tr"]
        try:
            ms, tr = timed(trav, 3)
            res[label] = (round(ms, 3), tr.num_contacts)
        except Exception as ex:
            res[label] = ("ERR " + str(ex)[:80],)
    cs = {v[1] for v in res.values() if len(v) == 2}
    print(f"{name:12s} n={n} build {tb:.3f} ms  " + "  ".join(f"{k}={v[0]}" for k, v in res.items()) + f"  contacts={sorted(cs)}  aux={bvh._handle and 0}", flush=True)
