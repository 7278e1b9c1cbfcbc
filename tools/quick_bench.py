#!/usr/bin/env python
"""Kernel-level timing harness (development aid, not the contract bench): per-kernel CUDA-event times
of build + traversal variants at several sizes."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibvh_b200 as ib
from ibvh_b200 import synth


def profile(handle):
    lib = ib.capi.lib()
    out = []
    name = C.create_string_buffer(64)
    ms = C.c_float()
    for i in range(lib.ibvh_profile_count(handle)):
        lib.ibvh_profile_get(handle, i, name, 64, C.byref(ms))
        out.append((name.value.decode(), ms.value))
    return out


def agg(rows):
    d = {}
    for k, v in rows:
        d.setdefault(k, []).append(v)
    return d


def timed(label, fn, h, reps, n, unit="leaves"):
    lib = ib.capi.lib()
    fn(); torch.cuda.synchronize()
    lib.ibvh_profile_enable(h, 1)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    dev_ms = e0.elapsed_time(e1) / reps
    rows = profile(h)
    lib.ibvh_profile_enable(h, 0)
    print(f"-- {label}: {dev_ms:.3f} ms/iter (wall {wall:.3f}) -> {n / dev_ms / 1e3:.1f} M {unit}/s", flush=True)
    for k, v in agg(rows).items():
        per = len(v) // reps
        print(f"     {k:26s} x{per:2d}/iter  {sum(v) / reps:8.3f} ms/iter")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,10000000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rays", type=int, default=0)
    ap.add_argument("--ref-shaped", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    for n in [int(x) for x in args.sizes.split(",")]:
        vols = synth.random_spheres_torch(n, dev, seed=42)
        src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
        bvh = ib.BVH(src, ib.BBox())
        h = bvh._handle
        tr = ib.traverse(bvh)
        C_ = tr.num_contacts
        big = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(C_ + 1024, ib.pair_dtype(), dev), tr.cache2)
        torch.cuda.synchronize()
        print(f"\n=== n={n} contacts={C_} ({C_ / n:.2f}/leaf) levels={bvh.tree.levels}", flush=True)
        cases = [
            ("build", lambda: ib.BVH(src, ib.BBox(), cache=bvh)),
            ("traverse ordered(pyramid)", lambda: ib.traverse(bvh, cache=big)),
            ("traverse unordered(pyramid)", lambda: ib.traverse(bvh, cache=big, ordered=False)),
            ("traverse ordered(walk)", lambda: ib.traverse(bvh, cache=big, walk=True)),
            ("traverse unordered(walk)", lambda: ib.traverse(bvh, cache=big, ordered=False, walk=True)),
            ("traverse ordered(packet)", lambda: ib.traverse(bvh, cache=big, packet=True)),
            ("traverse unordered(packet)", lambda: ib.traverse(bvh, cache=big, ordered=False, packet=True)),
        ]
        if args.ref_shaped:
            cases += [
                ("traverse ordered(ref-shaped)", lambda: ib.traverse(bvh, cache=big, reference_shaped=True)),
                ("traverse unordered(ref-shaped)", lambda: ib.traverse(bvh, cache=big, ordered=False, reference_shaped=True)),
            ]
        for label, fn in cases:
            timed(label, fn, h, args.reps, n)
        if args.rays:
            R = args.rays
            p, d = synth.random_rays_torch(R, dev, seed=7)
            p = (p * 0.4 + 0.5).contiguous()
            rt = ib.traverse_rays(bvh, p, d)
            bigr = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(rt.num_contacts + 1024, ib.pair_dtype(), dev), rt.cache2)
            print(f"   rays R={R} hits={rt.num_contacts}")
            timed("rays ordered", lambda: ib.traverse_rays(bvh, p, d, cache=bigr), h, args.reps, R, "rays")
            timed("rays unordered", lambda: ib.traverse_rays(bvh, p, d, cache=bigr, ordered=False), h, args.reps, R, "rays")


if __name__ == "__main__":
    main()
