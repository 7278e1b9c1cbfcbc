#!/usr/bin/env python
"""Traversal work counters (node tests, leaf tests, steps) for the packet and reference-shaped schedules."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibvh_b200 as ib
from ibvh_b200 import synth

dev = torch.device("cuda", 0)
lib = ib.capi.lib()
for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1000000,10000000").split(",")]:
    vols = synth.random_spheres_torch(n, dev, seed=42)
    src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
    bvh = ib.BVH(src, ib.BBox())
    cb = bvh._c_bvh()
    counts = ib.DeviceArray.empty(n, "int32", dev)
    for name, flags in (("packet", ib.capi.TRAVERSE_ORDERED | 8), ("thread", ib.capi.TRAVERSE_ORDERED | ib.capi.TRAVERSE_REFERENCE_SHAPED | 8)):
        params = ib.capi.TraverseParams(1, 0, -1, flags, 0, 0)
        total = C.c_int64()
        rc = lib.ibvh_traverse_single(bvh._handle, C.byref(cb), C.byref(params), counts.ptr, None, 0, C.byref(total),
                                      torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
        out = (C.c_int64 * 4)()
        lib.ibvh_last_traversal_stats(bvh._handle, out)
        nt, lt, s2, s3 = [int(v) for v in out]
        if name == "packet":
            print(f"n={n} packet: contacts={total.value} node_tests/query={nt / n:.1f} leaf_tests/query={lt / n:.1f} "
                  f"warp_steps/warp={s2 / (n / 32):.1f} warp_loads/warp={s3 / (n / 32):.1f}")
        else:
            print(f"n={n} thread: contacts={total.value} node_tests/query={nt / n:.1f} leaf_tests/query={lt / n:.1f} "
                  f"steps/query={s2 / n:.1f} slowest_lane_steps/warp={s3 / (n / 32):.1f}")
