#!/usr/bin/env python
"""BFSTraversal timing (development aid): per-kernel CUDA-event times of the BFS contact / ray traversals beside the LVT
ones on the same trees, with the result sets compared.   python tools/bfs_bench.py --sizes 1000000,10000000"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ibvh_b200 as ib
from ibvh_b200 import synth
from quick_bench import timed


def key(tr, n):
    t = tr.contacts.tensor.view(torch.int32).reshape(-1, 2).to(torch.int64)
    return torch.sort(t[:, 0] * (n + 1) + t[:, 1]).values


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,10000000")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--rays", type=int, default=2_000_000)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    for n in [int(x) for x in args.sizes.split(",")]:
        vols = synth.random_spheres_torch(n, dev, seed=42)
        src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
        bvh = ib.BVH(src, ib.BBox())
        h = bvh._handle
        lvt = ib.traverse(bvh, ordered=False)
        bfs = ib.traverse(bvh, ib.BFSTraversal())
        same = bool((key(lvt, n) == key(bfs, n)).all()) if lvt.num_contacts == bfs.num_contacts else False
        print(f"\n=== n={n} levels={bvh.tree.levels} contacts LVT={lvt.num_contacts} BFS={bfs.num_contacts} same_set={same} "
              f"num_checks BFS={bfs.num_checks} ({bfs.num_checks / max(1, bfs.num_contacts):.1f} per contact) LVT(pyramid tests)={lvt.num_checks} "
              f"device memory in use {torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9:.1f} GB", flush=True)
        timed("BFS traverse (default start level)", lambda: ib.traverse(bvh, ib.BFSTraversal(), cache=bfs), h, args.reps, n)
        timed("LVT traverse unordered", lambda: ib.traverse(bvh, cache=lvt, ordered=False), h, args.reps, n)
        del bvh, lvt, bfs
        ib.capi.lib().ibvh_release_workspace(h)
        torch.cuda.empty_cache()
    if args.rays:
        n = 1_000_000
        s = synth.shell_spheres_np(1000, 1000)
        bvh = ib.BVH(s, ib.BBox(), device=dev)
        p, d = synth.random_rays_torch(args.rays, dev, seed=7)
        lv = ib.traverse_rays(bvh, p, d, ordered=False)
        bf = ib.traverse_rays(bvh, p, d, ib.BFSTraversal())
        same = bool((key(lv, args.rays) == key(bf, args.rays)).all()) if lv.num_contacts == bf.num_contacts else False
        print(f"\n=== rays: {args.rays} rays, {n} leaves: hits LVT={lv.num_contacts} BFS={bf.num_contacts} same_set={same} num_checks={bf.num_checks}", flush=True)
        timed("BFS traverse_rays", lambda: ib.traverse_rays(bvh, p, d, ib.BFSTraversal(), cache=bf), bvh._handle, args.reps, args.rays, unit="rays")
        timed("LVT traverse_rays unordered", lambda: ib.traverse_rays(bvh, p, d, cache=lv, ordered=False), bvh._handle, args.reps, args.rays, unit="rays")


if __name__ == "__main__":
    main()
