"""Driver entry points: build() compiles every native artefact in-tree; smoke() runs one small
invocation of the hot path on cuda:0 and checks it against the CPU oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "implicitbvh.jl_b200", "csrc")
ORACLE = os.path.join(ROOT, "oracle")


def build() -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (csrc/Makefile) -> lib/libibvh_b200.so,
    g++ (oracle/Makefile) -> oracle/_build/libibvh_oracle.so, then import the package."""
    subprocess.check_call(["make", "-C", CSRC, "-s"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", ORACLE, "-s"])     # building the checker is not using it
    sys.path.insert(0, ROOT)
    import ibvh_b200
    assert ibvh_b200.capi.lib().ibvh_version() == 100
    so = ibvh_b200.capi.LIB_PATH
    sass = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass, f"{so} holds no sm_100a cubin: {sass[:200]}"


def smoke() -> None:
    """One small build + LVT contact traversal + ray query on cuda:0 through the C ABI, checked bit for
    bit against the oracle (leaves, BBox nodes, contact lists in reference order)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, ORACLE)
    import numpy as np
    import torch
    import ibvh_b200 as ib
    from ibvh_b200 import synth
    import oracle as O

    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    dev = torch.device("cuda", 0)
    n = 20_000
    s = synth.random_spheres_np(n, seed=42)
    bvh = ib.BVH(s, ib.BBox(), device=dev)
    tr = ib.traverse(bvh)
    un = ib.traverse(bvh, ordered=False, cache=tr)
    got_unordered = un.contacts.numpy().copy()
    tr = ib.traverse(bvh, cache=un)
    p, d = synth.random_rays_np(5_000, seed=7)
    rays = ib.traverse_rays(bvh, (p * 0.4 + 0.5).T, d.T)
    torch.cuda.synchronize()

    ol = O.wrap(s)
    on, _, _ = O.build(ol, O.BBOX)
    assert bvh.leaves.numpy().tobytes() == ol.tobytes(), "sorted leaves differ from the oracle"
    assert bvh.nodes.numpy().tobytes() == on.tobytes(), "BBox nodes differ from the oracle"
    want = O.traverse_single(ol, on, num_threads=4)
    assert tr.contacts.numpy().tobytes() == want.tobytes(), "contact list differs from the oracle"
    key = lambda c: np.sort(c["a"].astype(np.int64) * (n + 1) + c["b"])
    assert (key(got_unordered) == key(want)).all(), "unordered contact set differs from the oracle"
    wr = O.traverse_rays(ol, on, (p * 0.4 + 0.5).T, d.T, num_threads=4)
    assert rays.contacts.numpy().tobytes() == wr.tobytes(), "ray hits differ from the oracle"
    print(f"smoke ok: n={n} contacts={tr.num_contacts} ray_hits={rays.num_contacts} "
          f"lib={ib.capi.LIB_PATH} device={torch.cuda.get_device_name(0)}")


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
