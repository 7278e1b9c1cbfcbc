"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's known answers.

Bar (BASELINE.json north_star): Morton codes, sort order, BBox node volumes and contact lists are
BIT-EXACT; BSphere node merges within 1e-6 relative (we check exact equality first and fall back to
the tolerance). Run on the B200 box: `pytest -m gpu`.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import pairs_list, random_spheres, sorted_pairs

pytestmark = pytest.mark.gpu

F32_RTOL = 1e-6      # north_star tolerance for BSphere node merges


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def I_of(ibytes):
    return {4: np.int32, 8: np.int64}[ibytes]


def M_of(mbytes):
    return {2: np.uint16, 4: np.uint32, 8: np.uint64}[mbytes]


def opts(ib, ibytes=4, mbytes=4, **kw):
    return ib.BVHOptions(index=I_of(ibytes), morton=ib.DefaultMortonAlgorithm(M_of(mbytes), **kw))


def gpu_build(ib, vols, node="bbox", ibytes=4, mbytes=4, built_level=1, **kw):
    nt = ib.BBox() if node == "bbox" else ib.BSphere()
    return ib.BVH(vols, nt, built_level=built_level, options=opts(ib, ibytes, mbytes), **kw)


def oracle_build(O, vols, node="bbox", ibytes=4, mbytes=4, built_level=1):
    leaves = O.wrap(vols, ibytes, mbytes)
    nodes, mn, mx = O.build(leaves, O.BBOX if node == "bbox" else O.BSPHERE, built_level=built_level)
    return leaves, nodes


def assert_nodes_equal(got, want, node, lo=0):
    got, want = got[lo:], want[lo:]
    if node == "bbox":
        assert got.tobytes() == want.tobytes(), "BBox nodes must be bit-exact"
    else:
        if got.tobytes() != want.tobytes():
            np.testing.assert_allclose(got["x"], want["x"], rtol=F32_RTOL, atol=0)
            np.testing.assert_allclose(got["r"], want["r"], rtol=F32_RTOL, atol=0)


# ---------------------------------------------------------------------------------------------
def test_extension_is_loaded_not_a_fallback(ib, dev):
    import torch
    assert ib.capi.lib().ibvh_version() == 100
    h = ib.get_handle(dev)
    assert h.value
    maps = open("/proc/self/maps").read()
    assert "libibvh_b200.so" in maps
    assert torch.cuda.get_device_capability(0)[0] >= 10, "built for sm_100a only"


def test_synth_device_matches_host(ib, dev):
    from ibvh_b200 import synth
    n = 100_003
    a = synth.random_spheres_np(n, seed=42)
    b = synth.random_spheres_torch(n, dev, seed=42).cpu().numpy()
    assert (a["x"] == b[:, :3]).all() and (a["r"] == b[:, 3]).all()
    p, d = synth.random_rays_np(1000, seed=7)
    pt, _ = synth.random_rays_torch(1000, dev, seed=7)
    assert (p == pt.cpu().numpy()).all()


# ---- the reference's doctests / known answers through the product API --------------------------
def test_five_spheres_doctest(ib, golden, dev):
    g = golden["five_spheres"]
    for ibytes, mbytes in ((4, 4), (8, 8), (4, 2), (4, 8), (8, 4)):
        bvh = gpu_build(ib, ib.bspheres(g["centers"], g["radii"]), "bbox", ibytes, mbytes)
        tr = ib.traverse(bvh)
        assert pairs_list(tr.contacts.numpy()) == [tuple(p) for p in g["contacts_lvt_order"]]
        assert tr.contacts.numpy().dtype == ib.pair_dtype(I_of(ibytes))
        tr2 = ib.traverse(bvh, cache=tr)                     # traverse.jl:168-169
        assert pairs_list(tr2.contacts.numpy()) == [tuple(p) for p in g["contacts_lvt_order"]]
        assert tr2.cache1.ptr == tr.cache1.ptr and tr2.cache2.ptr == tr.cache2.ptr
    bvh = gpu_build(ib, ib.bspheres(g["centers"], g["radii"]), "sphere")
    assert pairs_list(ib.traverse(bvh).contacts.numpy()) == [tuple(p) for p in g["contacts_lvt_order"]]


def test_shuffled_doctest_inplace(ib, golden, dev):
    g = golden["five_spheres_shuffled"]
    ldt = ib.leaf_dtype(ib.BSphere())
    leaves = np.zeros(5, ldt)
    leaves["volume"] = ib.bspheres(g["centers"], g["radii"])
    leaves["index"] = g["indices"]
    d = ib.DeviceArray.from_numpy(leaves, device=dev)
    bvh = ib.BVH(d, ib.BBox())
    assert bvh.leaves.ptr == d.ptr                        # modified in place, no extra allocation
    first = d.numpy()[0]
    want = g["first_sorted_leaf"]
    assert int(first["index"]) == want["index"] and int(first["morton"]) == int(want["morton"], 16)
    assert first["volume"]["x"].tolist() == want["x"] and float(first["volume"]["r"]) == want["r"]


def test_pair_and_ray_doctests(ib, golden, dev):
    g = golden["pair_example"]
    b1 = gpu_build(ib, ib.bspheres(g["centers1"], g["radii1"]))
    b2 = gpu_build(ib, ib.bspheres(g["centers2"], g["radii2"]))
    tr = ib.traverse(b1, b2, start_level1=g["start_level1"], start_level2=g["start_level2"])
    assert pairs_list(tr.contacts.numpy()) == [tuple(p) for p in g["contacts_lvt_order"]]
    assert (tr.start_level1, tr.start_level2) == (2, 3)
    tr = ib.traverse(b1, b2, cache=tr)
    assert pairs_list(tr.contacts.numpy()) == [tuple(p) for p in g["contacts_lvt_order"]]
    g5, gr = golden["five_spheres"], golden["ray_example"]
    bvh = gpu_build(ib, ib.bspheres(g5["centers"], g5["radii"]))
    tr = ib.traverse_rays(bvh, np.array(gr["points"]), np.array(gr["directions"]))
    assert pairs_list(tr.contacts.numpy()) == [tuple(p) for p in gr["contacts_lvt_order"]]
    tr = ib.traverse_rays(bvh, np.array(gr["points"]), np.array(gr["directions"]), cache=tr)
    assert pairs_list(tr.contacts.numpy()) == [tuple(p) for p in gr["contacts_lvt_order"]]


def test_unordered_structure(ib, O, golden, dev):
    g = golden["unordered_contacts"]
    s = ib.bspheres(g["centers"], g["radii"])
    for node in ("bbox", "sphere"):
        bvh = gpu_build(ib, s, node)
        assert len(bvh.nodes) == g["num_nodes"]
        ol, on = oracle_build(O, s, node)
        assert bvh.leaves.numpy().tobytes() == ol.tobytes()
        assert_nodes_equal(bvh.nodes.numpy(), on, node)
        assert set(pairs_list(ib.traverse(bvh).contacts.numpy())) == {tuple(p) for p in g["contacts_set"]}
    b = O.boxes_of_spheres(s)
    bvh = gpu_build(ib, b, "bbox")
    assert set(pairs_list(ib.traverse(bvh).contacts.numpy())) == {tuple(p) for p in g["contacts_set"]}


# ---- Morton codes: bit-exact, n in 1:200 (test/gputests.jl:34-48) ------------------------------
@pytest.mark.parametrize("mbytes", [2, 4, 8])
@pytest.mark.parametrize("leaf", ["sphere", "box"])
def test_morton_bit_exact_small_sweep(ib, O, dev, mbytes, leaf):
    rng = np.random.default_rng(42)
    for n in range(1, 201):
        s = random_spheres(rng, n)
        vols = s if leaf == "sphere" else O.boxes_of_spheres(s)
        want = O.wrap(vols, 4, mbytes)
        mn, mx = O.morton_encode(want)
        d = ib.wrap_bounding_volumes(vols, opts(ib, 4, mbytes))
        gmn, gmx = ib.morton_encode(d, opts(ib, 4, mbytes))
        got = d.numpy()
        assert got.tobytes() == want.tobytes(), (n, mbytes, leaf)
        assert (gmn.astype(np.float32) == mn).all() and (gmx.astype(np.float32) == mx).all()


def test_morton_quirks_and_user_bounds(ib, O, dev):
    # all-negative axis (max seeded with floatmin), duplicates, single leaf, huge magnitudes, denormal extents
    cases = [
        ib.bspheres([[-5, -7, -9], [-1, -2, -3], [-4, -4, -4]], [1, 1, 1]),
        ib.bspheres([[3, 3, 3]] * 7, [1] * 7),
        ib.bspheres([[0, 0, 0]], [0.5]),
        ib.bspheres([[1e30, -1e30, 1e-30], [2e30, 1e30, 3e-30], [-1e30, 0, 0]], [1, 1, 1]),
        ib.bspheres([[0, 0, 0], [1e-40, 2e-40, 3e-40], [5e-39, 5e-39, 5e-39]], [1, 1, 1]),
    ]
    for s in cases:
        for mbytes in (2, 4, 8):
            want = O.wrap(s, 4, mbytes)
            O.morton_encode(want)
            d = ib.wrap_bounding_volumes(s, opts(ib, 4, mbytes))
            ib.morton_encode(d, opts(ib, 4, mbytes))
            assert d.numpy().tobytes() == want.tobytes()
    # user bounds (compute_extrema=false): used as given, unpadded
    rng = np.random.default_rng(1)
    s = random_spheres(rng, 1000)
    o = opts(ib, 4, 4, compute_extrema=False, mins=(-1.0, -1.0, -1.0), maxs=(7.0, 7.5, 8.0))
    want = O.wrap(s)
    O.morton_encode(want, compute_extrema=False, mins=(-1, -1, -1), maxs=(7, 7.5, 8))
    d = ib.wrap_bounding_volumes(s, o)
    ib.morton_encode(d, o)
    assert d.numpy().tobytes() == want.tobytes()
    with pytest.raises(ib.ArgumentError):
        ib.morton_encode(d, opts(ib, 4, 8))               # morton type mismatch, morton.jl:38-42


# ---- build: sort order + nodes -----------------------------------------------------------------
@pytest.mark.parametrize("node", ["bbox", "sphere"])
@pytest.mark.parametrize("ibytes,mbytes", [(4, 4), (8, 8), (4, 2)])
def test_build_bit_exact_sizes(ib, O, dev, node, ibytes, mbytes):
    rng = np.random.default_rng(7)
    for n in [1, 2, 3, 4, 5, 31, 32, 33, 255, 1023, 1024, 1025, 2049, 4096, 4097, 12345, 70001]:
        s = random_spheres(rng, n, spread=6.0 * max(1.0, (n / 200.0) ** (1 / 3)))
        ol, on = oracle_build(O, s, node, ibytes, mbytes)
        bvh = gpu_build(ib, s, node, ibytes, mbytes)
        assert bvh.leaves.numpy().tobytes() == ol.tobytes(), (n, "sorted leaves")
        assert_nodes_equal(bvh.nodes.numpy(), on, node)
        assert bvh.tree.levels == O.tree_shape(n)["levels"]
        assert (bvh.skips.cpu().numpy() == O.tree_shape(n)["skips"]).all()


def test_build_box_leaves(ib, O, dev):
    rng = np.random.default_rng(8)
    for n in (1, 6, 1000, 5000):
        b = O.boxes_of_spheres(random_spheres(rng, n))
        for ibytes, mbytes in ((4, 4), (8, 8), (4, 8), (8, 2)):
            ol, on = oracle_build(O, b, "bbox", ibytes, mbytes)
            bvh = gpu_build(ib, b, "bbox", ibytes, mbytes)
            assert bvh.leaves.numpy().tobytes() == ol.tobytes()
            assert_nodes_equal(bvh.nodes.numpy(), on, "bbox")
    with pytest.raises(ib.ArgumentError):
        gpu_build(ib, b, "sphere")                       # no BSphere(::BBox) in the reference


def test_build_many_ties_is_stable(ib, O, dev):
    """UInt16 codes (15 bits) on 100k leaves: ~3 leaves per code. Stable tie rule."""
    rng = np.random.default_rng(9)
    s = random_spheres(rng, 100_000, spread=50.0)
    ol, on = oracle_build(O, s, "bbox", 4, 2)
    bvh = gpu_build(ib, s, "bbox", 4, 2)
    got = bvh.leaves.numpy()
    assert (np.diff(got["morton"].astype(np.int64)) >= 0).all()
    assert got.tobytes() == ol.tobytes()
    assert_nodes_equal(bvh.nodes.numpy(), on, "bbox")
    # all-identical keys: order must be the input order
    s = ib.bspheres(np.full((5000, 3), 0.5), np.full(5000, 0.01))
    bvh = gpu_build(ib, s)
    assert (bvh.leaves.numpy()["index"] == np.arange(1, 5001)).all()


def test_built_level_and_cache(ib, O, dev):
    """build.jl:309-325 + cache rules build.jl:232-238,257-263; test/runtests.jl:904-918."""
    rng = np.random.default_rng(10)
    s = random_spheres(rng, 100)
    full = gpu_build(ib, s)
    levels = full.tree.levels
    ol, on = oracle_build(O, s)
    for bl, want in ((3, 3), (0.0, levels), (0.5, int(np.rint(levels + (1 - levels) * 0.5))), (1.0, 1)):
        import torch
        pre = ib.DeviceArray(torch.full((len(full.nodes) * 24,), 0xAB, dtype=torch.uint8, device=dev), ib.BBox().dtype)
        holder = gpu_build(ib, s)           # donor BVH whose node buffer we pre-fill to detect stray writes
        holder.nodes.tensor.copy_(pre.tensor)
        bvh = ib.BVH(s, ib.BBox(), built_level=bl, cache=holder)
        assert bvh.built_level == want
        assert bvh.nodes.ptr == holder.nodes.ptr      # node buffer reused
        got = bvh.nodes.numpy()
        lo = O.level_indices(100, min(want, levels - 1))[0] - 1
        assert got[lo:].tobytes() == on[lo:].tobytes()
        assert (got[:lo].view(np.uint8) == 0xAB).all(), "levels above built_level must stay untouched"
        sl = max(1, want)
        c = ib.traverse(bvh, start_level=sl)
        assert (sorted_pairs(c.contacts.numpy()) == sorted_pairs(O.traverse_single(ol, on))).all()
    with pytest.raises(ib.ArgumentError):
        ib.BVH(s, ib.BSphere(), cache=full)              # node type mismatch with the cache
    with pytest.raises(ib.ArgumentError):
        ib.BVH(s, ib.BBox(), built_level=0)
    with pytest.raises(ib.ArgumentError):
        ib.BVH(s, ib.BBox(), built_level=levels + 1)
    with pytest.raises(ib.ArgumentError):
        ib.BVH(s, ib.BBox(), built_level=1.5)
    with pytest.raises(ib.ArgumentError):
        ib.traverse(ib.BVH(s, ib.BBox(), built_level=4), start_level=2)   # start_level < built_level
    # wrapped input with mismatching index / morton types (runtests.jl:924-930)
    w64 = ib.wrap_bounding_volumes(s, opts(ib, 8, 4))
    with pytest.raises(ib.ArgumentError):
        ib.BVH(w64, ib.BBox())
    wm = ib.wrap_bounding_volumes(s, opts(ib, 4, 8))
    with pytest.raises(ib.ArgumentError):
        ib.BVH(wm, ib.BBox())


def test_stage_entry_points(ib, O, dev):
    rng = np.random.default_rng(11)
    s = random_spheres(rng, 3000)
    want = O.wrap(s)
    O.morton_encode(want)
    d = ib.wrap_bounding_volumes(s)
    assert d.numpy().tobytes() == O.wrap(s).tobytes()
    ib.morton_encode(d)
    assert d.numpy().tobytes() == want.tobytes()
    O.sort_leaves(want)
    ib.sort_leaves(d)
    assert d.numpy().tobytes() == want.tobytes()
    for node, nt in (("bbox", ib.BBox()), ("sphere", ib.BSphere())):
        for bl in (1, 5):
            on = O.aggregate(want, O.BBOX if node == "bbox" else O.BSPHERE, built_level=bl)
            gn = ib.aggregate(d, nt, built_level=bl).numpy()
            lo = O.level_indices(3000, bl)[0] - 1
            assert_nodes_equal(gn, on, node, lo)


def test_rebuild_is_idempotent_and_cached(ib, O, dev):
    """Config-5 shape: BVH(bvh.leaves, cache=bvh) on already sorted leaves gives the same tree."""
    rng = np.random.default_rng(12)
    s = random_spheres(rng, 20_000, spread=30.0)
    bvh = gpu_build(ib, s, "bbox", 8, 8)
    before_l, before_n = bvh.leaves.numpy().copy(), bvh.nodes.numpy().copy()
    again = ib.BVH(bvh.leaves, ib.BBox(), cache=bvh, options=opts(ib, 8, 8))
    assert again.leaves.ptr == bvh.leaves.ptr and again.nodes.ptr == bvh.nodes.ptr
    assert again.leaves.numpy().tobytes() == before_l.tobytes()
    assert again.nodes.numpy().tobytes() == before_n.tobytes()


# ---- traversal ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("node", ["bbox", "sphere"])
def test_single_all_start_levels(ib, O, dev, node):
    """gputests.jl:51-127 + runtests.jl:839-900: n in 1:11:200, every start level; exact reference ORDER
    in ordered mode, same set in unordered / reference-shaped mode, and == brute force."""
    rng = np.random.default_rng(42)
    for n in range(1, 200, 11):
        s = random_spheres(rng, n)
        ol, on = oracle_build(O, s, node)
        bvh = gpu_build(ib, s, node)
        brute = sorted_pairs(O.brute_single(s))
        for sl in range(1, bvh.tree.levels + 1):
            want = O.traverse_single(ol, on, start_level=sl)
            got = ib.traverse(bvh, start_level=sl)
            assert got.num_contacts == len(want)
            assert got.contacts.numpy().tobytes() == want.tobytes(), (n, sl, "ordered")
            assert (sorted_pairs(want) == brute).all()
            un = ib.traverse(bvh, start_level=sl, ordered=False)
            assert (sorted_pairs(un.contacts.numpy()) == brute).all(), (n, sl, "unordered")
            rs = ib.traverse(bvh, start_level=sl, reference_shaped=True)
            assert rs.contacts.numpy().tobytes() == want.tobytes(), (n, sl, "reference-shaped")
            pk = ib.traverse(bvh, start_level=sl, packet=True)
            assert pk.contacts.numpy().tobytes() == want.tobytes(), (n, sl, "packet")
            wk = ib.traverse(bvh, start_level=sl, walk=True)
            assert wk.contacts.numpy().tobytes() == want.tobytes(), (n, sl, "group walk")
            wu = ib.traverse(bvh, start_level=sl, walk=True, ordered=False)
            assert (sorted_pairs(wu.contacts.numpy()) == brute).all(), (n, sl, "group walk unordered")
            pu = ib.traverse(bvh, start_level=sl, packet=True, ordered=False)
            assert (sorted_pairs(pu.contacts.numpy()) == brute).all(), (n, sl, "packet unordered")


def test_single_counts_cache2_and_growth(ib, O, dev):
    rng = np.random.default_rng(13)
    s = random_spheres(rng, 5000, spread=12.0)
    ol, on = oracle_build(O, s)
    bvh = gpu_build(ib, s)
    want, counts = O.traverse_single(ol, on, want_counts=True)
    tr = ib.traverse(bvh)
    assert tr.contacts.numpy().tobytes() == want.tobytes()
    assert (tr.cache2.numpy()[:5000] == counts).all()          # inclusive scan of per-leaf counts
    # a cache that is too small is grown ("resize only if too small"), a big one is reused as is
    import torch
    small = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(3, ib.pair_dtype(), dev), ib.DeviceArray.empty(10, np.int32, dev))
    tr2 = ib.traverse(bvh, cache=small)
    assert tr2.contacts.numpy().tobytes() == want.tobytes() and len(tr2.cache1) >= len(want)
    big = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(len(want) + 1000, ib.pair_dtype(), dev), ib.DeviceArray.empty(9000, np.int32, dev))
    tr3 = ib.traverse(bvh, cache=big)
    assert tr3.cache1.ptr == big.cache1.ptr and tr3.cache2.ptr == big.cache2.ptr
    assert tr3.contacts.numpy().tobytes() == want.tobytes()
    tr4 = ib.traverse(bvh, cache=small, ordered=False)
    assert (sorted_pairs(tr4.contacts.numpy()) == sorted_pairs(want)).all()
    with pytest.raises(ib.ArgumentError):
        bad = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(8, ib.pair_dtype(np.int64), dev), ib.DeviceArray.empty(8, np.int64, dev))
        ib.traverse(bvh, cache=bad)
    # a `narrow` that accepts everything gives the same list (post-filter over leaf positions)
    assert ib.traverse(bvh, narrow=lambda a, b: np.ones(len(a), bool)).contacts.numpy().tobytes() == want.tobytes()


@pytest.mark.parametrize("ibytes,mbytes", [(8, 8), (4, 2)])
def test_single_other_index_types(ib, O, dev, ibytes, mbytes):
    rng = np.random.default_rng(14)
    for n in (2, 77, 4000):
        s = random_spheres(rng, n, spread=10.0)
        ol, on = oracle_build(O, s, "bbox", ibytes, mbytes)
        bvh = gpu_build(ib, s, "bbox", ibytes, mbytes)
        want = O.traverse_single(ol, on)
        assert ib.traverse(bvh).contacts.numpy().tobytes() == want.tobytes()
        assert (sorted_pairs(ib.traverse(bvh, ordered=False).contacts.numpy()) == sorted_pairs(want)).all()


def test_single_box_leaves(ib, O, dev):
    rng = np.random.default_rng(15)
    for n in (1, 9, 500, 3000):
        b = O.boxes_of_spheres(random_spheres(rng, n, spread=8.0))
        ol, on = oracle_build(O, b)
        bvh = gpu_build(ib, b)
        want = O.traverse_single(ol, on)
        assert ib.traverse(bvh).contacts.numpy().tobytes() == want.tobytes()
        assert (sorted_pairs(want) == sorted_pairs(O.brute_single(b))).all()


def test_single_degenerate_scenes(ib, O, dev):
    # everything touches everything: C = n(n-1)/2
    s = ib.bspheres(np.zeros((300, 3)), np.ones(300))
    bvh = gpu_build(ib, s)
    tr = ib.traverse(bvh)
    assert tr.num_contacts == 300 * 299 // 2
    ol, on = oracle_build(O, s)
    assert tr.contacts.numpy().tobytes() == O.traverse_single(ol, on).tobytes()
    # nothing touches
    s = ib.bspheres(np.arange(600, dtype=np.float32).reshape(200, 3) * 10, np.full(200, 0.1))
    tr = ib.traverse(gpu_build(ib, s))
    assert tr.num_contacts == 0 and len(tr.contacts) == 0
    # exact tangency on a lattice (closed comparisons must agree with the reference: <=, >=)
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    s = ib.bspheres(g, np.full(len(g), 0.5))
    ol, on = oracle_build(O, s)
    want = O.traverse_single(ol, on)
    assert len(want) == 3 * 8 * 8 * 7
    assert ib.traverse(gpu_build(ib, s)).contacts.numpy().tobytes() == want.tobytes()
    # single leaf / two leaves
    assert ib.traverse(gpu_build(ib, ib.bspheres([[0, 0, 0]], [1]))).num_contacts == 0
    assert pairs_list(ib.traverse(gpu_build(ib, ib.bspheres([[0, 0, 0], [1, 0, 0]], [1, 1]))).contacts.numpy()) == [(1, 2)]


def test_single_query_range_shards_concatenate(ib, O, dev):
    """Multi-GPU sharding unit: contiguous query ranges concatenate to the full list."""
    rng = np.random.default_rng(16)
    s = random_spheres(rng, 7001, spread=14.0)
    bvh = gpu_build(ib, s)
    full = ib.traverse(bvh).contacts.numpy()
    parts = []
    bounds = [0, 1000, 1001, 4096, 7001]
    for a, b in zip(bounds[:-1], bounds[1:]):
        parts.append(ib.traverse(bvh, query_range=(a, b - a)).contacts.numpy())
    assert np.concatenate(parts).tobytes() == full.tobytes()


def test_pair_all_start_levels(ib, O, dev):
    """runtests.jl:1009-1081 + gputests.jl:132-208."""
    rng = np.random.default_rng(44)
    sizes = [1, 22, 64, 127, 190]
    for node in ("bbox", "sphere"):
        for n1 in sizes:
            for n2 in sizes:
                s1, s2 = random_spheres(rng, n1), random_spheres(rng, n2)
                o1, on1 = oracle_build(O, s1, node)
                o2, on2 = oracle_build(O, s2, node)
                b1, b2 = gpu_build(ib, s1, node), gpu_build(ib, s2, node)
                brute = sorted_pairs(O.brute_pair(s1, s2))
                for sl1 in {1, b1.tree.levels}:
                    for sl2 in range(1, b2.tree.levels + 1, 2):
                        want = O.traverse_pair(o1, on1, o2, on2, start_level1=sl1, start_level2=sl2)
                        got = ib.traverse(b1, b2, start_level1=sl1, start_level2=sl2)
                        assert got.contacts.numpy().tobytes() == want.tobytes(), (n1, n2, sl1, sl2)
                        assert (sorted_pairs(want) == brute).all()
                        un = ib.traverse(b1, b2, start_level1=sl1, start_level2=sl2, ordered=False)
                        assert (sorted_pairs(un.contacts.numpy()) == brute).all()
                        pk = ib.traverse(b1, b2, start_level1=sl1, start_level2=sl2, packet=True)
                        assert pk.contacts.numpy().tobytes() == want.tobytes(), (n1, n2, sl1, sl2, "packet")
                        wk = ib.traverse(b1, b2, start_level1=sl1, start_level2=sl2, walk=True)
                        assert wk.contacts.numpy().tobytes() == want.tobytes(), (n1, n2, sl1, sl2, "group walk")


def test_pair_self_equivalence_and_partial_build(ib, O, dev):
    rng = np.random.default_rng(17)
    s = random_spheres(rng, 3000, spread=10.0)
    bvh = gpu_build(ib, s)
    single = sorted_pairs(ib.traverse(bvh).contacts.numpy())
    both = sorted_pairs(ib.traverse(bvh, bvh).contacts.numpy())
    sset = {tuple(p) for p in both.tolist()}
    assert all((i, i) in sset for i in range(1, 3001))
    upper = sorted(p for p in sset if p[0] < p[1])
    assert upper == [tuple(p) for p in single.tolist()]
    # config-3 shape: target built only up to a low level, every query scans all roots
    s2 = random_spheres(rng, 2500, spread=10.0)
    levels2 = ib.ImplicitTree(2500).levels
    for bl in (levels2 - 3, levels2 - 7, 1):
        t = gpu_build(ib, s2, built_level=bl)
        o1, on1 = oracle_build(O, s)
        o2, on2 = oracle_build(O, s2, built_level=bl)
        want = O.traverse_pair(o1, on1, o2, on2, built_level2=bl)
        got = ib.traverse(bvh, t)
        assert got.contacts.numpy().tobytes() == want.tobytes()
        assert (sorted_pairs(want) == sorted_pairs(O.brute_pair(s, s2))).all()


def test_rays_small_and_axis_aligned(ib, O, dev):
    rng = np.random.default_rng(45)
    for n in (1, 2, 37, 200, 3000):
        for leaf in ("sphere", "box"):
            s = random_spheres(rng, n, spread=6.0 if n < 1000 else 15.0)
            vols = s if leaf == "sphere" else O.boxes_of_spheres(s)
            ol, on = oracle_build(O, vols)
            bvh = gpu_build(ib, vols)
            R = 500
            p = (8 * rng.random((3, R)) - 1).astype(np.float32)
            d = (rng.random((3, R)) - 0.5).astype(np.float32)
            d[:, :5] = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, 0, -1]], np.float32).T
            d[:, 5] = 0                                   # zero direction: NaNs in the slab test
            for sl in sorted({1, bvh.tree.levels, max(1, bvh.tree.levels // 2)}):
                want = O.traverse_rays(ol, on, p, d, start_level=sl)
                got = ib.traverse_rays(bvh, p, d, start_level=sl)
                assert got.contacts.numpy().tobytes() == want.tobytes(), (n, leaf, sl)
                un = ib.traverse_rays(bvh, p, d, start_level=sl, ordered=False)
                assert (sorted_pairs(un.contacts.numpy()) == sorted_pairs(want)).all()
            # brute force agrees except for the degenerate zero-direction ray (a = b = 0 makes the sphere test
            # vacuously true while the slab test prunes): compare without it
            keep = np.arange(R) != 5
            assert (sorted_pairs(O.traverse_rays(ol, on, p[:, keep], d[:, keep])) == sorted_pairs(O.brute_rays(vols, p[:, keep], d[:, keep]))).all()
    assert ib.traverse_rays(bvh, np.zeros((3, 0), np.float32), np.zeros((3, 0), np.float32)).num_contacts == 0
    with pytest.raises(ib.ArgumentError):
        ib.traverse_rays(bvh, np.zeros((2, 4), np.float32), np.zeros((2, 4), np.float32))
    with pytest.raises(ib.ArgumentError):
        ib.traverse_rays(bvh, np.zeros((3, 4), np.float32), np.zeros((3, 5), np.float32))


def test_ray_grid_against_analytic_sphere(ib, O, dev):
    """runtests.jl:1086-1225 in Float32 — exact order of ray ids equals the oracle's."""
    x, r = np.array([0.9, 1.1, 0.0], np.float32), np.float32(9.9)
    s = ib.bspheres([x], [r])
    bvh = gpu_build(ib, s)
    ol, on = oracle_build(O, s)
    rng_ = [np.arange(x[k] - r, x[k] + r + 1e-6, 1.0) for k in range(3)]
    pts = np.array([[px, py, pz] for pz in rng_[2] for py in rng_[1] for px in rng_[0]], np.float32).T
    for axis in range(3):
        for sign in (1.0, -1.0):
            d = np.zeros_like(pts)
            d[axis, :] = sign
            want = O.traverse_rays(ol, on, pts, d)
            got = ib.traverse_rays(bvh, pts, d)
            assert got.contacts.numpy().tobytes() == want.tobytes()
            assert len(want) > 100


# ---- medium size: oracle in seconds -------------------------------------------------------------------
def test_config1_100k_against_oracle(ib, O, dev):
    """BASELINE config 1 shape (100 k random spheres, BBox nodes, UInt32 / Int32), whole pipeline."""
    from ibvh_b200 import synth
    n = 100_000
    s = synth.random_spheres_np(n, seed=42)
    ol, on = oracle_build(O, s)
    bvh = gpu_build(ib, s)
    assert bvh.leaves.numpy().tobytes() == ol.tobytes()
    assert bvh.nodes.numpy().tobytes() == on.tobytes()
    want = O.traverse_single(ol, on, num_threads=8)
    got = ib.traverse(bvh)
    assert got.contacts.numpy().tobytes() == want.tobytes()
    assert 3.0 * n < len(want) < 5.0 * n                    # C ~ 4 N by construction
    un = ib.traverse(bvh, ordered=False)
    assert (sorted_pairs(un.contacts.numpy()) == sorted_pairs(want)).all()
    for kw in (dict(packet=True), dict(reference_shaped=True), dict(walk=True)):
        assert ib.traverse(bvh, **kw).contacts.numpy().tobytes() == want.tobytes(), kw
        assert (sorted_pairs(ib.traverse(bvh, ordered=False, **kw).contacts.numpy()) == sorted_pairs(want)).all(), kw


def test_rays_mesh_like_against_oracle(ib, O, dev):
    """Config-4 shape at reduced size: 200 x 200 shell, 20 k rays."""
    from ibvh_b200 import synth
    s = synth.shell_spheres_np(200, 200)
    ol, on = oracle_build(O, s)
    bvh = gpu_build(ib, s)
    assert bvh.nodes.numpy().tobytes() == on.tobytes()
    p, d = synth.random_rays_np(20_000, seed=7)
    want = O.traverse_rays(ol, on, p.T, d.T, num_threads=8)
    got = ib.traverse_rays(bvh, p.T, d.T)
    assert got.contacts.numpy().tobytes() == want.tobytes()
    assert len(want) > 20_000
    un = ib.traverse_rays(bvh, p.T, d.T, ordered=False)
    assert (sorted_pairs(un.contacts.numpy()) == sorted_pairs(want)).all()
    for sl in (3, bvh.tree.levels - 2):                      # several roots per ray
        w2 = O.traverse_rays(ol, on, p.T[:, :4000], d.T[:, :4000], start_level=sl, num_threads=8)
        assert ib.traverse_rays(bvh, p.T[:, :4000], d.T[:, :4000], start_level=sl).contacts.numpy().tobytes() == w2.tobytes()
        un2 = ib.traverse_rays(bvh, p.T[:, :4000], d.T[:, :4000], start_level=sl, ordered=False)
        assert (sorted_pairs(un2.contacts.numpy()) == sorted_pairs(w2)).all()


@pytest.mark.parametrize("after", ["2", "40", "0"])
def test_rays_long_ray_queue_thresholds(after):
    """The ray tests again, in a fresh process, with long rays exported to rays_wide_kernel after 2 node steps (nearly every
    ray; the 20 k-ray case overflows the 4096-entry queue, so the queue-full path runs too), after 40, and never (0):
    the same oracle lists come out whichever kernel finishes a ray."""
    import subprocess, sys
    env = dict(os.environ, IBVH_RAYS_WIDE=after)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "rays_small or ray_grid or pair_and_ray_doctests or rays_mesh_like"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "4 passed" in r.stdout, r.stdout[-1000:]


# ---- full BASELINE sizes: size-independent properties ------------------------------------------------
def test_config2_10M_properties(ib, dev):
    """10 M leaves: sortedness, permutation checksum, parent-contains-children, ordered == unordered,
    idempotent rebuild. (The oracle is not run at this size.)"""
    import torch
    from ibvh_b200 import synth
    n = 10_000_000
    vols = synth.random_spheres_torch(n, dev, seed=42)
    src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
    bvh = ib.BVH(src, ib.BBox())
    assert bvh.tree.levels == 25 and len(bvh.nodes) == 10_000_009
    L = bvh.leaves.tensor.view(torch.int32).reshape(n, 6)
    mort = L[:, 5].to(torch.int64) & 0xFFFFFFFF
    assert bool((mort[1:] >= mort[:-1]).all()), "Morton keys ascending"
    idx = L[:, 4].to(torch.int64)
    assert int(idx.sum()) == n * (n + 1) // 2 and int(idx.min()) == 1 and int(idx.max()) == n
    assert int(torch.bincount(idx, minlength=n + 1)[1:].min()) == 1, "indices are a permutation of 1..n"
    # volumes travelled with their index
    sph = bvh.leaves.tensor.view(torch.float32).reshape(n, 6)[:, :4]
    assert bool((sph == vols[(idx - 1)]).all())
    # stable ties: equal keys keep ascending original index
    same = mort[1:] == mort[:-1]
    assert bool((idx[1:][same] > idx[:-1][same]).all())
    # every parent box contains both children (exact min/max merge), bottom-up over all levels
    nodes = bvh.nodes.tensor.view(torch.float32).reshape(-1, 6)
    t = bvh.tree
    skips = t.skips()
    for lvl in range(1, t.levels - 1):
        s0, e0 = ib.level_indices(t, lvl)
        s1, e1 = ib.level_indices(t, lvl + 1)
        par = nodes[s0 - 1:e0]
        ch = nodes[s1 - 1:e1]
        left = ch[0::2][: len(par)]
        assert bool((par[:, :3] <= left[:, :3]).all() and (par[:, 3:] >= left[:, 3:]).all())
        right = ch[1::2]
        pr = par[: len(right)]
        assert bool((pr[:, :3] <= right[:, :3]).all() and (pr[:, 3:] >= right[:, 3:]).all())
        both = torch.minimum(left[: len(right), :3], right[:, :3])
        assert bool((pr[:, :3] == both).all()), "parent lo == min(children lo) exactly"
    # root == bounds of all sphere boxes
    lo = (sph[:, :3] - sph[:, 3:4]).amin(0)
    up = (sph[:, :3] + sph[:, 3:4]).amax(0)
    assert bool((nodes[0, :3] == lo).all() and (nodes[0, 3:] == up).all())
    # traversal: ordered and unordered agree as sets; counts scan ends at the total; pairs are (min, max)
    tr = ib.traverse(bvh)
    C_ = tr.num_contacts
    assert 3.5 * n < C_ < 4.5 * n
    assert int(tr.cache2.tensor.view(torch.int32)[n - 1]) == C_
    pairs = tr.contacts.tensor.view(torch.int32).reshape(-1, 2).to(torch.int64)
    assert bool((pairs[:, 0] < pairs[:, 1]).all())
    key_o = (pairs[:, 0] * (n + 1) + pairs[:, 1]).sort().values
    assert bool((key_o[1:] != key_o[:-1]).all()), "no duplicate pairs"
    un = ib.traverse(bvh, ordered=False, cache=ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(C_ + 16, ib.pair_dtype(), dev), tr.cache2))
    assert un.num_contacts == C_
    pu = un.contacts.tensor.view(torch.int32).reshape(-1, 2).to(torch.int64)
    key_u = (pu[:, 0] * (n + 1) + pu[:, 1]).sort().values
    assert bool((key_o == key_u).all())
    # spot-check 2000 reported contacts and 2000 random non-contacts with the sphere predicate
    orig = vols
    sel = pairs[torch.randint(0, C_, (2000,), device=dev)]
    a, b = orig[sel[:, 0] - 1], orig[sel[:, 1] - 1]
    d2 = ((a[:, :3] - b[:, :3]) ** 2).sum(1)
    assert bool((d2 <= (a[:, 3] + b[:, 3]) ** 2 * (1 + 1e-5)).all())
    # idempotent rebuild on the sorted leaves (cache=bvh)
    before = bvh.nodes.tensor.clone()
    again = ib.BVH(bvh.leaves, ib.BBox(), cache=bvh)
    assert bool((again.nodes.tensor == before).all())
    mort2 = again.leaves.tensor.view(torch.int32).reshape(n, 6)[:, 5]
    assert bool((mort2 == L[:, 5]).all())


def test_config2_10M_against_oracle(ib, O, dev):
    """configs[1] at FULL size against the oracle (test/gputests.jl:34-48,51-127 — GPU structs == CPU structs, sorted GPU
    contacts == sorted CPU contacts): sorted leaves and BBox nodes byte-identical, the ordered contact list byte-identical
    (order included), the unordered list identical as a sorted list."""
    import os
    import torch
    from ibvh_b200 import synth
    n = 10_000_000
    threads = os.cpu_count() or 8
    s = synth.random_spheres_np(n, seed=42)
    ol = O.wrap(s)
    on, _, _ = O.build(ol, O.BBOX, num_threads=threads)
    want = O.traverse_single(ol, on, num_threads=threads)
    bvh = ib.BVH(s, ib.BBox(), device=dev)
    got_leaves = bvh.leaves.tensor.cpu().numpy()
    assert got_leaves.tobytes() == ol.tobytes(), "10 M sorted leaves differ from the oracle"
    got_nodes = bvh.nodes.tensor.cpu().numpy()
    assert got_nodes.tobytes() == on.tobytes(), "10 M BBox nodes differ from the oracle"
    del got_leaves, got_nodes
    tr = ib.traverse(bvh)
    assert tr.num_contacts == len(want)
    got = tr.contacts.tensor.cpu().numpy()
    assert got.tobytes() == want.tobytes(), "10 M ordered contact list differs from the oracle"
    del got
    un = ib.traverse(bvh, ordered=False, cache=ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(len(want) + 16, ib.pair_dtype(), dev), tr.cache2))
    assert un.num_contacts == len(want)
    # sorted lists: (a, b) pairs as one 64-bit key (little endian: a is the low word) -> compare a * 2^32 + b
    pu = un.contacts.tensor.view(torch.int32).reshape(-1, 2).to(torch.int64)
    key_u = ((pu[:, 0] << 32) | pu[:, 1]).sort().values.cpu().numpy()
    key_w = np.sort((want["a"].astype(np.int64) << 32) | want["b"].astype(np.int64))
    assert (key_u == key_w).all(), "10 M unordered contact list differs from the oracle as a sorted list"


def test_config4_rays_1M_shell_against_oracle(ib, O, dev):
    """configs[3] scene at FULL size (1000 x 1000 shell = 1 M leaves) with 2 M rays of the bench's own ray law against
    the oracle: hit list byte-identical in the reference's order (test/gputests.jl:211-248), unordered == sorted."""
    import os
    import torch
    from ibvh_b200 import synth
    threads = os.cpu_count() or 8
    s = synth.shell_spheres_np(1000, 1000)
    ol = O.wrap(s)
    on, _, _ = O.build(ol, O.BBOX, num_threads=threads)
    bvh = ib.BVH(s, ib.BBox(), device=dev)
    assert bvh.leaves.numpy().tobytes() == ol.tobytes() and bvh.nodes.numpy().tobytes() == on.tobytes()
    R = 2_000_000
    p, d = synth.random_rays_np(R, seed=7)
    want = O.traverse_rays(ol, on, p.T, d.T, num_threads=threads)
    got = ib.traverse_rays(bvh, p.T, d.T)
    assert got.num_contacts == len(want) > R
    assert got.contacts.numpy().tobytes() == want.tobytes(), "ray hits differ from the oracle (order included)"
    un = ib.traverse_rays(bvh, p.T, d.T, ordered=False)
    pu = un.contacts.tensor.view(torch.int32).reshape(-1, 2).to(torch.int64)
    key_u = ((pu[:, 0] << 32) | pu[:, 1]).sort().values.cpu().numpy()
    key_w = np.sort((want["a"].astype(np.int64) << 32) | want["b"].astype(np.int64))
    assert (key_u == key_w).all()


# ---- BASELINE configs[2] and configs[4] at full size: size-independent properties ------------------------
def _keys(t, n_mult):
    """Sorted 64-bit keys of an IndexPair tensor view (int32 or int64 pairs)."""
    return (t[:, 0].to(__import__("torch").int64) * n_mult + t[:, 1].to(__import__("torch").int64)).sort().values


def test_config3_pair_5M_partial_build_properties(ib, dev):
    """configs[2]: BVH-vs-BVH, 5 M + 5 M leaves, target built only up to levels-13 (about 2^10 roots):
    the partial-build traversal (group walk from every root, as the reference scans them) must give the same
    contact set as the fully built tree (pyramid schedule); flip symmetry; spot-check with the sphere predicate."""
    import torch
    from ibvh_b200 import synth
    n = 5_000_000
    s = synth.sphere_radius_scale(n)
    v1 = synth.random_spheres_torch(n, dev, seed=42, scale=s)
    v2 = synth.random_spheres_torch(n, dev, seed=43, scale=s)
    d1 = ib.DeviceArray(v1.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
    d2 = ib.DeviceArray(v2.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
    b1 = ib.BVH(d1, ib.BBox())
    levels = b1.tree.levels
    b2_full = ib.BVH(d2, ib.BBox())
    b2_part = ib.BVH(d2, ib.BBox(), built_level=levels - 13)
    assert b2_part.built_level == levels - 13
    full = ib.traverse(b1, b2_full, ordered=False)
    part = ib.traverse(b1, b2_part, ordered=False)                     # start_level2 defaults to built_level
    assert part.start_level2 == levels - 13
    assert full.num_contacts == part.num_contacts > n
    kf = _keys(full.contacts.tensor.view(torch.int32).reshape(-1, 2), n + 1)
    kp = _keys(part.contacts.tensor.view(torch.int32).reshape(-1, 2), n + 1)
    assert bool((kf == kp).all())
    assert bool((kf[1:] != kf[:-1]).all()), "no duplicates"
    # ordered mode gives the same set again, and its scan ends at the total
    o = ib.traverse(b1, b2_full)
    assert o.num_contacts == full.num_contacts
    assert int(o.cache2.tensor.view(torch.int32)[n - 1]) == o.num_contacts
    assert bool((_keys(o.contacts.tensor.view(torch.int32).reshape(-1, 2), n + 1) == kf).all())
    # swapped arguments: same pairs with the roles exchanged (both have n leaves, so no flip; exchange manually)
    sw = ib.traverse(b2_full, b1, ordered=False)
    ps = sw.contacts.tensor.view(torch.int32).reshape(-1, 2)
    ks = (ps[:, 1].to(torch.int64) * (n + 1) + ps[:, 0].to(torch.int64)).sort().values
    assert bool((ks == kf).all())
    # every reported pair satisfies the sphere predicate (2000 samples) and so do none of 2000 random pairs' complement
    pairs = full.contacts.tensor.view(torch.int32).reshape(-1, 2).to(torch.int64)
    sel = pairs[torch.randint(0, full.num_contacts, (2000,), device=dev)]
    a, b = v1[sel[:, 0] - 1], v2[sel[:, 1] - 1]
    d2_ = ((a[:, :3] - b[:, :3]) ** 2).sum(1)
    assert bool((d2_ <= (a[:, 3] + b[:, 3]) ** 2 * (1 + 1e-5)).all())


def test_config5_100M_uint64_int64_cached_rebuild(ib, dev):
    """configs[4]: 100 M leaves, UInt64 Morton, Int64 index: build, perturb, cached rebuild in place
    (BVH(bvh.leaves, cache=bvh)), contact traversal with reused caches. Properties only."""
    import torch
    from ibvh_b200 import synth
    n = 100_000_000
    o = opts(ib, 8, 8)
    vols = synth.random_spheres_torch(n, dev, seed=42)
    src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
    bvh = ib.BVH(src, ib.BBox(), options=o)
    del src
    assert bvh.tree.levels == 28 and len(bvh.nodes) == 100_000_007 and bvh.leaves.dtype.itemsize == 32
    L = bvh.leaves.tensor.view(torch.int64).reshape(n, 4)
    mort = L[:, 3]
    assert bool((mort[1:] >= mort[:-1]).all()) and int(mort.min()) >= 0, "63-bit keys ascending"
    idx = L[:, 2]
    assert int(idx.sum()) == n * (n + 1) // 2 and int(idx.min()) == 1 and int(idx.max()) == n
    tr = ib.traverse(bvh, ordered=False)
    C_ = tr.num_contacts
    assert 3.5 * n < C_ < 4.5 * n
    pairs = tr.contacts.tensor.view(torch.int64).reshape(-1, 2)
    assert bool((pairs[:, 0] < pairs[:, 1]).all())
    sel = pairs[torch.randint(0, C_, (2000,), device=dev)]
    a, b = vols[sel[:, 0] - 1], vols[sel[:, 1] - 1]
    d2_ = ((a[:, :3] - b[:, :3]) ** 2).sum(1)
    assert bool((d2_ <= (a[:, 3] + b[:, 3]) ** 2 * (1 + 1e-5)).all())
    del pairs, sel
    # one simulation step: perturb the centres in place (leaf order = previous Morton order), rebuild with cache
    s = synth.sphere_radius_scale(n)
    F = bvh.leaves.tensor.view(torch.float32).reshape(n, 8)
    for k in range(3):
        F[:, k] += (synth.uniform_torch(1234, n, k, dev) - 0.5) * (s / 2)
    moved = F[:, :4].clone()
    idx_before = L[:, 2].clone()
    nodes_ptr, leaves_ptr = bvh.nodes.ptr, bvh.leaves.ptr
    again = ib.BVH(bvh.leaves, ib.BBox(), cache=bvh, options=o)
    assert again.nodes.ptr == nodes_ptr and again.leaves.ptr == leaves_ptr, "in place, node buffer reused"
    L2 = again.leaves.tensor.view(torch.int64).reshape(n, 4)
    assert bool((L2[1:, 3] >= L2[:-1, 3]).all())
    assert int(L2[:, 2].sum()) == n * (n + 1) // 2
    # each sphere travelled with its index: compare through the index -> previous position map
    pos_of_index = torch.empty(n + 1, dtype=torch.int64, device=dev)
    pos_of_index[idx_before] = torch.arange(n, device=dev)
    F2 = again.leaves.tensor.view(torch.float32).reshape(n, 8)[:, :4]
    assert bool((F2 == moved[pos_of_index[L2[:, 2]]]).all())
    del moved, pos_of_index, idx_before
    tr2 = ib.traverse(again, ordered=False, cache=tr)
    assert tr2.cache1.ptr == tr.cache1.ptr or tr2.num_contacts > len(tr.cache1)
    assert 3.0 * n < tr2.num_contacts < 5.0 * n


# ---- Float64 volumes (the reference's own tests are mostly Float64; runtests.jl:596-900) -------------------
@pytest.mark.parametrize("node", ["bbox", "sphere"])
def test_float64_build_and_traversals(ib, O, dev, node):
    rng = np.random.default_rng(64)
    nt = ib.BBox(np.float64) if node == "bbox" else ib.BSphere(np.float64)
    for n in (1, 2, 5, 37, 200, 3000, 20_000):
        s = random_spheres(rng, n, fbytes=8, spread=6.0 * max(1.0, (n / 200.0) ** (1 / 3)))
        for ibytes, mbytes in ((4, 4), (8, 8), (4, 2)):
            ol, on = oracle_build(O, s, node, ibytes, mbytes)
            bvh = ib.BVH(s, nt, options=opts(ib, ibytes, mbytes))
            assert bvh.leaves.numpy().tobytes() == ol.tobytes(), (n, ibytes, mbytes, "sorted leaves")
            assert_nodes_equal(bvh.nodes.numpy(), on, node)
        want = O.traverse_single(ol, on)
        assert ib.traverse(bvh).contacts.numpy().tobytes() == want.tobytes(), (n, node)
        assert (sorted_pairs(ib.traverse(bvh, ordered=False).contacts.numpy()) == sorted_pairs(want)).all()
        assert (sorted_pairs(want) == sorted_pairs(O.brute_single(s))).all()
        for sl in range(1, bvh.tree.levels + 1, 3):
            w2 = O.traverse_single(ol, on, start_level=sl)
            assert ib.traverse(bvh, start_level=sl).contacts.numpy().tobytes() == w2.tobytes(), (n, node, sl)
    # pair + rays in Float64
    s1, s2 = random_spheres(rng, 700, 8), random_spheres(rng, 450, 8)
    o1, on1 = oracle_build(O, s1, node)
    o2, on2 = oracle_build(O, s2, node)
    b1, b2 = ib.BVH(s1, nt), ib.BVH(s2, nt)
    want = O.traverse_pair(o1, on1, o2, on2)
    assert ib.traverse(b1, b2).contacts.numpy().tobytes() == want.tobytes()
    assert (sorted_pairs(want) == sorted_pairs(O.brute_pair(s1, s2))).all()
    if node == "bbox":
        p = (8 * rng.random((3, 400)) - 1)
        d = (rng.random((3, 400)) - 0.5)
        wr = O.traverse_rays(o1, on1, p, d)
        gr = ib.traverse_rays(b1, p, d)
        assert gr.contacts.numpy().tobytes() == wr.tobytes()
        assert (sorted_pairs(ib.traverse_rays(b1, p, d, ordered=False).contacts.numpy()) == sorted_pairs(wr)).all()


def test_float64_reference_doctest(ib, golden, dev):
    """README.md:37-54 uses Float64 spheres; with Float64 nodes the contacts are the documented ones."""
    g = golden["five_spheres"]
    bvh = ib.BVH(ib.bspheres(g["centers"], g["radii"], np.float64), ib.BBox(np.float64))
    assert pairs_list(ib.traverse(bvh).contacts.numpy()) == [tuple(p) for p in g["contacts_lvt_order"]]
    # the reference's DEFAULT call: Float64 spheres, node type left at BBox{Float32} (README.md:38-46, build.jl:198-205)
    s64 = ib.bspheres(g["centers"], g["radii"], np.float64)
    mixed = ib.BVH(s64)
    assert mixed.nodes.dtype == ib.BBox(np.float32).dtype and mixed.leaves.dtype["volume"] == ib.BSphere(np.float64).dtype
    assert pairs_list(ib.traverse(mixed).contacts.numpy()) == [tuple(p) for p in g["contacts_lvt_order"]]
    with pytest.raises(ib.ArgumentError):
        ib.traverse_rays(mixed, np.zeros((3, 1)), np.ones((3, 1)))                       # no mixed-type ray test in the reference either
    with pytest.raises(NotImplementedError):
        ib.BVH(ib.bspheres(g["centers"], g["radii"], np.float32), ib.BBox(np.float64))   # Float64 nodes over Float32 leaves: not built


def test_mixed_float64_leaves_float32_nodes(ib, O, dev):
    """Float64 leaves under Float32 nodes (the reference's default node type): sorted leaves and nodes byte-identical to the
    oracle's converting merges (merge.jl:47-81: arithmetic in the leaf type, result converted), contacts in order, every
    start level, both node kinds, pair traversal; and the sets equal brute force."""
    rng = np.random.default_rng(6432)
    for node in ("bbox", "sphere"):
        nt = ib.BBox(np.float32) if node == "bbox" else ib.BSphere(np.float32)
        nk = O.BBOX if node == "bbox" else O.BSPHERE
        for n in (1, 2, 3, 5, 64, 201, 3000, 40_000):
            s = random_spheres(rng, n, fbytes=8, spread=6.0 * max(1.0, (n / 200.0) ** (1 / 3)))
            for ibytes, mbytes in ((4, 4), (8, 8)):
                ol = O.wrap(s, ibytes, mbytes)
                on, _, _ = O.build(ol, nk, node_fbytes=4)
                bvh = ib.BVH(s, nt, options=opts(ib, ibytes, mbytes), device=dev)
                assert bvh.leaves.numpy().tobytes() == ol.tobytes(), (node, n, "leaves")
                assert_nodes_equal(bvh.nodes.numpy(), on, node)
            want = O.traverse_single(ol, on)
            assert ib.traverse(bvh).contacts.numpy().tobytes() == want.tobytes(), (node, n)
            assert (sorted_pairs(ib.traverse(bvh, ordered=False).contacts.numpy()) == sorted_pairs(want)).all()
            # (no brute-force check here: Float32 node boxes are ROUNDED conversions of Float64 volumes and need not contain
            # them, so the reference's own contact set may differ from the sphere predicate by tangency-level pairs)
            for sl in range(1, bvh.tree.levels + 1, 4):
                w2 = O.traverse_single(ol, on, start_level=sl)
                assert ib.traverse(bvh, start_level=sl).contacts.numpy().tobytes() == w2.tobytes(), (node, n, sl)
        s1, s2 = random_spheres(rng, 900, 8), random_spheres(rng, 500, 8)
        o1 = O.wrap(s1); o2 = O.wrap(s2)
        on1, _, _ = O.build(o1, nk, node_fbytes=4)
        on2, _, _ = O.build(o2, nk, node_fbytes=4)
        b1, b2 = ib.BVH(s1, nt, device=dev), ib.BVH(s2, nt, device=dev)
        assert ib.traverse(b1, b2).contacts.numpy().tobytes() == O.traverse_pair(o1, on1, o2, on2).tobytes(), node


def test_triangles_to_volumes_and_mesh_pipeline(ib, O, golden, dev):
    """leaf volumes from triangles on the device, bit-exact with the oracle, then the whole
    mesh -> volumes -> BVH -> contacts pipeline (benchmark/bvh_contact.jl shape) without leaving the device."""
    rng = np.random.default_rng(5)
    for T, fb in ((np.float32, 4), (np.float64, 8)):
        tris = (rng.random((4000, 1, 3)) * 10 + rng.random((4000, 3, 3)) * 0.4).astype(T)
        tris[:50, 2] = tris[:50, 0] + T(0.5) * (tris[:50, 1] - tris[:50, 0])      # collinear
        tris[50:60] = tris[50:60, :1]                                              # all three vertices equal
        for kind, vt in ((O.BSPHERE, ib.BSphere(T)), (O.BBOX, ib.BBox(T))):
            want = O.volumes_from_triangles(tris, kind, fb)
            got = ib.volumes_from_triangles(tris, vt, device=dev)
            assert got.numpy().tobytes() == want.tobytes(), (T, kind)
        vols = ib.volumes_from_triangles(tris, ib.BSphere(T), device=dev)
        bvh = ib.BVH(vols, ib.BBox(T))
        ol, on = oracle_build(O, O.volumes_from_triangles(tris, O.BSPHERE, fb))
        assert bvh.leaves.numpy().tobytes() == ol.tobytes() and bvh.nodes.numpy().tobytes() == on.tobytes()
        assert ib.traverse(bvh).contacts.numpy().tobytes() == O.traverse_single(ol, on).tobytes()
    g = golden["triangle_volumes"]
    for c in g["bsphere"]:
        s = ib.volumes_from_triangles(np.array([c["tri"]], np.float64), ib.BSphere(np.float64), device=dev).numpy()[0]
        assert np.allclose(s["x"], c["x"], rtol=1e-12, atol=1e-15) and np.isclose(s["r"], c["r"], rtol=1e-12)
    with pytest.raises(ib.ArgumentError):
        ib.volumes_from_triangles(np.zeros((4, 3, 2), np.float32), device=dev)


@pytest.mark.gpu
def test_multi_gpu_fused_traversal_and_peer_gather():
    """Multi-GPU exchange through the C ABI: ibvh_allgather_pairs reproduces the single-GPU ordered list
    from rank-ordered shards, and the fused traversal (ibvh_traverse_params_t.peer) the single-GPU contact set.
    One process per GPU under torchrun; with one visible GPU the same code runs as a world of 1."""
    import os
    import subprocess
    import sys
    import torch
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ngpu = torch.cuda.device_count()
    world = 2 if ngpu >= 2 else 1
    env = dict(os.environ, LEAVES="300000", TIMING="0", NCCL_DEBUG="WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tools", "fused_test.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PARITY ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.gpu
def test_deferred_traversal_matches_synchronous(ib, O, dev):
    """IBVH_TRAVERSE_DEFER + ibvh_traverse_finish: same contact set as the synchronous call; the next build may be
    enqueued before the count is read; a second traversal on the handle finishes the outstanding one first; a too small
    contacts buffer is sorted out by the fall-back; a dropped result cancels itself."""
    import torch
    from ibvh_b200 import synth
    n = 200_000
    vols = synth.random_spheres_np(n, seed=11)
    bvh = ib.BVH(vols, ib.BBox(), device=dev)
    ref = ib.traverse(bvh, ordered=False)
    want = torch.sort(ref.contacts.tensor.view(torch.int64)).values
    cache = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(ref.num_contacts + 64, ib.pair_dtype(), dev), ref.cache2)
    for rep in range(3):
        tr = ib.traverse(bvh, cache=cache, ordered=False, defer=True)
        assert tr._resolve is not None, "the call must have been deferred"
        bvh2 = ib.BVH(vols, ib.BBox(), device=dev, cache=bvh)          # enqueued while the traversal is outstanding
        assert bvh2.nodes.ptr != bvh.nodes.ptr, "a BVH with a pending traversal keeps its node buffer to itself"
        assert tr.num_contacts == ref.num_contacts
        got = torch.sort(tr.contacts.tensor.view(torch.int64)).values
        assert torch.equal(got, want)
        cache = tr
    # contacts buffer too small: finish reports the need, the fall-back call regrows cache1
    small = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(1000, ib.pair_dtype(), dev), ref.cache2)
    tr = ib.traverse(bvh, cache=small, ordered=False, defer=True)
    assert tr.num_contacts == ref.num_contacts and len(tr.cache1) >= ref.num_contacts
    assert torch.equal(torch.sort(tr.contacts.tensor.view(torch.int64)).values, want)
    # scratch pair lists sized from a sparse scene, then a dense scene deferred: finish answers IBVH_ERR_AGAIN and
    # the fall-back repeats the traversal
    dense = ib.BVH(synth.random_spheres_np(n, seed=13, scale=2.5 * synth.sphere_radius_scale(n)), ib.BBox(), device=dev)
    refd = ib.traverse(dense, ordered=False)
    sparse = ib.BVH(synth.random_spheres_np(n, seed=14, scale=0.05 * synth.sphere_radius_scale(n)), ib.BBox(), device=dev)
    ib.traverse(sparse, ordered=False)
    cd = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(refd.num_contacts + 64, ib.pair_dtype(), dev), refd.cache2)
    trd = ib.traverse(dense, cache=cd, ordered=False, defer=True)
    assert trd.num_contacts == refd.num_contacts
    assert torch.equal(torch.sort(trd.contacts.tensor.view(torch.int64)).values, torch.sort(refd.contacts.tensor.view(torch.int64)).values)
    # pair traversal
    vols_b = synth.random_spheres_np(n // 2, seed=12, scale=synth.sphere_radius_scale(n))
    bvh_b = ib.BVH(vols_b, ib.BBox(), device=dev)
    refp = ib.traverse(bvh, bvh_b, ordered=False)
    cp = ib.BVHTraversal(1, 1, 0, 0, ib.DeviceArray.empty(refp.num_contacts + 64, ib.pair_dtype(), dev), refp.cache2)
    trp = ib.traverse(bvh, bvh_b, cache=cp, ordered=False, defer=True)
    assert trp.num_contacts == refp.num_contacts
    assert torch.equal(torch.sort(trp.contacts.tensor.view(torch.int64)).values, torch.sort(refp.contacts.tensor.view(torch.int64)).values)


@pytest.mark.gpu
def test_deferred_traversal_survives_rebuild_with_other_geometry(ib, O, dev):
    """between a deferred traversal and the read of its count, a build with DIFFERENT geometry and
    cache=bvh is enqueued. If the deferred call then has to be repeated (IBVH_ERR_AGAIN: scratch lists too small;
    IBVH_ERR_CAPACITY: cache1 too small) the repeat must still see step k's tree, not step k+1's nodes."""
    import torch
    from ibvh_b200 import synth
    n = 200_000
    dense_vols = synth.random_spheres_np(n, seed=13, scale=2.5 * synth.sphere_radius_scale(n))
    other_vols = synth.random_spheres_np(n, seed=99)
    dense = ib.BVH(dense_vols, ib.BBox(), device=dev)
    refd = ib.traverse(dense, ordered=False)
    want = torch.sort(refd.contacts.tensor.view(torch.int64)).values.clone()
    n_want = refd.num_contacts
    for mode in ("again", "capacity"):
        dense = ib.BVH(dense_vols, ib.BBox(), device=dev)
        if mode == "again":
            # learn small scratch lists from a sparse scene, then defer the dense one with a large enough cache1
            sparse = ib.BVH(synth.random_spheres_np(n, seed=14, scale=0.05 * synth.sphere_radius_scale(n)), ib.BBox(), device=dev)
            ib.traverse(sparse, ordered=False)
            cache = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(n_want + 64, ib.pair_dtype(), dev), refd.cache2)
        else:
            ib.traverse(dense, ordered=False)                              # scratch lists sized right, cache1 far too small
            cache = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(1000, ib.pair_dtype(), dev), refd.cache2)
        tr = ib.traverse(dense, cache=cache, ordered=False, defer=True)
        assert tr._resolve is not None
        nodes_before = dense.nodes.tensor.clone()
        nxt = ib.BVH(other_vols, ib.BBox(), device=dev, cache=dense)       # different geometry, wants to reuse dense.nodes
        torch.cuda.synchronize()
        assert torch.equal(dense.nodes.tensor, nodes_before), "the pending BVH's nodes were overwritten by the next build"
        assert tr.num_contacts == n_want, mode
        assert torch.equal(torch.sort(tr.contacts.tensor.view(torch.int64)).values, want), mode
        # the in-place form: rebuilding over the pending BVH's own leaves resolves the traversal first
        tr2 = ib.traverse(dense, cache=tr, ordered=False, defer=True)
        if tr2._resolve is not None:
            again = ib.BVH(dense.leaves, ib.BBox(), device=dev, cache=dense)
            assert tr2._resolve is None, "an in-place rebuild must finish the traversal that still reads those leaves"
        assert tr2.num_contacts == n_want
        del nxt


@pytest.mark.gpu
def test_deferred_traversal_cancel_and_rays_guard(ib, O, dev):
    """a deferred result that is dropped unread must not block the handle; traverse_rays must not
    run over an outstanding deferred traversal's read-back slots."""
    import gc
    import torch
    from ibvh_b200 import synth
    n = 100_000
    vols = synth.random_spheres_np(n, seed=21)
    bvh = ib.BVH(vols, ib.BBox(), device=dev)
    ref = ib.traverse(bvh, ordered=False)
    mk = lambda: ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(ref.num_contacts + 64, ib.pair_dtype(), dev), ref.cache2)
    tr = ib.traverse(bvh, cache=mk(), ordered=False, defer=True)
    assert tr._resolve is not None
    del tr
    gc.collect()                                                           # dropped unread: cancels itself
    t2 = ib.traverse(bvh, cache=mk(), ordered=False)
    assert t2.num_contacts == ref.num_contacts
    # rays while a deferred traversal is outstanding: the host layer finishes the traversal first ...
    tr = ib.traverse(bvh, cache=mk(), ordered=False, defer=True)
    p, d = synth.random_rays_np(2000, seed=3)
    pts = (p * 0.4 + 0.5).T
    rays = ib.traverse_rays(bvh, pts, d.T)
    assert tr._resolve is None and tr.num_contacts == ref.num_contacts
    ol = O.wrap(vols)
    on, _, _ = O.build(ol, O.BBOX)
    assert rays.contacts.numpy().tobytes() == O.traverse_rays(ol, on, pts, d.T).tobytes()
    # ... and the C entry point itself refuses (status ArgumentError) when called underneath the host layer
    import ctypes as C
    tr = ib.traverse(bvh, cache=mk(), ordered=False, defer=True)
    assert tr._resolve is not None
    lib = ib.capi.lib()
    cb = bvh._c_bvh()
    params = ib.capi.TraverseParams(1, 0, -1, ib.capi.TRAVERSE_ORDERED, 0, 0, None)
    pt = torch.from_numpy(np.ascontiguousarray(pts.T.astype(np.float32))).to(dev)
    dt = torch.from_numpy(np.ascontiguousarray(d.astype(np.float32))).to(dev)
    total = C.c_int64(0)
    rc = lib.ibvh_traverse_rays(bvh._handle, C.byref(cb), pt.data_ptr(), dt.data_ptr(), 2000, C.byref(params), None, None, 0, C.byref(total),
                                torch.cuda.current_stream().cuda_stream)
    assert rc == ib.capi.ERR_ARGUMENT
    assert tr.num_contacts == ref.num_contacts
    assert lib.ibvh_traverse_cancel(bvh._handle) == ib.capi.OK              # nothing outstanding: no-op


@pytest.mark.gpu
def test_reference_shaped_proxy_build_is_bit_identical(ib, O, dev):
    """The reference-shaped proxy (struct-moving merge sort, two mapreduce passes, one merge launch per level; the stand-in
    for the reference's CUDA.jl backend in bench.py) must give the oracle's leaves and nodes bit for bit, ties included."""
    from ibvh_b200 import synth
    rng = np.random.default_rng(8)
    for n in (1, 2, 3, 511, 512, 513, 1025, 5000, 100_003):
        s = random_spheres(rng, n, spread=6.0 * max(1.0, (n / 200.0) ** (1 / 3)))
        ol, on = oracle_build(O, s)
        bvh = ib.BVH(s, ib.BBox(), device=dev, reference_shaped=True)
        assert bvh.leaves.numpy().tobytes() == ol.tobytes(), n
        assert bvh.nodes.numpy().tobytes() == on.tobytes(), n
    # many ties (few distinct Morton codes): the merge sort must be stable like the product's radix sort
    s = synth.random_spheres_np(60_000, seed=5)
    s["x"] = np.round(s["x"] * 4) / 4
    ol, on = oracle_build(O, s)
    bvh = ib.BVH(s, ib.BBox(), device=dev, reference_shaped=True)
    assert bvh.leaves.numpy().tobytes() == ol.tobytes() and bvh.nodes.numpy().tobytes() == on.tobytes()
    want = O.traverse_single(ol, on, num_threads=4)
    assert ib.traverse(bvh, reference_shaped=True).contacts.numpy().tobytes() == want.tobytes()
    with pytest.raises(NotImplementedError):
        ib.BVH(s, ib.BBox(), device=dev, reference_shaped=True, options=opts(ib, 8, 8))


@pytest.mark.gpu
def test_ordered_long_segments_and_narrow(ib, O, dev):
    """Dense clusters: query leaves with tens, hundreds and > 10^4 contacts (one sphere covering the whole scene) — the
    ordered fix-up sorts their segments with the warp / block bitonic tiers instead of round 1's O(m^2) insertion sort.
    Then `narrow` as a post-filter over leaf positions (IBVH_TRAVERSE_POSITIONS) against the same predicate applied to the
    oracle's list, and BVHTraversal.num_checks."""
    rng = np.random.default_rng(77)
    n = 40_000
    s = random_spheres(rng, n, spread=40.0)
    s["r"] *= 0.5
    s["x"][0] = 20.0; s["r"][0] = 80.0                       # touches every other leaf
    s["x"][1:400] = 5.0 + rng.random((399, 3)).astype(np.float32) * 0.5      # a cluster: ~400 mutual contacts each
    s["x"][400:460] = 30.0 + rng.random((60, 3)).astype(np.float32) * 0.2    # a smaller one: ~60 each
    ol, on = oracle_build(O, s)
    want = O.traverse_single(ol, on, num_threads=8)
    bvh = ib.BVH(s, ib.BBox(), device=dev)
    got = ib.traverse(bvh)
    assert got.num_contacts == len(want) > n
    assert got.contacts.numpy().tobytes() == want.tobytes()
    assert got.num_checks > got.num_contacts                  # box + leaf tests of the pyramid schedule
    # with a pre-sized cache (the one-pass stash protocol) and through the packet schedule
    again = ib.traverse(bvh, cache=got)
    assert again.contacts.numpy().tobytes() == want.tobytes()
    assert ib.traverse(bvh, packet=True).contacts.numpy().tobytes() == want.tobytes()
    # pair traversal with long segments
    s2 = random_spheres(rng, 5000, spread=40.0)
    o2, on2 = oracle_build(O, s2)
    wp = O.traverse_pair(ol, on, o2, on2, num_threads=8)
    b2 = ib.BVH(s2, ib.BBox(), device=dev)
    assert ib.traverse(bvh, b2).contacts.numpy().tobytes() == wp.tobytes()
    # narrow: keep pairs whose index sum is even and whose first volume is the smaller one
    pred = lambda a, b: ((a["index"] + b["index"]) % 2 == 0) & (a["volume"]["r"] <= b["volume"]["r"])
    pos_of = {int(ix): k for k, ix in enumerate(ol["index"])}
    def filt(pairs, la, lb, single):
        keep = []
        for a, b in zip(pairs["a"], pairs["b"]):
            pa, pb = pos_of_a[int(a)], pos_of_b[int(b)]
            if single and pa > pb:
                pa, pb = pb, pa                              # the query is the leaf at the lower position
            keep.append(bool(pred(la[pa:pa + 1], lb[pb:pb + 1])[0]))
        return pairs[np.array(keep, bool)]
    pos_of_a = pos_of_b = pos_of
    want_n = filt(want, ol, ol, True)
    got_n = ib.traverse(bvh, narrow=pred)
    assert got_n.contacts.numpy().tobytes() == want_n.tobytes()
    assert int(got_n.cache2.numpy()[-1]) == len(want_n)
    scalar_pred = lambda a, b: bool((int(a["index"]) + int(b["index"])) % 2 == 0 and a["volume"]["r"] <= b["volume"]["r"])
    assert ib.traverse(bvh, narrow=scalar_pred, ordered=True).contacts.numpy().tobytes() == want_n.tobytes()     # element-wise fallback
    pos_of_b = {int(ix): k for k, ix in enumerate(o2["index"])}
    want_pn = filt(wp, ol, o2, False)
    assert ib.traverse(bvh, b2, narrow=pred).contacts.numpy().tobytes() == want_pn.tobytes()
    # rays: narrow(leaf, point, direction)
    p = (rng.random((3, 3000)) * 40).astype(np.float32)
    d = (rng.random((3, 3000)) - 0.5).astype(np.float32)
    wr = O.traverse_rays(ol, on, p, d, num_threads=8)
    rpred = lambda leaf, pt, dr: (leaf["index"] % 3 == 0) & (pt[:, 0] > 10.0)
    keep = (wr["a"] % 3 == 0) & (p[0][wr["b"] - 1] > 10.0)
    gr = ib.traverse_rays(bvh, p, d, narrow=rpred)
    assert gr.contacts.numpy().tobytes() == wr[keep].tobytes()


@pytest.mark.gpu
def test_sidecar_reuse_eviction_and_fallback(ib, O, dev):
    """The build's sidecar (ibvh_bvh_t.build_id: packed records, aligned node levels, pyramid levels, index array) is an
    accelerator only: a traversal gives the same bytes whether its BVH's sidecar is still held (2 slots per handle), was
    evicted by later builds, or never existed; an in-place rebuild gets a fresh one; pair traversals use both trees'."""
    from ibvh_b200 import synth
    n = 150_000
    sets = [synth.random_spheres_np(n, seed=s) for s in (31, 32, 33, 34)]
    want = []
    for s in sets:
        ol, on = oracle_build(O, s)
        want.append((ol, on, O.traverse_single(ol, on, num_threads=8)))
    bvhs = [ib.BVH(s, ib.BBox(), device=dev) for s in sets]              # four builds: the first two sidecars are evicted
    ids = [b._build_id for b in bvhs]
    assert all(i > 0 for i in ids) and len(set(ids)) == 4
    for k in (0, 3, 1, 2):                                               # evicted (pack on the fly) and held sidecars alike
        assert ib.traverse(bvhs[k]).contacts.numpy().tobytes() == want[k][2].tobytes(), k
        assert (sorted_pairs(ib.traverse(bvhs[k], ordered=False).contacts.numpy()) == sorted_pairs(want[k][2])).all(), k
    # a stale / foreign id must not be trusted: same n, same id value, other arrays
    fake = ib.BVH(sets[0], ib.BBox(), device=dev)
    fake._build_id = bvhs[3]._build_id
    assert ib.traverse(fake, ordered=False).num_contacts >= 0            # (contract broken by the caller: must not crash or hang)
    # pair traversal: queries' and target's sidecars, held and evicted
    wp = O.traverse_pair(want[2][0], want[2][1], want[3][0], want[3][1], num_threads=8)
    assert ib.traverse(bvhs[2], bvhs[3]).contacts.numpy().tobytes() == wp.tobytes()
    wp2 = O.traverse_pair(want[0][0], want[0][1], want[3][0], want[3][1], num_threads=8)
    assert ib.traverse(bvhs[0], bvhs[3]).contacts.numpy().tobytes() == wp2.tobytes()
    # in-place rebuild (cache=bvh): new id, same result as before (idempotent on sorted leaves)
    again = ib.BVH(bvhs[3].leaves, ib.BBox(), cache=bvhs[3])
    assert again._build_id not in ids
    assert ib.traverse(again).contacts.numpy().tobytes() == want[3][2].tobytes()
    # sharded query ranges with the full-range pyramid of the sidecar (boundary groups carry boxes of out-of-range leaves)
    parts = []
    for qb, qc in ((0, 50_001), (50_001, 49_999), (100_000, 50_000)):
        parts.append(ib.traverse(again, query_range=(qb, qc)).contacts.numpy())
    assert np.concatenate(parts).tobytes() == want[3][2].tobytes()


# ---- BFSTraversal (src/traverse/breadth_first, src/raytrace/breadth_first) ------------------------------------------------
# The reference's GPU backend appends with atomics, so its BFS lists are sets: contacts are compared as sorted lists
# (test/gputests.jl:73-78 sorts too); num_checks (the summed BVTT lengths) is deterministic and compared exactly.
@pytest.mark.gpu
def test_bfs_doctests_and_old_interface(ib, golden, dev):
    g = golden["five_spheres"]
    want = sorted(tuple(p) for p in g["contacts_lvt_order"])
    for node in ("bbox", "sphere"):
        for ibytes, mbytes in ((4, 4), (8, 8), (4, 2)):
            bvh = gpu_build(ib, ib.bspheres(g["centers"], g["radii"]), node, ibytes, mbytes)
            tr = ib.traverse(bvh, ib.BFSTraversal())
            assert sorted(pairs_list(tr.contacts.numpy())) == want
            assert tr.start_level1 == ib.default_start_level(bvh, ib.BFSTraversal()) == 2      # breadth_first.jl:4-6
            assert tr.contacts.numpy().dtype == ib.pair_dtype(I_of(ibytes)) and tr.cache2.dtype == ib.pair_dtype(I_of(ibytes))
            tr2 = ib.traverse(bvh, ib.BFSTraversal(), cache=tr)
            assert tr2.cache1.ptr == tr.cache1.ptr and sorted(pairs_list(tr2.contacts.numpy())) == want
            old = ib.traverse(bvh, 3)                                                         # traverse(bvh, start_level) = BFS, traverse.jl:233-241
            assert old.start_level1 == 3 and sorted(pairs_list(old.contacts.numpy())) == want
    bvh = gpu_build(ib, ib.bspheres(g["centers"], g["radii"]))
    with pytest.raises(ib.ArgumentError):
        ib.traverse(bvh, ib.BFSTraversal(), cache=ib.traverse(bvh))          # an LVT cache: eltype(cache.cache2) === IndexPair{I} fails
    with pytest.raises(ib.ArgumentError):
        ib.traverse(bvh, ib.BFSTraversal(), start_level=5)
    g = golden["pair_example"]
    b1 = gpu_build(ib, ib.bspheres(g["centers1"], g["radii1"]))
    b2 = gpu_build(ib, ib.bspheres(g["centers2"], g["radii2"]))
    tr = ib.traverse(b1, b2, ib.BFSTraversal(), start_level1=g["start_level1"], start_level2=g["start_level2"])
    assert sorted(pairs_list(tr.contacts.numpy())) == sorted(tuple(p) for p in g["contacts_lvt_order"])
    old = ib.traverse(b1, b2, g["start_level1"], start_level2=g["start_level2"])      # traverse(bvh1, bvh2, start_level1, start_level2) = BFS, traverse.jl:244-256
    assert (old.start_level1, old.start_level2, old.num_checks) == (tr.start_level1, tr.start_level2, tr.num_checks)
    assert sorted(pairs_list(old.contacts.numpy())) == sorted(tuple(p) for p in g["contacts_lvt_order"])
    gr = golden["ray_example"]
    tr = ib.traverse_rays(bvh, np.array(gr["points"]), np.array(gr["directions"]), ib.BFSTraversal())
    assert sorted(pairs_list(tr.contacts.numpy())) == sorted(tuple(p) for p in gr["contacts_lvt_order"])
    one = gpu_build(ib, ib.bspheres([[0, 0, 0]], [1.0]))
    assert ib.traverse(one, ib.BFSTraversal()).num_contacts == 0             # traverse_single.jl:17-21


@pytest.mark.gpu
@pytest.mark.parametrize("node", ["bbox", "sphere"])
def test_bfs_single_all_start_levels(ib, O, dev, node):
    """runtests.jl:839-900 / gputests.jl:51-127 with alg = BFSTraversal()."""
    rng = np.random.default_rng(42)
    for n in range(1, 200, 11):
        s = random_spheres(rng, n)
        ol, on = oracle_build(O, s, node)
        bvh = gpu_build(ib, s, node)
        brute = sorted_pairs(O.brute_single(s))
        cache = None
        for sl in range(1, bvh.tree.levels + 1):
            want, checks = O.traverse_bfs_single(ol, on, start_level=sl)
            got = ib.traverse(bvh, ib.BFSTraversal(), start_level=sl, cache=cache)
            assert (sorted_pairs(got.contacts.numpy()) == sorted_pairs(want)).all(), (n, sl)
            assert (sorted_pairs(want) == brute).all()
            assert got.num_checks == checks, (n, sl, got.num_checks, checks)
            cache = got


@pytest.mark.gpu
def test_bfs_pair_all_start_levels(ib, O, dev):
    """runtests.jl:1009-1081 / gputests.jl:132-208 with alg = BFSTraversal(): every stage of traverse_pair.jl:39-140
    (both descend, one side at its last node level, one side already at its leaves — incl. single-leaf trees)."""
    rng = np.random.default_rng(44)
    sizes = [1, 2, 22, 64, 127, 190]
    for node in ("bbox", "sphere"):
        for n1 in sizes:
            for n2 in sizes:
                s1, s2 = random_spheres(rng, n1), random_spheres(rng, n2)
                o1, on1 = oracle_build(O, s1, node)
                o2, on2 = oracle_build(O, s2, node)
                b1, b2 = gpu_build(ib, s1, node), gpu_build(ib, s2, node)
                brute = sorted_pairs(O.brute_pair(s1, s2))
                lv1, lv2 = b1.tree.levels, b2.tree.levels
                for sl1 in sorted({1, max(1, lv1 // 2), max(1, lv1 - 1), lv1}):
                    for sl2 in sorted({1, max(1, lv2 // 2), max(1, lv2 - 1), lv2}):
                        want, checks = O.traverse_bfs_pair(o1, on1, o2, on2, start_level1=sl1, start_level2=sl2)
                        got = ib.traverse(b1, b2, ib.BFSTraversal(), start_level1=sl1, start_level2=sl2)
                        assert (sorted_pairs(got.contacts.numpy()) == brute).all(), (node, n1, n2, sl1, sl2)
                        assert (sorted_pairs(want) == brute).all()
                        assert got.num_checks == checks, (node, n1, n2, sl1, sl2)
                        assert (got.start_level1, got.start_level2) == (sl1, sl2)


@pytest.mark.gpu
def test_bfs_rays_types_partial_builds_and_narrow(ib, O, dev):
    rng = np.random.default_rng(45)
    # rays: raytrace/breadth_first, every start level, both node kinds; Float64 too
    for fbytes in (4, 8):
        for node in ("bbox", "sphere"):
            for n in (1, 7, 190):
                s = random_spheres(rng, n, fbytes=fbytes)
                ol, on = oracle_build(O, s, node)
                f = np.float32 if fbytes == 4 else np.float64
                bvh = ib.BVH(s, ib.BBox(f) if node == "bbox" else ib.BSphere(f))
                p = (6 * rng.random((3, 300))).astype(f)
                d = rng.standard_normal((3, 300)).astype(f)
                for sl in range(1, bvh.tree.levels + 1):
                    want, checks = O.traverse_bfs_rays(ol, on, p, d, start_level=sl)
                    got = ib.traverse_rays(bvh, p, d, ib.BFSTraversal(), start_level=sl)
                    assert (sorted_pairs(got.contacts.numpy()) == sorted_pairs(want)).all(), (fbytes, node, n, sl)
                    assert got.num_checks == checks
                    assert (sorted_pairs(want) == sorted_pairs(O.brute_rays(s, p, d))).all()
    # the reference's default call: Float64 leaves under Float32 nodes, partially built; box leaves; Int64 / UInt64
    s = random_spheres(rng, 3000, fbytes=8, spread=12.0)
    for node_t, onode in ((ib.BBox(np.float32), O.BBOX), (ib.BSphere(np.float32), O.BSPHERE)):
        bvh = ib.BVH(s, node_t, built_level=3)
        ol = O.wrap(s)
        on, _, _ = O.build(ol, onode, node_fbytes=4, built_level=3)
        for sl in (3, 6, bvh.tree.levels):
            want, checks = O.traverse_bfs_single(ol, on, built_level=3, start_level=sl)
            got = ib.traverse(bvh, ib.BFSTraversal(), start_level=sl)
            assert (sorted_pairs(got.contacts.numpy()) == sorted_pairs(want)).all() and got.num_checks == checks
        one = ib.BVH(random_spheres(rng, 1, fbytes=8, spread=12.0), node_t)
        o1 = O.wrap(one.leaves.numpy()["volume"].copy())
        on1, _, _ = O.build(o1, onode, node_fbytes=4)
        want, checks = O.traverse_bfs_pair(ol, on, o1, on1, built_level1=3, start_level1=4, start_level2=1)
        got = ib.traverse(bvh, one, ib.BFSTraversal(), start_level1=4, start_level2=1)     # node (Float32) against leaf volume (Float64)
        assert (sorted_pairs(got.contacts.numpy()) == sorted_pairs(want)).all() and got.num_checks == checks
    lo = (6 * rng.random((500, 3))).astype(np.float32)
    boxes = ib.bboxes(lo, lo + (0.2 + 0.6 * rng.random((500, 3))).astype(np.float32))
    bvh = gpu_build(ib, boxes, "bbox", 8, 8)
    ol, on = oracle_build(O, boxes, "bbox", 8, 8)
    want, checks = O.traverse_bfs_single(ol, on)
    got = ib.traverse(bvh, ib.BFSTraversal())
    assert (sorted_pairs(got.contacts.numpy()) == sorted_pairs(want)).all() and got.num_checks == checks
    # narrow: runtests.jl:1228-1266 / gputests.jl:251-290 — BFS == LVT under the same predicate
    s = random_spheres(rng, 190)
    bvh = gpu_build(ib, s)
    pred = lambda a, b: a["morton"] < b["morton"]
    bfs = ib.traverse(bvh, ib.BFSTraversal(), narrow=pred)
    lvt = ib.traverse(bvh, ib.LVTTraversal(), narrow=pred)
    assert (sorted_pairs(bfs.contacts.numpy()) == sorted_pairs(lvt.contacts.numpy())).all()
    by_index = lambda a, b: a["index"] < b["index"]          # (bv1 is the LEFT leaf in both algorithms: about half survive)
    bfs = ib.traverse(bvh, ib.BFSTraversal(), narrow=by_index)
    lvt = ib.traverse(bvh, ib.LVTTraversal(), narrow=by_index)
    assert 0 < bfs.num_contacts < ib.traverse(bvh).num_contacts
    assert (sorted_pairs(bfs.contacts.numpy()) == sorted_pairs(lvt.contacts.numpy())).all()
    b2 = gpu_build(ib, random_spheres(rng, 64))
    bfs = ib.traverse(bvh, b2, ib.BFSTraversal(), narrow=pred)
    lvt = ib.traverse(bvh, b2, ib.LVTTraversal(), narrow=pred)
    assert (sorted_pairs(bfs.contacts.numpy()) == sorted_pairs(lvt.contacts.numpy())).all()


@pytest.mark.gpu
def test_bfs_100k_against_oracle_and_capacity_protocol(ib, O, dev):
    """configs[0] scene: BFS contacts == the oracle's BFS == the LVT list as sets, num_checks exact; a too-small cache1 is
    grown and only the leaf level repeated."""
    from ibvh_b200 import synth
    n = 100_000
    s = synth.random_spheres_np(n, seed=42)
    ol, on = oracle_build(O, s)
    bvh = gpu_build(ib, s)
    want, checks = O.traverse_bfs_single(ol, on)
    got = ib.traverse(bvh, ib.BFSTraversal())
    assert got.num_checks == checks and got.num_contacts == len(want)
    assert (sorted_pairs(got.contacts.numpy()) == sorted_pairs(want)).all()
    assert (sorted_pairs(want) == sorted_pairs(O.traverse_single(ol, on, num_threads=8))).all()
    small = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(1000, ib.pair_dtype(np.int32), dev), ib.DeviceArray.empty(0, ib.pair_dtype(np.int32), dev))
    again = ib.traverse(bvh, ib.BFSTraversal(), cache=small)
    assert again.num_contacts == len(want) and len(again.cache1) == len(want) and again.num_checks == checks
    assert (sorted_pairs(again.contacts.numpy()) == sorted_pairs(want)).all()
    # pair: two trees of different depth
    s2 = synth.random_spheres_np(30_000, seed=43)
    o2, on2 = oracle_build(O, s2)
    b2 = gpu_build(ib, s2)
    want, checks = O.traverse_bfs_pair(ol, on, o2, on2)
    got = ib.traverse(bvh, b2, ib.BFSTraversal())
    assert got.num_checks == checks
    assert (sorted_pairs(got.contacts.numpy()) == sorted_pairs(want)).all()
    assert (sorted_pairs(want) == sorted_pairs(O.traverse_pair(ol, on, o2, on2, num_threads=8))).all()


@pytest.mark.gpu
def test_sort_contacts_sorted_and_unique(ib, O, dev):
    """the unordered lists sorted on the device == the oracle's list sorted on the host
    (`sort(traversal.contacts)`, gputests.jl:73-78); unique drops repeated pairs."""
    import torch
    from ibvh_b200 import synth
    for n, ibytes, mbytes in ((100_000, 4, 4), (20_000, 8, 8), (3, 4, 4)):
        s = synth.random_spheres_np(n, seed=42) if n > 3 else ib.bspheres([[0, 0, 0], [0, 0, 1], [0, 0, 5]], [0.6, 0.6, 0.1])
        ol, on = oracle_build(O, s, "bbox", ibytes, mbytes)
        bvh = gpu_build(ib, s, "bbox", ibytes, mbytes)
        want = sorted_pairs(O.traverse_single(ol, on, num_threads=8))
        for tr in (ib.traverse(bvh, ordered=False), ib.traverse(bvh, ib.BFSTraversal())):
            out = ib.sort_contacts(tr)
            got = out.contacts.numpy()
            assert out.num_contacts == len(want)
            assert (np.stack([got["a"], got["b"]], 1).astype(np.int64) == want).all(), (n, ibytes)
    # unique: a list that holds every pair three times
    I = np.int32
    tr = ib.traverse(gpu_build(ib, synth.random_spheres_np(50_000, seed=1)), ordered=False)
    k = tr.num_contacts
    trip = torch.cat([tr.contacts.tensor] * 3)
    rep = ib.BVHTraversal(1, 0, 0, 3 * k, ib.DeviceArray(trip, ib.pair_dtype(I)), tr.cache2)
    uni = ib.sort_contacts(rep, unique=True)
    assert uni.num_contacts == k
    ref = ib.sort_contacts(tr).contacts.numpy()
    assert uni.contacts.numpy().tobytes() == ref.tobytes()
    assert ib.sort_contacts(rep).num_contacts == k          # (already unique: sorting again keeps it)
    # indices that do not fit the 64-bit key: refused, list untouched
    bad = np.zeros(4, ib.pair_dtype(np.int64))
    bad["a"] = [1, 2, -3, 4]; bad["b"] = [5, 6, 7, 1 << 40]
    d = ib.DeviceArray.from_numpy(bad, device=dev)
    with pytest.raises(Exception):
        ib.sort_contacts(ib.BVHTraversal(1, 0, 0, 4, d, d))
    assert d.numpy().tobytes() == bad.tobytes()
