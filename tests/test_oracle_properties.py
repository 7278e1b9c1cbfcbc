"""Properties the reference's randomised tests assert (test/runtests.jl:839-1081, 1230-1270), re-run on
our own seeded inputs against the oracle. CPU only."""
import numpy as np
import pytest

from conftest import random_spheres, sorted_pairs


def _to_input_pairs_single(contacts):
    return sorted_pairs(contacts)


@pytest.mark.parametrize("node_kind", ["sphere", "bbox"])
@pytest.mark.parametrize("fbytes", [4, 8])
def test_single_equals_brute_force_all_start_levels(O, node_kind, fbytes):
    """runtests.jl:839-900: n in 1:11:200, every start_level; LVT == O(n^2) brute force (as sets)."""
    rng = np.random.default_rng(42)
    nk = O.BSPHERE if node_kind == "sphere" else O.BBOX
    for n in range(1, 200, 11):
        levels = O.tree_shape(n)["levels"]
        for start_level in range(1, levels + 1):
            s = random_spheres(rng, n, fbytes)
            brute = sorted_pairs(O.brute_single(s))
            leaves = O.wrap(s)
            nodes, _, _ = O.build(leaves, nk, fbytes)
            got = sorted_pairs(O.traverse_single(leaves, nodes, start_level=start_level))
            assert got.shape == brute.shape and (got == brute).all(), (n, start_level)


def test_pair_self_equivalent_to_single(O):
    """runtests.jl:936-1004: traverse(bvh, bvh) == single contacts + diagonal + mirrored pairs."""
    rng = np.random.default_rng(43)
    for n in range(1, 200, 33):
        levels = O.tree_shape(n)["levels"]
        for sl1 in range(1, levels + 1, 2):
            for sl2 in range(1, levels + 1):
                s = random_spheres(rng, n)
                leaves = O.wrap(s)
                nodes, _, _ = O.build(leaves, O.BBOX)
                c1 = sorted_pairs(O.traverse_single(leaves, nodes, start_level=sl1))
                c2 = sorted_pairs(O.traverse_pair(leaves, nodes, leaves, nodes, start_level1=sl1, start_level2=sl2))
                s2 = {tuple(p) for p in c2.tolist()}
                assert all((i, i) in s2 for i in range(1, n + 1))
                off = {p for p in s2 if p[0] != p[1]}
                assert all((j, i) in off for (i, j) in off)
                upper = sorted(p for p in off if p[0] < p[1])
                assert upper == [tuple(p) for p in c1.tolist()]


@pytest.mark.parametrize("node_kind", ["sphere", "bbox"])
def test_pair_equals_brute_force(O, node_kind):
    """runtests.jl:1009-1081: n1, n2 in 1:21:200 (here a thinned grid), every (start_level1, start_level2)."""
    rng = np.random.default_rng(44)
    nk = O.BSPHERE if node_kind == "sphere" else O.BBOX
    sizes = [1, 22, 64, 127, 190]
    for n1 in sizes:
        for n2 in sizes:
            l1n, l2n = O.tree_shape(n1)["levels"], O.tree_shape(n2)["levels"]
            for sl1 in {1, l1n}:
                for sl2 in range(1, l2n + 1, 2):
                    s1, s2 = random_spheres(rng, n1), random_spheres(rng, n2)
                    brute = sorted_pairs(O.brute_pair(s1, s2))
                    a, b = O.wrap(s1), O.wrap(s2)
                    na, _, _ = O.build(a, nk)
                    nb, _, _ = O.build(b, nk)
                    # contacts carry .index == original position, so they compare directly with brute force
                    got = sorted_pairs(O.traverse_pair(a, na, b, nb, start_level1=sl1, start_level2=sl2))
                    assert got.shape == brute.shape and (got == brute).all(), (n1, n2, sl1, sl2)


def test_rays_equal_brute_force(O):
    rng = np.random.default_rng(45)
    for n in (1, 2, 37, 200):
        for leaf in ("sphere", "box"):
            s = random_spheres(rng, n)
            vols = s if leaf == "sphere" else O.boxes_of_spheres(s)
            leaves = O.wrap(vols)
            nodes, _, _ = O.build(leaves, O.BBOX)
            R = 300
            p = (8 * rng.random((3, R)) - 1).astype(np.float32)
            d = (rng.random((3, R)) - 0.5).astype(np.float32)
            d[:, :5] = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, 0, -1]], np.float32).T   # axis rays: 1/0 = Inf
            brute = sorted_pairs(O.brute_rays(vols, p, d))
            levels = O.tree_shape(n)["levels"]
            for sl in range(1, levels + 1):
                got = sorted_pairs(O.traverse_rays(leaves, nodes, p, d, start_level=sl))
                assert got.shape == brute.shape and (got == brute).all(), (n, leaf, sl)


def test_built_level_partial_trees(O):
    """Partial builds: nodes above built_level are untouched; traversal from start_level >= built_level
    gives the same contacts."""
    rng = np.random.default_rng(46)
    n = 150
    s = random_spheres(rng, n)
    full_leaves = O.wrap(s)
    full_nodes, _, _ = O.build(full_leaves, O.BBOX)
    want = sorted_pairs(O.traverse_single(full_leaves, full_nodes))
    levels = O.tree_shape(n)["levels"]
    for bl in range(1, levels + 1):
        leaves = O.wrap(s)
        nodes, _, _ = O.build(leaves, O.BBOX, built_level=bl)
        got = sorted_pairs(O.traverse_single(leaves, nodes, built_level=bl, start_level=bl))
        assert (got == want).all()
        # levels >= min(bl, levels-1) are identical to the full build, levels above are left alone (zeros here)
        lo, _ = O.level_indices(n, min(bl, levels - 1))
        assert (nodes[lo - 1:] == full_nodes[lo - 1:]).all()
        assert (nodes[: lo - 1]["lo"] == 0).all()


def test_multithreaded_oracle_matches_single_thread(O):
    rng = np.random.default_rng(47)
    s = random_spheres(rng, 5000, spread=20.0)
    a, b = O.wrap(s), O.wrap(s)
    na, _, _ = O.build(a, O.BBOX, num_threads=1)
    nb, _, _ = O.build(b, O.BBOX, num_threads=8, min_elems=10)
    assert (a == b).all() and (na == nb).all()
    ca = O.traverse_single(a, na, num_threads=1)
    cb = O.traverse_single(b, nb, num_threads=8, min_elems=10)
    assert (ca == cb).all()       # same order too: tasks are contiguous ascending ranges


def test_stable_sort_ties_keep_input_order(O):
    """Tie rule adopted in SURVEY.md §8c: equal Morton keys keep input order."""
    s = O.spheres(np.zeros((64, 3)) + 0.25, np.full(64, 0.1))
    s["x"][32:] = 0.75
    leaves = O.wrap(s, mbytes=2)
    O.build(leaves)
    idx = leaves["index"]
    assert (np.diff(leaves["morton"].astype(np.int64)) >= 0).all()
    for m in np.unique(leaves["morton"]):
        grp = idx[leaves["morton"] == m]
        assert (np.diff(grp) > 0).all()
