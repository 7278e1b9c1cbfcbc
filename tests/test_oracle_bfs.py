"""The oracle's BFS traversals (src/traverse/breadth_first, src/raytrace/breadth_first) against brute force, the
reference's known answers and the oracle's own LVT traversals — the reference runs every traversal test over
`for alg in (BFSTraversal(), LVTTraversal())` (test/runtests.jl:600-1225, 1228-1266)."""
import numpy as np

from conftest import pairs_list, random_spheres, sorted_pairs


def build(O, s, node="bbox", built_level=1, ib=4, mb=4):
    leaves = O.wrap(s, ib, mb)
    nodes, _, _ = O.build(leaves, O.BBOX if node == "bbox" else O.BSPHERE, built_level=built_level)
    return leaves, nodes


def test_bfs_known_answers(O, golden):
    g = golden["five_spheres"]                          # runtests.jl:600-668: the three contacts, any order
    for node in ("bbox", "sphere"):
        leaves, nodes = build(O, O.spheres(g["centers"], g["radii"]), node)
        for sl in range(1, 5):
            c, checks = O.traverse_bfs_single(leaves, nodes, start_level=sl)
            assert sorted(pairs_list(c)) == sorted(tuple(p) for p in g["contacts_lvt_order"])
            assert checks >= len(c)
    assert O.bfs_default_start_level(5) == 2 and O.bfs_default_start_level(5, 3) == 3     # breadth_first.jl:4-6
    g = golden["pair_example"]                          # traverse.jl:183-208
    l1, n1 = build(O, O.spheres(g["centers1"], g["radii1"]))
    l2, n2 = build(O, O.spheres(g["centers2"], g["radii2"]))
    c, _ = O.traverse_bfs_pair(l1, n1, l2, n2, start_level1=g["start_level1"], start_level2=g["start_level2"])
    assert sorted(pairs_list(c)) == sorted(tuple(p) for p in g["contacts_lvt_order"])
    g5, gr = golden["five_spheres"], golden["ray_example"]          # raytrace.jl:33-69
    leaves, nodes = build(O, O.spheres(g5["centers"], g5["radii"]))
    c, _ = O.traverse_bfs_rays(leaves, nodes, np.array(gr["points"], np.float32), np.array(gr["directions"], np.float32))
    assert sorted(pairs_list(c)) == sorted(tuple(p) for p in gr["contacts_lvt_order"])


def test_bfs_initial_bvtt_is_the_triangle_of_the_start_level(O):
    """traverse_single.jl:137-156: self-checks + all pairs of the real nodes; no self-checks at the leaf level."""
    rng = np.random.default_rng(5)
    s = random_spheres(rng, 11, spread=100.0)           # far apart: (almost) no node contacts
    leaves, nodes = build(O, s)
    levels = O.tree_shape(11)["levels"]
    _, checks = O.traverse_bfs_single(leaves, nodes, start_level=levels)
    assert checks == 11 * 10 // 2


def test_bfs_single_equals_brute_force_all_start_levels(O):
    rng = np.random.default_rng(42)
    for node in ("bbox", "sphere"):
        for n in range(1, 200, 11):
            s = random_spheres(rng, n)
            leaves, nodes = build(O, s, node)
            brute = sorted_pairs(O.brute_single(s))
            lvt = sorted_pairs(O.traverse_single(leaves, nodes))
            assert (lvt == brute).all()
            prev_checks = None
            for sl in range(1, O.tree_shape(n)["levels"] + 1):
                c, checks = O.traverse_bfs_single(leaves, nodes, start_level=sl)
                assert (sorted_pairs(c) == brute).all(), (node, n, sl)
                assert (c["a"] < c["b"]).all()
                prev_checks = checks
            pos, _ = O.traverse_bfs_single(leaves, nodes, positions=True)
            assert len(pos) == len(brute)
            if len(pos):
                a, b = leaves["index"][pos["a"] - 1], leaves["index"][pos["b"] - 1]
                assert (pos["a"] < pos["b"]).all()
                assert (sorted_pairs(np.stack([np.minimum(a, b), np.maximum(a, b)], 1)) == brute).all()


def test_bfs_pair_equals_brute_force(O):
    rng = np.random.default_rng(44)
    sizes = [1, 2, 22, 64, 127, 190]
    for node in ("bbox", "sphere"):
        for n1 in sizes:
            for n2 in sizes:
                s1, s2 = random_spheres(rng, n1), random_spheres(rng, n2)
                l1, nd1 = build(O, s1, node)
                l2, nd2 = build(O, s2, node)
                brute = sorted_pairs(O.brute_pair(s1, s2))
                lv1, lv2 = O.tree_shape(n1)["levels"], O.tree_shape(n2)["levels"]
                for sl1 in sorted({1, max(1, lv1 // 2), max(1, lv1 - 1), lv1}):
                    for sl2 in sorted({1, max(1, lv2 // 2), max(1, lv2 - 1), lv2}):
                        c, checks = O.traverse_bfs_pair(l1, nd1, l2, nd2, start_level1=sl1, start_level2=sl2)
                        assert (sorted_pairs(c) == brute).all(), (node, n1, n2, sl1, sl2)
                        assert checks >= len(c)


def test_bfs_partial_builds_and_mixed_float_types(O):
    rng = np.random.default_rng(3)
    s = random_spheres(rng, 300, fbytes=8)
    brute = sorted_pairs(O.brute_single(s))
    for node in (O.BBOX, O.BSPHERE):
        leaves = O.wrap(s)
        nodes, _, _ = O.build(leaves, node, node_fbytes=4, built_level=3)      # the reference's default call shape: F64 leaves, F32 nodes
        for sl in (3, 5, O.tree_shape(300)["levels"]):
            c, _ = O.traverse_bfs_single(leaves, nodes, built_level=3, start_level=sl)
            assert (sorted_pairs(c) == brute).all()
        s2 = random_spheres(rng, 1, fbytes=8)
        l2 = O.wrap(s2)
        n2, _, _ = O.build(l2, node, node_fbytes=4)
        c, _ = O.traverse_bfs_pair(leaves, nodes, l2, n2, built_level1=3, start_level1=4, start_level2=1)   # node (F32) vs leaf volume (F64)
        assert (sorted_pairs(c) == sorted_pairs(O.brute_pair(s, s2))).all()
        c, _ = O.traverse_bfs_pair(l2, n2, leaves, nodes, built_level2=3, start_level1=1, start_level2=3)
        assert (sorted_pairs(c) == sorted_pairs(O.brute_pair(s2, s))).all()


def test_bfs_rays_equal_brute_force(O):
    rng = np.random.default_rng(45)
    for node in ("bbox", "sphere"):
        for n in (1, 7, 64, 190):
            s = random_spheres(rng, n)
            leaves, nodes = build(O, s, node)
            p = (6 * rng.random((3, 50))).astype(np.float32)
            d = rng.standard_normal((3, 50)).astype(np.float32)
            brute = sorted_pairs(O.brute_rays(s, p, d))
            for sl in range(1, O.tree_shape(n)["levels"] + 1):
                c, checks = O.traverse_bfs_rays(leaves, nodes, p, d, start_level=sl)
                assert (sorted_pairs(c) == brute).all(), (node, n, sl)
            lvt = O.traverse_rays(leaves, nodes, p, d)
            assert (sorted_pairs(lvt) == brute).all()
