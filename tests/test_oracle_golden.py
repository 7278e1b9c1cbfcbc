"""The oracle against every known answer the reference's own tests / doctests hold for the hot path
(SURVEY.md §8c). CPU only."""
import numpy as np
import pytest

from conftest import pairs_list


def test_implicit_tree_known_answers(O, golden):
    for case in golden["implicit_tree"]:
        t = O.tree_shape(case["n"])
        for k in ("levels", "real_leaves", "virtual_leaves", "real_nodes", "virtual_nodes"):
            assert t[k] == case[k], (case["source"], k)
        for idx, want in case["memory_index"]:
            assert O.memory_index(case["n"], idx) == want
        for lvl, want in case["level_indices"]:
            assert list(O.level_indices(case["n"], lvl)) == want
        for idx, want in case["isvirtual"]:
            assert O.isvirtual(case["n"], idx) == want


def test_tree_domain_error(O):
    with pytest.raises(ValueError):
        O.tree_shape(0)


def test_morton_split3(O, golden):
    g = golden["morton_split3"]
    for bits in g["bits"]:
        assert O.morton_split3(g["input"], bits) == g["output"]


def test_build_level(O, golden):
    g = golden["build_level"]
    for c in g["cases"]:
        if isinstance(c["built_level"], float):
            assert O.build_level_float(g["levels"], c["built_level"]) == c["expect"]


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_five_spheres_doctest(O, golden, variant):
    g = golden["five_spheres"]
    v = g["variants"][variant]
    s = O.spheres(g["centers"], g["radii"], v["float"])
    leaves = O.wrap(s, v["index"], v["morton"])
    nodes, _, _ = O.build(leaves, O.BBOX, v["node_float"])
    c = O.traverse_single(leaves, nodes)
    assert pairs_list(c) == [tuple(p) for p in g["contacts_lvt_order"]]


def test_five_spheres_sphere_nodes(O, golden):
    g = golden["five_spheres"]
    for fb in (4, 8):
        s = O.spheres(g["centers"], g["radii"], fb)
        leaves = O.wrap(s)
        nodes, _, _ = O.build(leaves, O.BSPHERE, fb)
        assert len(nodes) == 6
        assert pairs_list(O.traverse_single(leaves, nodes)) == [tuple(p) for p in g["contacts_lvt_order"]]


def test_shuffled_doctest_morton(O, golden):
    g = golden["five_spheres_shuffled"]
    s = O.spheres(g["centers"], g["radii"], 4)
    leaves = O.wrap(s)
    leaves["index"] = g["indices"]
    O.build(leaves)
    first = leaves[0]
    want = g["first_sorted_leaf"]
    assert first["index"] == want["index"]
    assert int(first["morton"]) == int(want["morton"], 16)
    assert first["volume"]["x"].tolist() == want["x"] and float(first["volume"]["r"]) == want["r"]


def test_unordered_structure(O, golden):
    """test/runtests.jl:718-834: node structure of the 5-leaf tree for unordered input."""
    g = golden["unordered_contacts"]
    for fb in (8, 4):
        s = O.spheres(g["centers"], g["radii"], fb)
        # BSphere nodes
        leaves = O.wrap(s)
        nodes, _, _ = O.build(leaves, O.BSPHERE, fb)
        assert len(nodes) == g["num_nodes"]

        def m(a, b):
            out = np.zeros(4, np.float64)
            aa = np.array([*s[a - 1]["x"], s[a - 1]["r"]], np.float64) if isinstance(a, int) else a
            bb = np.array([*s[b - 1]["x"], s[b - 1]["r"]], np.float64) if isinstance(b, int) else b
            O.lib().orc_merge_sphere_f64(aa.ctypes.data, bb.ctypes.data, out.ctypes.data)
            return out
        n4, n5 = m(3, 1), m(2, 5)
        tol = 1e-5 if fb == 4 else 1e-12
        assert np.allclose(nodes[3]["x"], n4[:3], atol=tol) and np.allclose(nodes[4]["x"], n5[:3], atol=tol)
        assert np.allclose(nodes[5]["x"], s[3]["x"])
        n2 = m(n4, n5)
        assert np.allclose(nodes[1]["x"], n2[:3], atol=tol) and np.allclose(nodes[2]["x"], s[3]["x"])
        s4 = np.array([*s[3]["x"], s[3]["r"]], np.float64)
        assert np.allclose(nodes[0]["x"], m(n2, s4)[:3], atol=tol)
        got = set(pairs_list(O.traverse_single(leaves, nodes)))
        assert got == {tuple(p) for p in g["contacts_set"]}
        # BBox nodes
        leaves = O.wrap(s)
        nodes, _, _ = O.build(leaves, O.BBOX, fb)
        assert len(nodes) == g["num_nodes"]
        assert set(pairs_list(O.traverse_single(leaves, nodes))) == {tuple(p) for p in g["contacts_set"]}
        # box leaves (bvh_single_bbox_small_unordered)
        b = O.boxes_of_spheres(s)
        leaves = O.wrap(b)
        nodes, _, _ = O.build(leaves, O.BBOX, fb)
        assert set(pairs_list(O.traverse_single(leaves, nodes))) == {tuple(p) for p in g["contacts_set"]}


def test_pair_doctest(O, golden):
    g = golden["pair_example"]
    l1 = O.wrap(O.spheres(g["centers1"], g["radii1"]))
    l2 = O.wrap(O.spheres(g["centers2"], g["radii2"]))
    n1, _, _ = O.build(l1)
    n2, _, _ = O.build(l2)
    c = O.traverse_pair(l1, n1, l2, n2, start_level1=g["start_level1"], start_level2=g["start_level2"])
    assert pairs_list(c) == [tuple(p) for p in g["contacts_lvt_order"]]


def test_ray_doctest(O, golden):
    g5, g = golden["five_spheres"], golden["ray_example"]
    leaves = O.wrap(O.spheres(g5["centers"], g5["radii"]))
    nodes, _, _ = O.build(leaves)
    c = O.traverse_rays(leaves, nodes, np.array(g["points"]), np.array(g["directions"]))
    assert pairs_list(c) == [tuple(p) for p in g["contacts_lvt_order"]]


def test_ray_box_truth_table(O, golden):
    g = golden["ray_box"]
    box = np.array(g["box"], np.float64).reshape(6)
    for c in g["cases"]:
        p, d = np.array(c["p"], np.float64), np.array(c["d"], np.float64)
        assert bool(O.lib().orc_ray_box_f64(box.ctypes.data, p.ctypes.data, d.ctypes.data)) == c["hit"], c


def test_ray_sphere_truth_table(O, golden):
    g = golden["ray_sphere"]
    s = np.array(g["sphere"], np.float64)
    for c in g["cases"]:
        p, d = np.array(c["p"], np.float64), np.array(c["d"], np.float64)
        assert bool(O.lib().orc_ray_sphere_f64(s.ctypes.data, p.ctypes.data, d.ctypes.data)) == c["hit"], c


def triangle_sphere(p1, p2, p3):
    """BSphere{Float64}(p1, p2, p3) — src/bounding_volumes/bsphere.jl:43-112 (input preparation, only
    needed to rebuild the spheres the reference's ray tests use)."""
    a, b, c = (np.array(p, np.float64) for p in (p1, p2, p3))
    abab = float(np.dot(b - a, b - a)); abac = float(np.dot(b - a, c - a)); acac = float(np.dot(c - a, c - a))
    d = 2.0 * (abab * acac - abac * abac)
    if abs(d) <= np.finfo(np.float64).eps:
        lo, up = np.minimum(np.minimum(a, b), c), np.maximum(np.maximum(a, b), c)
        ctr = 0.5 * (lo + up)
        return np.array([*ctr, np.linalg.norm(ctr - up)])
    s = (abab * acac - acac * abac) / d
    t = (acac * abab - abab * abac) / d
    if s <= 0:
        ctr = 0.5 * (a + c); r = np.linalg.norm(ctr - a)
    elif t <= 0:
        ctr = 0.5 * (a + b); r = np.linalg.norm(ctr - a)
    elif s + t >= 1:
        ctr = 0.5 * (b + c); r = np.linalg.norm(ctr - b)
    else:
        ctr = a + s * (b - a) + t * (c - a); r = np.linalg.norm(ctr - a)
    return np.array([*ctr, r])


def test_ray_sphere_triangle_cases(O, golden):
    g = golden["ray_sphere_triangle_cases"]
    d = np.array(g["direction"], np.float64)
    for tri in g["triangles"]:
        s = triangle_sphere(*tri)
        for p in g["points"]:
            p = np.array(p, np.float64)
            assert O.lib().orc_ray_sphere_f64(s.ctypes.data, p.ctypes.data, d.ctypes.data) == 1
            nd = -d
            assert O.lib().orc_ray_sphere_f64(s.ctypes.data, p.ctypes.data, nd.ctypes.data) == 1


def test_sphere_and_box_merges(O, golden):
    for c in golden["sphere_merges"]["cases"]:
        a, b = np.array(c["a"], np.float64), np.array(c["b"], np.float64)
        out = np.zeros(4)
        O.lib().orc_merge_sphere_f64(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        assert np.allclose(out, c["c"], rtol=1e-12), c
    for c in golden["box_merges"]["cases"]:
        a, b = np.array(c["a"], np.float64), np.array(c["b"], np.float64)
        out = np.zeros(6)
        O.lib().orc_merge_box_f64(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        assert np.allclose(out, c["c"], rtol=1e-12), c


def test_ray_grid_against_analytic_sphere(O, golden):
    """test/runtests.jl:1086-1225 — exact order of the ray ids, six axis directions."""
    g = golden["ray_grid"]
    sph = triangle_sphere(*g["triangle"])
    x, r = sph[:3], sph[3]
    s = O.spheres([x], [r], 8)
    leaves = O.wrap(s)
    nodes, _, _ = O.build(leaves, O.BBOX, 8)
    rng = [np.arange(x[k] - r, x[k] + r + 1e-12, 1.0) for k in range(3)]
    # Julia comprehension order: x fastest, then y, then z
    pts = np.array([[px, py, pz] for pz in rng[2] for py in rng[1] for px in rng[0]]).T
    for axis in range(3):
        for sign in (1.0, -1.0):
            d = np.zeros_like(pts)
            d[axis, :] = sign
            got = [int(b) for b in O.traverse_rays(leaves, nodes, pts, d)["b"]]
            others = [k for k in range(3) if k != axis]
            want = []
            for i in range(pts.shape[1]):
                p = pts[:, i]
                behind = p[axis] <= x[axis] if sign > 0 else p[axis] >= x[axis]
                if behind and np.linalg.norm(p[others] - x[others]) <= r:
                    want.append(i + 1)
                elif (not behind) and np.linalg.norm(p - x) <= r:
                    want.append(i + 1)
            assert got == want, (axis, sign)


def test_extrema_bracket_centres(O):
    """test/runtests.jl:510-559: padded extrema strictly bracket all centres, incl. degenerate inputs."""
    rng = np.random.default_rng(42)
    cases = []
    for fb, scale in ((4, 1000.0), (8, 1000.0), (4, 1.0)):
        f = {4: np.float32, 8: np.float64}[fb]
        cases.append(O.spheres((scale * rng.random((100, 3))).astype(f), rng.random(100), fb))
    cases.append(O.spheres([[0, 0, 0]], [1.0], 8))
    cases.append(O.spheres([[1000, 0, 0], [1000, 0, 0]], [1.0, 1.0], 8))
    cases.append(O.spheres([[-5, -7, -9], [-1, -2, -3]], [1.0, 1.0], 4))
    for s in cases:
        leaves = O.wrap(s)
        mn, mx = O.morton_encode(leaves)
        for k in range(3):
            assert (s["x"][:, k] > mn[k]).all() and (s["x"][:, k] < mx[k]).all()


def test_max_seed_quirk(O):
    """the max reduction is seeded with floatmin, so an all-negative axis
    reports max = floatmin (+ padding), not the true maximum."""
    s = O.spheres([[-5, -7, -9], [-1, -2, -3]], [1.0, 1.0], 4)
    leaves = O.wrap(s)
    mn, mx = O.morton_encode(leaves)
    fm = np.finfo(np.float32).tiny
    assert (mx == np.float32(np.float32(fm + np.float32(1e-5) * fm) + fm)).all()


def test_triangle_volume_constructors(O, golden):
    """BSphere{T}(p1,p2,p3) / BBox{T}(p1,p2,p3) known answers (runtests.jl:185-210, 260-276) and agreement with the
    independent numpy restatement above."""
    g = golden["triangle_volumes"]
    for c in g["bsphere"]:
        s = O.volumes_from_triangles([c["tri"]], O.BSPHERE, 8)[0]
        assert np.allclose(s["x"], c["x"], rtol=1e-12, atol=1e-15) and np.isclose(s["r"], c["r"], rtol=1e-12), c
        ref = triangle_sphere(*c["tri"])
        assert np.allclose([*s["x"], s["r"]], ref, rtol=1e-12, atol=1e-15)
    for c in g["bbox"]:
        b = O.volumes_from_triangles([c["tri"]], O.BBOX, 8)[0]
        assert b["lo"].tolist() == c["lo"] and b["up"].tolist() == c["up"]
    rng = np.random.default_rng(3)
    tris = rng.random((500, 3, 3))
    tris[:20, 2] = tris[:20, 0] + 0.5 * (tris[:20, 1] - tris[:20, 0])          # degenerate (collinear) triangles
    s = O.volumes_from_triangles(tris, O.BSPHERE, 8)
    for i in range(len(tris)):
        ref = triangle_sphere(*tris[i])
        assert np.allclose([*s[i]["x"], s[i]["r"]], ref, rtol=1e-9, atol=1e-12), i
        d = np.linalg.norm(tris[i] - s[i]["x"], axis=1)
        assert (d <= s[i]["r"] * (1 + 1e-9) + 1e-12).all(), "the sphere encloses the three vertices"
