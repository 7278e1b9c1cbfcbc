"""ctypes binding of the CPU oracle (oracle/_build/libibvh_oracle.so).

TEST INFRASTRUCTURE ONLY. Importers allowed: tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / ``--impl reference`` legs. The product package never imports this module.

All arrays are numpy; leaves are structured arrays with the reference's isbits layouts
(`leaf_dtype`), indices on the API surface are 1-based as in the reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libibvh_oracle.so")

BSPHERE, BBOX = 0, 1


def compile_lib(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++, -ffp-contract=off)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("ibvh_oracle.hpp", "ibvh_oracle_capi.cpp", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        compile_lib()
        _lib = C.CDLL(_SO)
        i64, vp, ci = C.c_int64, C.c_void_p, C.c_int
        _lib.orc_tree_shape.argtypes = [i64, vp, vp]
        _lib.orc_memory_index.restype = i64
        _lib.orc_memory_index.argtypes = [i64, i64]
        _lib.orc_level_indices.argtypes = [i64, i64, vp, vp]
        _lib.orc_isvirtual.argtypes = [i64, i64]
        _lib.orc_build_level_float.restype = i64
        _lib.orc_build_level_float.argtypes = [i64, C.c_double]
        _lib.orc_morton_split3_u16.restype = C.c_uint16
        _lib.orc_morton_split3_u16.argtypes = [C.c_uint16]
        _lib.orc_morton_split3_u32.restype = C.c_uint32
        _lib.orc_morton_split3_u32.argtypes = [C.c_uint32]
        _lib.orc_morton_split3_u64.restype = C.c_uint64
        _lib.orc_morton_split3_u64.argtypes = [C.c_uint64]
        for name in ("orc_ray_box_f64", "orc_ray_sphere_f64", "orc_ray_box_f32", "orc_ray_sphere_f32"):
            getattr(_lib, name).argtypes = [vp, vp, vp]
        for name in ("orc_merge_sphere_f64", "orc_merge_sphere_f32", "orc_merge_box_f64",
                     "orc_merge_spheres_to_box_f64", "orc_merge_spheres_to_box_f32"):
            getattr(_lib, name).argtypes = [vp, vp, vp]
            getattr(_lib, name).restype = None
        _lib.orc_contact_sphere_f32.argtypes = [vp, vp]
        _lib.orc_contact_sphere_f64.argtypes = [vp, vp]
        _lib.orc_volumes_from_triangles.argtypes = [vp, i64, ci, ci, vp]
        _lib.orc_leaf_bytes.restype = i64
        _lib.orc_leaf_bytes.argtypes = [ci, ci, ci, ci]
        _lib.orc_wrap.argtypes = [vp, i64, ci, ci, ci, ci, vp]
        _lib.orc_morton_encode.argtypes = [vp, i64, ci, ci, ci, ci, ci, vp, vp, ci, i64]
        _lib.orc_sort_leaves.argtypes = [vp, i64, ci, ci, ci, ci, ci, i64]
        _lib.orc_aggregate.argtypes = [vp, i64, ci, ci, ci, ci, vp, ci, ci, i64, ci, i64]
        _lib.orc_build.argtypes = [vp, i64, ci, ci, ci, ci, vp, ci, ci, i64, ci, vp, vp, ci, i64, i64, i64]
        _lib.orc_traverse_single.restype = i64
        _lib.orc_traverse_single.argtypes = [vp, i64, ci, ci, ci, ci, vp, ci, ci, i64, i64, vp, i64, vp, ci, i64]
        _lib.orc_traverse_pair.restype = i64
        _lib.orc_traverse_pair.argtypes = [vp, i64, vp, i64, i64, vp, i64, vp, i64, i64, ci, ci, ci, ci, ci, ci, vp, i64, vp, ci, i64]
        _lib.orc_traverse_rays.restype = i64
        _lib.orc_traverse_rays.argtypes = [vp, i64, ci, ci, ci, ci, vp, ci, ci, i64, i64, vp, vp, ci, i64, vp, i64, vp, ci, i64]
        _lib.orc_bfs_fetch.restype = i64
        _lib.orc_bfs_fetch.argtypes = [vp, i64]
        _lib.orc_traverse_bfs_single.restype = i64
        _lib.orc_traverse_bfs_single.argtypes = [vp, i64, ci, ci, ci, ci, vp, ci, ci, i64, i64, ci, vp]
        _lib.orc_traverse_bfs_pair.restype = i64
        _lib.orc_traverse_bfs_pair.argtypes = [vp, i64, vp, i64, i64, vp, i64, vp, i64, i64, ci, ci, ci, ci, ci, ci, ci, vp]
        _lib.orc_traverse_bfs_rays.restype = i64
        _lib.orc_traverse_bfs_rays.argtypes = [vp, i64, ci, ci, ci, ci, vp, ci, ci, i64, i64, vp, vp, i64, ci, vp]
        _lib.orc_brute_single.restype = i64
        _lib.orc_brute_single.argtypes = [vp, i64, ci, ci, vp, i64]
        _lib.orc_brute_pair.restype = i64
        _lib.orc_brute_pair.argtypes = [vp, i64, vp, i64, ci, ci, vp, i64]
        _lib.orc_brute_rays.restype = i64
        _lib.orc_brute_rays.argtypes = [vp, i64, ci, ci, vp, vp, i64, vp, i64]
    return _lib


# ---------------------------------------------------------------------------------------------
# dtypes mirroring the reference's isbits layouts (SURVEY.md §8 layout table)
# ---------------------------------------------------------------------------------------------
def _f(fbytes):
    return {4: np.float32, 8: np.float64}[fbytes]


def volume_dtype(kind: int, fbytes: int = 4) -> np.dtype:
    f = _f(fbytes)
    if kind == BSPHERE:
        return np.dtype([("x", f, 3), ("r", f)])
    return np.dtype([("lo", f, 3), ("up", f, 3)])


def leaf_dtype(kind: int, fbytes: int = 4, ibytes: int = 4, mbytes: int = 4) -> np.dtype:
    i = {4: np.int32, 8: np.int64}[ibytes]
    m = {2: np.uint16, 4: np.uint32, 8: np.uint64}[mbytes]
    dt = np.dtype([("volume", volume_dtype(kind, fbytes)), ("index", i), ("morton", m)], align=True)
    assert dt.itemsize == lib().orc_leaf_bytes(kind, fbytes, ibytes, mbytes), (dt.itemsize, kind, fbytes, ibytes, mbytes)
    return dt


def pair_dtype(ibytes: int = 4) -> np.dtype:
    i = {4: np.int32, 8: np.int64}[ibytes]
    return np.dtype([("a", i), ("b", i)])


def _desc(leaves: np.ndarray):
    dt = leaves.dtype
    vol = dt["volume"]
    kind = BSPHERE if "r" in vol.names else BBOX
    fbytes = vol[vol.names[0]].base.itemsize
    return kind, fbytes, dt["index"].itemsize, dt["morton"].itemsize


def _vdesc(vols: np.ndarray):
    vol = vols.dtype
    kind = BSPHERE if "r" in vol.names else BBOX
    return kind, vol[vol.names[0]].base.itemsize


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    pass


def _check(rc, what):
    if rc < 0:
        raise OracleError(f"{what}: oracle status {rc}")
    return rc


# ---------------------------------------------------------------------------------------------
# API
# ---------------------------------------------------------------------------------------------
def tree_shape(n: int):
    tree = np.zeros(5, np.int64)
    skips = np.zeros(64, np.int64)
    rc = lib().orc_tree_shape(n, _p(tree), _p(skips))
    if rc == -2:
        raise ValueError("DomainError: must have at least one geometry!")
    d = dict(zip(("levels", "real_leaves", "real_nodes", "virtual_leaves", "virtual_nodes"), map(int, tree)))
    d["skips"] = skips[: d["levels"]].copy()
    return d


def memory_index(n, idx):
    return int(lib().orc_memory_index(n, idx))


def level_indices(n, level):
    a, b = C.c_int64(), C.c_int64()
    lib().orc_level_indices(n, level, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def isvirtual(n, idx):
    return bool(lib().orc_isvirtual(n, idx))


def build_level_float(levels, f):
    return int(lib().orc_build_level_float(levels, float(f)))


def morton_split3(v, bits):
    return int(getattr(lib(), f"orc_morton_split3_u{bits}")(v))


def wrap(volumes: np.ndarray, ibytes=4, mbytes=4) -> np.ndarray:
    kind, fbytes = _vdesc(volumes)
    volumes = np.ascontiguousarray(volumes)
    out = np.zeros(len(volumes), leaf_dtype(kind, fbytes, ibytes, mbytes))
    _check(lib().orc_wrap(_p(volumes), len(volumes), kind, fbytes, ibytes, mbytes, _p(out)), "wrap")
    return out


def morton_encode(leaves: np.ndarray, compute_extrema=True, mins=None, maxs=None, num_threads=1, min_elems=100):
    kind, fbytes, ib, mb = _desc(leaves)
    f = _f(fbytes)
    mn = np.zeros(3, f) if mins is None else np.asarray(mins, f).copy()
    mx = np.zeros(3, f) if maxs is None else np.asarray(maxs, f).copy()
    _check(lib().orc_morton_encode(_p(leaves), len(leaves), kind, fbytes, ib, mb, int(compute_extrema), _p(mn), _p(mx),
                                   num_threads, min_elems), "morton_encode")
    return mn, mx


def sort_leaves(leaves: np.ndarray, num_threads=1, min_elems=100):
    kind, fbytes, ib, mb = _desc(leaves)
    _check(lib().orc_sort_leaves(_p(leaves), len(leaves), kind, fbytes, ib, mb, num_threads, min_elems), "sort")


def num_nodes(n: int) -> int:
    t = tree_shape(n)
    return t["real_nodes"] - t["real_leaves"]


def aggregate(leaves: np.ndarray, node_kind=BBOX, node_fbytes=None, built_level=1, num_threads=1, min_elems=100):
    kind, fbytes, ib, mb = _desc(leaves)
    node_fbytes = node_fbytes or fbytes
    nodes = np.zeros(num_nodes(len(leaves)), volume_dtype(node_kind, node_fbytes))
    _check(lib().orc_aggregate(_p(leaves), len(leaves), kind, fbytes, ib, mb, _p(nodes), node_kind, node_fbytes,
                               built_level, num_threads, min_elems), "aggregate")
    return nodes


def build(leaves: np.ndarray, node_kind=BBOX, node_fbytes=None, built_level=1, compute_extrema=True, mins=None, maxs=None,
          num_threads=1, min_elems=100):
    """BVH(leaves, NodeType; built_level) — leaves are encoded + sorted IN PLACE; returns (nodes, mins, maxs)."""
    kind, fbytes, ib, mb = _desc(leaves)
    node_fbytes = node_fbytes or fbytes
    f = _f(fbytes)
    mn = np.zeros(3, f) if mins is None else np.asarray(mins, f).copy()
    mx = np.zeros(3, f) if maxs is None else np.asarray(maxs, f).copy()
    nodes = np.zeros(num_nodes(len(leaves)), volume_dtype(node_kind, node_fbytes))
    rc = lib().orc_build(_p(leaves), len(leaves), kind, fbytes, ib, mb, _p(nodes), node_kind, node_fbytes, built_level,
                         int(compute_extrema), _p(mn), _p(mx), num_threads, min_elems, min_elems, min_elems)
    _check(rc, "build")
    return nodes, mn, mx


def _node_desc(nodes: np.ndarray):
    return _vdesc(nodes)


def traverse_single(leaves, nodes, built_level=1, start_level=None, num_threads=1, min_elems=100, want_counts=False):
    kind, fbytes, ib, mb = _desc(leaves)
    nk, nf = _node_desc(nodes)
    start_level = start_level or max(1, built_level)
    n = len(leaves)
    counts = np.zeros(n, {4: np.int32, 8: np.int64}[ib]) if want_counts else None
    args = (_p(leaves), n, kind, fbytes, ib, mb, _p(nodes), nk, nf, built_level, start_level)
    total = _check(lib().orc_traverse_single(*args, None, 0, _p(counts), num_threads, min_elems), "traverse_single")
    contacts = np.zeros(total, pair_dtype(ib))
    if total:
        _check(lib().orc_traverse_single(*args, _p(contacts), total, None, num_threads, min_elems), "traverse_single")
    return (contacts, counts) if want_counts else contacts


def traverse_pair(leaves1, nodes1, leaves2, nodes2, built_level1=1, built_level2=1, start_level1=None, start_level2=None,
                  num_threads=1, min_elems=100):
    kind, fbytes, ib, mb = _desc(leaves1)
    assert _desc(leaves2) == (kind, fbytes, ib, mb)
    nk, nf = _node_desc(nodes2 if len(nodes2) or not len(nodes1) else nodes1)
    start_level1 = start_level1 or max(1, built_level1)
    start_level2 = start_level2 or max(1, built_level2)
    args = (_p(leaves1), len(leaves1), _p(nodes1), built_level1, start_level1,
            _p(leaves2), len(leaves2), _p(nodes2), built_level2, start_level2, kind, fbytes, ib, mb, nk, nf)
    total = _check(lib().orc_traverse_pair(*args, None, 0, None, num_threads, min_elems), "traverse_pair")
    contacts = np.zeros(total, pair_dtype(ib))
    if total:
        _check(lib().orc_traverse_pair(*args, _p(contacts), total, None, num_threads, min_elems), "traverse_pair")
    return contacts


def traverse_rays(leaves, nodes, points, directions, built_level=1, start_level=1, num_threads=1, min_elems=100):
    """points/directions: (3, R) arrays as in the reference (column-major 3xR == C-order (R, 3))."""
    kind, fbytes, ib, mb = _desc(leaves)
    nk, nf = _node_desc(nodes)
    points = np.asarray(points)
    directions = np.asarray(directions)
    assert points.shape[0] == 3 and directions.shape == points.shape
    rf = 8 if points.dtype == np.float64 else 4
    p = np.ascontiguousarray(points.T.astype(_f(rf)))
    d = np.ascontiguousarray(directions.T.astype(_f(rf)))
    nr = p.shape[0]
    args = (_p(leaves), len(leaves), kind, fbytes, ib, mb, _p(nodes), nk, nf, built_level, start_level, _p(p), _p(d), rf, nr)
    total = _check(lib().orc_traverse_rays(*args, None, 0, None, num_threads, min_elems), "traverse_rays")
    contacts = np.zeros(total, pair_dtype(ib))
    if total:
        _check(lib().orc_traverse_rays(*args, _p(contacts), total, None, num_threads, min_elems), "traverse_rays")
    return contacts


# ---- BFS traversals (src/traverse/breadth_first, src/raytrace/breadth_first): (list in the reference's CPU order, num_checks)
def bfs_default_start_level(n: int, built_level: int = 1) -> int:
    """default_start_level(bvh, ::BFSTraversal) — breadth_first/breadth_first.jl:4-6."""
    return max(tree_shape(n)["levels"] // 2, built_level)


def _bfs_fetch(total, ib):
    out = np.zeros(total, pair_dtype(ib))
    _check(lib().orc_bfs_fetch(_p(out), out.nbytes), "bfs_fetch")
    return out


def traverse_bfs_single(leaves, nodes, built_level=1, start_level=None, positions=False):
    kind, fbytes, ib, mb = _desc(leaves)
    nk, nf = _node_desc(nodes)
    n = len(leaves)
    start_level = start_level or bfs_default_start_level(n, built_level)
    checks = C.c_int64(0)
    total = _check(lib().orc_traverse_bfs_single(_p(leaves), n, kind, fbytes, ib, mb, _p(nodes), nk, nf, built_level, start_level,
                                                 int(positions), C.byref(checks)), "traverse_bfs_single")
    return _bfs_fetch(total, ib), int(checks.value)


def traverse_bfs_pair(leaves1, nodes1, leaves2, nodes2, built_level1=1, built_level2=1, start_level1=None, start_level2=None,
                      positions=False):
    kind, fbytes, ib, mb = _desc(leaves1)
    assert _desc(leaves2) == (kind, fbytes, ib, mb)
    nk, nf = _node_desc(nodes2 if len(nodes2) or not len(nodes1) else nodes1)
    start_level1 = start_level1 or bfs_default_start_level(len(leaves1), built_level1)
    start_level2 = start_level2 or bfs_default_start_level(len(leaves2), built_level2)
    checks = C.c_int64(0)
    total = _check(lib().orc_traverse_bfs_pair(_p(leaves1), len(leaves1), _p(nodes1), built_level1, start_level1,
                                               _p(leaves2), len(leaves2), _p(nodes2), built_level2, start_level2,
                                               kind, fbytes, ib, mb, nk, nf, int(positions), C.byref(checks)), "traverse_bfs_pair")
    return _bfs_fetch(total, ib), int(checks.value)


def traverse_bfs_rays(leaves, nodes, points, directions, built_level=1, start_level=1, positions=False):
    kind, fbytes, ib, mb = _desc(leaves)
    nk, nf = _node_desc(nodes)
    p = np.ascontiguousarray(np.asarray(points).T.astype(_f(fbytes)))
    d = np.ascontiguousarray(np.asarray(directions).T.astype(_f(fbytes)))
    checks = C.c_int64(0)
    total = _check(lib().orc_traverse_bfs_rays(_p(leaves), len(leaves), kind, fbytes, ib, mb, _p(nodes), nk, nf, built_level,
                                               start_level, _p(p), _p(d), len(p), int(positions), C.byref(checks)), "traverse_bfs_rays")
    return _bfs_fetch(total, ib), int(checks.value)


def brute_single(volumes):
    kind, fb = _vdesc(volumes)
    volumes = np.ascontiguousarray(volumes)
    c = lib().orc_brute_single(_p(volumes), len(volumes), kind, fb, None, 0)
    out = np.zeros((c, 2), np.int64)
    lib().orc_brute_single(_p(volumes), len(volumes), kind, fb, _p(out), c)
    return out


def brute_pair(v1, v2):
    kind, fb = _vdesc(v1)
    v1, v2 = np.ascontiguousarray(v1), np.ascontiguousarray(v2)
    c = lib().orc_brute_pair(_p(v1), len(v1), _p(v2), len(v2), kind, fb, None, 0)
    out = np.zeros((c, 2), np.int64)
    lib().orc_brute_pair(_p(v1), len(v1), _p(v2), len(v2), kind, fb, _p(out), c)
    return out


def brute_rays(volumes, points, directions):
    kind, fb = _vdesc(volumes)
    volumes = np.ascontiguousarray(volumes)
    p = np.ascontiguousarray(np.asarray(points).T.astype(_f(fb)))
    d = np.ascontiguousarray(np.asarray(directions).T.astype(_f(fb)))
    c = lib().orc_brute_rays(_p(volumes), len(volumes), kind, fb, _p(p), _p(d), len(p), None, 0)
    out = np.zeros((c, 2), np.int64)
    lib().orc_brute_rays(_p(volumes), len(volumes), kind, fb, _p(p), _p(d), len(p), _p(out), c)
    return out


def volumes_from_triangles(tris, kind=BSPHERE, fbytes=4) -> np.ndarray:
    """BSphere{T}(p1, p2, p3) / BBox{T}(p1, p2, p3) for an (n, 3, 3) array of triangle vertices."""
    t = np.ascontiguousarray(np.asarray(tris, _f(fbytes)).reshape(-1, 3, 3))
    out = np.zeros(len(t), volume_dtype(kind, fbytes))
    _check(lib().orc_volumes_from_triangles(_p(t), len(t), kind, fbytes, _p(out)), "volumes_from_triangles")
    return out


def spheres(centers, radii, fbytes=4) -> np.ndarray:
    centers = np.asarray(centers, _f(fbytes)).reshape(-1, 3)
    out = np.zeros(len(centers), volume_dtype(BSPHERE, fbytes))
    out["x"] = centers
    out["r"] = np.asarray(radii, _f(fbytes))
    return out


def boxes_of_spheres(sph: np.ndarray) -> np.ndarray:
    """BBox(BSphere) — merge.jl:47-51."""
    fb = sph.dtype["r"].itemsize
    out = np.zeros(len(sph), volume_dtype(BBOX, fb))
    out["lo"] = sph["x"] - sph["r"][:, None]
    out["up"] = sph["x"] + sph["r"][:, None]
    return out
