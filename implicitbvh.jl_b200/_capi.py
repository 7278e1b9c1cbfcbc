"""ctypes binding of libibvh_b200.so — one Python stub per entry point of include/ibvh.h.

This is the in-container stand-in for the `ccall`s of the Julia package extension.
There is no fallback path: if the shared library is missing or does not load, importing the
product API raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IBVH_B200_LIB") or os.path.join(_HERE, "lib", "libibvh_b200.so")

# status codes (include/ibvh.h)
OK, ERR_ARGUMENT, ERR_DOMAIN, ERR_UNSUPPORTED, ERR_CUDA, ERR_CAPACITY, ERR_ALLOC, ERR_PEER, ERR_AGAIN = range(9)
MAX_PEERS = 16
BSPHERE, BBOX = 0, 1
TRAVERSE_ORDERED, TRAVERSE_UNORDERED, TRAVERSE_REFERENCE_SHAPED, TRAVERSE_COUNTS_VALID = 0, 1, 2, 4
TRAVERSE_STATS, TRAVERSE_PACKET, TRAVERSE_WALK, TRAVERSE_DEFER, TRAVERSE_POSITIONS = 8, 16, 32, 64, 128


class Types(C.Structure):
    _fields_ = [("leaf_kind", C.c_int32), ("float_bytes", C.c_int32), ("index_bytes", C.c_int32),
                ("morton_bytes", C.c_int32), ("node_kind", C.c_int32), ("node_float_bytes", C.c_int32)]


class Tree(C.Structure):
    _fields_ = [("levels", C.c_int64), ("real_leaves", C.c_int64), ("real_nodes", C.c_int64),
                ("virtual_leaves", C.c_int64), ("virtual_nodes", C.c_int64)]


class Bvh(C.Structure):
    _fields_ = [("d_leaves", C.c_void_p), ("d_nodes", C.c_void_p), ("n", C.c_int64), ("built_level", C.c_int64),
                ("types", Types), ("build_id", C.c_uint64)]


class Peer(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("buffers", C.c_uint64 * MAX_PEERS), ("multicast", C.c_uint64),
                ("header_bytes", C.c_int64), ("capacity_bytes", C.c_int64), ("epoch", C.c_uint64), ("fused_seq", C.c_uint64),
                ("region_begin", C.c_int64 * (MAX_PEERS + 1))]


class TraverseParams(C.Structure):
    _fields_ = [("start_level", C.c_int64), ("query_begin", C.c_int64), ("query_count", C.c_int64),
                ("flags", C.c_uint32), ("flip", C.c_int32), ("id_base", C.c_int64), ("peer", C.POINTER(Peer))]


# name -> (restype, argtypes); kept as a table so tests can check every header symbol is exported
_i64, _vp, _ci, _dp = C.c_int64, C.c_void_p, C.c_int, C.POINTER(C.c_double)
SIGNATURES = {
    "ibvh_version": (_ci, []),
    "ibvh_status_string": (C.c_char_p, [_ci]),
    "ibvh_last_error": (C.c_char_p, [_vp]),
    "ibvh_tree_shape": (_ci, [_i64, C.POINTER(Tree), C.POINTER(_i64)]),
    "ibvh_memory_index": (_i64, [C.POINTER(Tree), _i64]),
    "ibvh_level_indices": (_ci, [C.POINTER(Tree), _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "ibvh_isvirtual": (_ci, [C.POINTER(Tree), _i64]),
    "ibvh_compute_build_level": (_ci, [_i64, _ci, _i64, C.c_double, C.POINTER(_i64)]),
    "ibvh_leaf_bytes": (_i64, [C.POINTER(Types)]),
    "ibvh_volume_bytes": (_i64, [C.c_int32, C.c_int32]),
    "ibvh_num_nodes": (_i64, [_i64]),
    "ibvh_create": (_ci, [C.POINTER(_vp), _ci]),
    "ibvh_destroy": (_ci, [_vp]),
    "ibvh_workspace_query": (_i64, [C.POINTER(Types), _i64]),
    "ibvh_workspace_bytes": (_i64, [_vp]),
    "ibvh_release_workspace": (_ci, [_vp]),
    "ibvh_wrap": (_ci, [_vp, _vp, _i64, C.POINTER(Types), _vp, _vp]),
    "ibvh_volumes_from_triangles": (_ci, [_vp, _vp, _i64, C.c_int32, C.c_int32, _vp, _vp]),
    "ibvh_morton_encode": (_ci, [_vp, _vp, _i64, C.POINTER(Types), _ci, _dp, _dp, _dp, _dp, _vp]),
    "ibvh_sort_leaves": (_ci, [_vp, _vp, _i64, C.POINTER(Types), _vp]),
    "ibvh_aggregate": (_ci, [_vp, _vp, _i64, C.POINTER(Types), _vp, _i64, _vp]),
    "ibvh_build": (_ci, [_vp, _vp, _vp, _i64, C.POINTER(Types), _vp, _i64, _ci, _dp, _dp, _vp]),
    "ibvh_build_reference_shaped": (_ci, [_vp, _vp, _vp, _i64, C.POINTER(Types), _vp, _i64, _vp]),
    "ibvh_traverse_single": (_ci, [_vp, C.POINTER(Bvh), C.POINTER(TraverseParams), _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "ibvh_traverse_pair": (_ci, [_vp, C.POINTER(Bvh), C.POINTER(Bvh), C.POINTER(TraverseParams), _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "ibvh_traverse_rays": (_ci, [_vp, C.POINTER(Bvh), _vp, _vp, _i64, C.POINTER(TraverseParams), _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "ibvh_bfs_default_start_level": (_i64, [_i64, _i64]),
    "ibvh_traverse_bfs_single": (_ci, [_vp, C.POINTER(Bvh), C.POINTER(TraverseParams), _vp, _i64, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "ibvh_traverse_bfs_pair": (_ci, [_vp, C.POINTER(Bvh), C.POINTER(Bvh), _i64, _i64, C.c_uint32, _vp, _i64, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "ibvh_traverse_bfs_rays": (_ci, [_vp, C.POINTER(Bvh), _vp, _vp, _i64, C.POINTER(TraverseParams), _vp, _i64, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "ibvh_sort_contacts": (_ci, [_vp, _vp, _i64, C.c_int32, _ci, C.POINTER(_i64), _vp]),
    "ibvh_profile_enable": (_ci, [_vp, _ci]),
    "ibvh_profile_count": (_ci, [_vp]),
    "ibvh_profile_get": (_ci, [_vp, _ci, C.c_char_p, _ci, C.POINTER(C.c_float)]),
    "ibvh_profile_reset": (_ci, [_vp]),
    "ibvh_last_traversal_stats": (_ci, [_vp, C.POINTER(_i64)]),
    "ibvh_traverse_finish": (_ci, [_vp, C.POINTER(_i64)]),
    "ibvh_traverse_cancel": (_ci, [_vp]),
    "ibvh_last_build_id": (C.c_uint64, [_vp]),
    "ibvh_peer_last_counts": (_ci, [_vp, C.POINTER(_i64), C.c_int32]),
    "ibvh_peer_compact_plan": (_ci, [C.c_int32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.c_int32]),
    "ibvh_allgather_pairs": (_ci, [_vp, C.POINTER(Peer), _vp, _i64, C.c_int32, C.POINTER(_i64), C.POINTER(_i64), _vp]),
}

_lib = None


class LibraryMissing(ImportError):
    pass


def lib() -> C.CDLL:
    """Load libibvh_b200.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)        # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def status_string(code: int) -> str:
    return lib().ibvh_status_string(code).decode()
