"""ibvh-b200 — B200-native (sm_100a) hot path of ImplicitBVH.jl behind a C ABI.

The package directory is named `implicitbvh.jl_b200` (not importable by that dotted name); import it
through the repo-root shim: `import ibvh_b200`.
"""
from . import _capi as capi
from .api import (ArgumentError, BBox, BBOX, BFSTraversal, BSphere, BSPHERE, BVH, BVHOptions, BVHTraversal, CudaError,
                  DefaultMortonAlgorithm, DeviceArray, DomainError, ImplicitTree, LVTTraversal, VolumeType,
                  aggregate, bboxes, bspheres, default_start_level, get_handle, isvirtual, leaf_dtype,
                  level_indices, memory_index, morton_encode, pair_dtype, sort_contacts, sort_leaves, traverse,
                  traverse_rays, volumes_from_triangles, wrap_bounding_volumes)

__all__ = [
    "BVH", "BVHTraversal", "BVHOptions", "traverse", "traverse_rays", "sort_contacts", "default_start_level",
    "ImplicitTree", "memory_index", "level_indices", "isvirtual", "DefaultMortonAlgorithm", "LVTTraversal", "BFSTraversal",
    "BSphere", "BBox", "DeviceArray", "ArgumentError", "DomainError", "CudaError",
]
