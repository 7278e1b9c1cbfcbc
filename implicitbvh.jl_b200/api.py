"""Host-side mirror of the ImplicitBVH.jl public API for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference (src/ImplicitBVH.jl:20-24 exports):
`BVH`, `BVHOptions`, `BVHTraversal`, `traverse`, `traverse_rays`, `default_start_level`,
`ImplicitTree`, `memory_index`, `level_indices`, `isvirtual`, `DefaultMortonAlgorithm`,
`LVTTraversal`. Julia is not available in this image, so this Python layer plays the role of the
Julia methods that would `ccall` libibvh_b200.so; torch is used only for device
memory and streams. Nothing here computes on the CPU: every geometric result comes from the CUDA
kernels, and the module refuses to work without the shared library and a CUDA device.

Device arrays of isbits structs are `DeviceArray`s: a torch.uint8 tensor of raw bytes plus the numpy
structured dtype that describes one element (the stand-in for `CuVector{BoundingVolume{...}}`).
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field
from typing import Optional, Union

import numpy as np
import torch

from . import _capi as capi
from ._capi import BBOX, BSPHERE


# ---------------------------------------------------------------------------------------------
# errors (the reference throws ArgumentError / DomainError; both are ValueErrors here)
# ---------------------------------------------------------------------------------------------
class ArgumentError(ValueError):
    pass


class DomainError(ValueError):
    pass


class CudaError(RuntimeError):
    pass


def _raise(code: int, handle=None, what: str = ""):
    msg = f"{what}: {capi.status_string(code)}"
    if code == capi.ERR_ARGUMENT:
        raise ArgumentError(msg)
    if code == capi.ERR_DOMAIN:
        raise DomainError(msg)
    if code in (capi.ERR_CUDA, capi.ERR_ALLOC) and handle is not None:
        msg += " — " + capi.lib().ibvh_last_error(handle).decode()
    if code == capi.ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CudaError(msg)


# ---------------------------------------------------------------------------------------------
# isbits layouts (SURVEY.md §8 layout table)
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class VolumeType:
    """`BSphere{T}` / `BBox{T}` as a type object (bsphere.jl:26-29, bbox.jl:35-38)."""
    kind: int
    float_bytes: int = 4

    @property
    def dtype(self) -> np.dtype:
        f = {4: np.float32, 8: np.float64}[self.float_bytes]
        if self.kind == BSPHERE:
            return np.dtype([("x", f, 3), ("r", f)])
        return np.dtype([("lo", f, 3), ("up", f, 3)])

    def __repr__(self):
        return f"{'BSphere' if self.kind == BSPHERE else 'BBox'}{{Float{8 * self.float_bytes}}}"


def BSphere(T=np.float32) -> VolumeType:
    return VolumeType(BSPHERE, np.dtype(T).itemsize)


def BBox(T=np.float32) -> VolumeType:
    return VolumeType(BBOX, np.dtype(T).itemsize)


def leaf_dtype(vol: VolumeType, index=np.int32, morton=np.uint32) -> np.dtype:
    """BoundingVolume{V, I, M} (bounding_volumes.jl:55-59), natural alignment."""
    return np.dtype([("volume", vol.dtype), ("index", np.dtype(index)), ("morton", np.dtype(morton))], align=True)


def pair_dtype(index=np.int32) -> np.dtype:
    """IndexPair{I} (traverse/traverse.jl:6)."""
    return np.dtype([("a", np.dtype(index)), ("b", np.dtype(index))])


def _volume_of_dtype(dt: np.dtype) -> Optional[VolumeType]:
    if dt.names == ("x", "r"):
        return VolumeType(BSPHERE, dt["r"].itemsize)
    if dt.names == ("lo", "up"):
        return VolumeType(BBOX, dt["lo"].base.itemsize)
    return None


def bspheres(centers, radii, T=np.float32) -> np.ndarray:
    """Convenience: array of BSphere{T} from (n,3) centres and (n,) radii."""
    centers = np.asarray(centers, T).reshape(-1, 3)
    out = np.zeros(len(centers), BSphere(T).dtype)
    out["x"] = centers
    out["r"] = np.asarray(radii, T)
    return out


def bboxes(lo, up, T=np.float32) -> np.ndarray:
    lo = np.asarray(lo, T).reshape(-1, 3)
    out = np.zeros(len(lo), BBox(T).dtype)
    out["lo"] = lo
    out["up"] = np.asarray(up, T).reshape(-1, 3)
    return out


# ---------------------------------------------------------------------------------------------
# device arrays
# ---------------------------------------------------------------------------------------------
class DeviceArray:
    """A vector of isbits structs in device (or pinned host) memory: raw bytes + element dtype."""

    def __init__(self, tensor: torch.Tensor, dtype: np.dtype):
        assert tensor.dtype == torch.uint8 and tensor.dim() == 1 and tensor.is_contiguous()
        self.dtype = np.dtype(dtype)
        assert tensor.numel() % self.dtype.itemsize == 0
        self.tensor = tensor

    @classmethod
    def empty(cls, n: int, dtype, device) -> "DeviceArray":
        dtype = np.dtype(dtype)
        return cls(torch.empty(int(n) * dtype.itemsize, dtype=torch.uint8, device=device), dtype)

    @classmethod
    def from_numpy(cls, arr: np.ndarray, device=None, pin: bool = False) -> "DeviceArray":
        arr = np.ascontiguousarray(arr)
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1))
        if pin:
            t = t.pin_memory()
        if device is not None:
            t = t.to(device, non_blocking=True)
        return cls(t, arr.dtype)

    def to(self, device) -> "DeviceArray":
        return DeviceArray(self.tensor.to(device, non_blocking=True), self.dtype)

    def numpy(self) -> np.ndarray:
        return self.tensor.cpu().numpy().view(self.dtype)

    def __len__(self):
        return self.tensor.numel() // self.dtype.itemsize

    def __getitem__(self, s: slice) -> "DeviceArray":
        assert isinstance(s, slice) and s.step in (None, 1)
        a, b, _ = s.indices(len(self))
        it = self.dtype.itemsize
        return DeviceArray(self.tensor[a * it:max(a, b) * it], self.dtype)

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    @property
    def device(self):
        return self.tensor.device

    @property
    def is_cuda(self):
        return self.tensor.is_cuda

    def __repr__(self):
        return f"DeviceArray({len(self)} x {self.dtype.itemsize} B on {self.device})"


# ---------------------------------------------------------------------------------------------
# handles (one per device)
# ---------------------------------------------------------------------------------------------
_handles = {}


def _device_index(device) -> int:
    if not torch.cuda.is_available():
        raise CudaError("no CUDA device available; ibvh-b200 has no CPU fallback")
    d = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if d.type != "cuda":
        raise CudaError("ibvh-b200 runs on CUDA devices only; there is no CPU fallback")
    return d.index if d.index is not None else torch.cuda.current_device()


def get_handle(device=None):
    if not torch.cuda.is_available():
        raise CudaError("no CUDA device available; ibvh-b200 has no CPU fallback")
    idx = _device_index(device)
    h = _handles.get(idx)
    if h is None:
        out = C.c_void_p()
        rc = capi.lib().ibvh_create(C.byref(out), idx)
        if rc != capi.OK:
            _raise(rc, None, "ibvh_create")
        h = out
        _handles[idx] = h
    return h


def _stream_ptr(device_index: int) -> int:
    return torch.cuda.current_stream(device_index).cuda_stream


# At most one deferred traversal (traverse(..., defer=True)) is outstanding per device handle. The registry lets
# the calls that would invalidate it — another traversal on the handle, a build that overwrites buffers of the BVH
# it still reads — resolve it first, and lets a dropped result cancel itself.
_outstanding = {}                # device index -> weakref to the pending BVHTraversal


def _pending_on(device_index: int):
    ref = _outstanding.get(device_index)
    tr = ref() if ref is not None else None
    if tr is None or tr._resolve is None:
        _outstanding.pop(device_index, None)
        return None
    return tr


def _resolve_outstanding(device_index: int):
    """Finish the deferred traversal outstanding on this device, if any (blocks until its count is known)."""
    tr = _pending_on(device_index)
    if tr is not None:
        tr.num_contacts                     # noqa: B018 — property access resolves it
    _outstanding.pop(device_index, None)


def _overlaps(a: "DeviceArray", b: "DeviceArray") -> bool:
    if a is None or b is None or a.tensor.numel() == 0 or b.tensor.numel() == 0 or a.device != b.device:
        return False
    a0, b0 = a.ptr, b.ptr
    return a0 < b0 + b.tensor.numel() and b0 < a0 + a.tensor.numel()


# ---------------------------------------------------------------------------------------------
# options (utils.jl:34-93), Morton algorithm (morton/default.jl:22-42), traversal algorithm tag
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class DefaultMortonAlgorithm:
    exemplar: type = np.uint32
    compute_extrema: bool = True
    mins: tuple = (float("nan"),) * 3
    maxs: tuple = (float("nan"),) * 3

    def __post_init__(self):
        if np.dtype(self.exemplar) not in (np.dtype(np.uint16), np.dtype(np.uint32), np.dtype(np.uint64)):
            raise ArgumentError("Morton type must be UInt16, UInt32 or UInt64")

    @property
    def eltype(self) -> np.dtype:
        return np.dtype(self.exemplar)


@dataclass(frozen=True)
class BVHOptions:
    index: type = np.int32
    morton: DefaultMortonAlgorithm = field(default_factory=DefaultMortonAlgorithm)
    num_threads: int = 1
    min_mortons_per_thread: int = 100
    min_sorts_per_thread: int = 100
    min_boundings_per_thread: int = 100
    min_traversals_per_thread: int = 100
    block_size: int = 256

    def __post_init__(self):
        for name in ("num_threads", "min_mortons_per_thread", "min_sorts_per_thread", "min_boundings_per_thread",
                     "min_traversals_per_thread", "block_size"):
            if not getattr(self, name) > 0:                       # utils.jl:74-79
                raise ArgumentError(f"{name} > 0 must hold")
        if np.dtype(self.index) not in (np.dtype(np.int32), np.dtype(np.int64)):
            raise NotImplementedError("index type must be Int32 or Int64 in this build")

    @property
    def index_dtype(self) -> np.dtype:
        return np.dtype(self.index)


class LVTTraversal:
    """Leaf-vs-tree traversal (traverse/leaf_vs_tree/leaf_vs_tree.jl:1)."""

    def __repr__(self):
        return "LVTTraversal()"


class BFSTraversal:
    """Simultaneous breadth-first traversal, nodes paired level by level (traverse/breadth_first/breadth_first.jl:1;
    traverse/traverse.jl:19-24): fewest checks, BVTT lists of 10-20x the contacts in library scratch. Contacts come back
    as a set in unspecified order, like the reference's GPU backend; `num_checks` is exact."""

    def __repr__(self):
        return "BFSTraversal()"


# ---------------------------------------------------------------------------------------------
# implicit tree (implicit_tree.jl) — host-only integer math done by the C library
# ---------------------------------------------------------------------------------------------
class ImplicitTree:
    def __init__(self, num_leaves: int):
        t = capi.Tree()
        rc = capi.lib().ibvh_tree_shape(int(num_leaves), C.byref(t), None)
        if rc == capi.ERR_DOMAIN:
            raise DomainError(f"{num_leaves}: must have at least one geometry!")      # implicit_tree.jl:78-80
        self._c = t
        self.levels, self.real_leaves, self.real_nodes = t.levels, t.real_leaves, t.real_nodes
        self.virtual_leaves, self.virtual_nodes = t.virtual_leaves, t.virtual_nodes

    def skips(self) -> np.ndarray:
        t = capi.Tree()
        s = (C.c_int64 * 64)()
        capi.lib().ibvh_tree_shape(self.real_leaves, C.byref(t), s)
        return np.array(s[: self.levels], np.int64)

    def __repr__(self):
        return f"ImplicitTree(levels: {self.levels}, real_leaves: {self.real_leaves})"


def memory_index(tree: ImplicitTree, implicit_index: int) -> int:
    if not (1 <= implicit_index <= 2 ** tree.levels - 1):
        raise IndexError(implicit_index)
    return int(capi.lib().ibvh_memory_index(C.byref(tree._c), int(implicit_index)))


def level_indices(tree: ImplicitTree, level: int):
    if not (1 <= level <= tree.levels):
        raise IndexError(level)
    a, b = C.c_int64(), C.c_int64()
    capi.lib().ibvh_level_indices(C.byref(tree._c), int(level), C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def isvirtual(tree: ImplicitTree, implicit_index: int) -> bool:
    if not (1 <= implicit_index <= 2 ** tree.levels - 1):
        raise IndexError(implicit_index)
    return bool(capi.lib().ibvh_isvirtual(C.byref(tree._c), int(implicit_index)))


# ---------------------------------------------------------------------------------------------
# BVH (build.jl:155-271)
# ---------------------------------------------------------------------------------------------
def _types(leaf_vol: VolumeType, node: VolumeType, index: np.dtype, morton: np.dtype) -> capi.Types:
    nfb = 0 if node.float_bytes == leaf_vol.float_bytes else node.float_bytes
    return capi.Types(leaf_vol.kind, leaf_vol.float_bytes, index.itemsize, morton.itemsize, node.kind, nfb)


class BVH:
    """`BVH(bounding_volumes, node_type=BBox{Float32}; built_level=1, cache=nothing, options=BVHOptions())`.

    `reference_shaped=True` (extension, measurement aid): build with the naive reference-shaped kernels instead.
    `bounding_volumes`: numpy structured array or DeviceArray of raw volumes (BSphere/BBox: wrapped on
    device, index = position) or of BoundingVolume structs (sorted IN PLACE when already on the device,
    as in the reference, build.jl:147-152). Fields mirror build.jl:155-166.
    """

    def __init__(self, bounding_volumes: Union[np.ndarray, DeviceArray], node_type: VolumeType = None, *,
                 built_level: Union[int, float] = 1, cache: Optional["BVH"] = None, options: BVHOptions = None,
                 device=None, reference_shaped: bool = False):
        options = options or BVHOptions()
        node_type = node_type or BBox(np.float32)
        lib = capi.lib()
        if isinstance(bounding_volumes, np.ndarray):
            dev = torch.device("cuda", _device_index(device))
            src = DeviceArray.from_numpy(bounding_volumes, device=dev)
        else:
            src = bounding_volumes
            if not src.is_cuda:
                src = src.to(torch.device("cuda", _device_index(device)))
        didx = src.device.index
        self._handle = get_handle(src.device)
        I, M = options.index_dtype, options.morton.eltype
        n = len(src)

        vol = _volume_of_dtype(src.dtype)
        wrapped_input = vol is None
        if wrapped_input:
            if src.dtype.names != ("volume", "index", "morton"):
                raise ArgumentError("bounding_volumes must be BSphere / BBox volumes or BoundingVolume structs")
            vol = _volume_of_dtype(src.dtype["volume"])
            # check_bounding_volume_types, build.jl:355-361
            if src.dtype["index"] != I:
                raise ArgumentError(f"BoundingVolume index type {src.dtype['index']} does not match BVHOptions index_exemplar type {I}")
            if src.dtype["morton"] != M:
                raise ArgumentError(f"BoundingVolume morton type {src.dtype['morton']} does not match BVHOptions morton type {M}")
        if node_type.float_bytes != vol.float_bytes and not (vol.float_bytes == 8 and node_type.float_bytes == 4):
            raise NotImplementedError("node float types built: the leaf float type, or Float32 nodes over Float64 leaves (the reference's default)")
        ldt = leaf_dtype(vol, I, M)
        self.types = _types(vol, node_type, I, M)
        assert ldt.itemsize == lib.ibvh_leaf_bytes(C.byref(self.types))

        if n < 1:
            raise DomainError(f"{n}: must have at least one geometry!")
        self.tree = ImplicitTree(n)

        # skips: reuse from cache (build.jl:232-239)
        if cache is not None and cache.index_dtype != I:
            raise ArgumentError("eltype(cache.skips) === I must hold")
        if cache is not None and cache.tree.real_leaves == n and cache.skips.device == src.device:
            self.skips = cache.skips                      # same n => same skips: reused, as the reference does
        else:
            self.skips = torch.from_numpy(self.tree.skips().astype(I)).to(src.device, non_blocking=True)

        # built level (build.jl:309-325)
        out = C.c_int64()
        if isinstance(built_level, (int, np.integer)) and not isinstance(built_level, bool):
            rc = lib.ibvh_compute_build_level(self.tree.levels, 0, int(built_level), 0.0, C.byref(out))
        elif isinstance(built_level, (float, np.floating)):
            rc = lib.ibvh_compute_build_level(self.tree.levels, 1, 0, float(built_level), C.byref(out))
        else:
            raise TypeError("built_level (the level to build BVH up to) must be Integer or AbstractFloat")
        if rc != capi.OK:
            _raise(rc, self._handle, "compute_build_level")
        self.built_level = int(out.value)

        # A deferred traversal may still be outstanding on this device. If it has to be repeated when it is resolved
        # (IBVH_ERR_AGAIN / IBVH_ERR_CAPACITY) it reads its BVH again, so that BVH's buffers must not be overwritten by
        # this build: an in-place build over its leaves resolves the traversal first, and its node buffer is not
        # taken over from `cache` (a fresh one is allocated; the caching allocator ping-pongs two buffers).
        pending = _pending_on(didx)
        pending_bvhs = pending._bvhs if pending is not None else ()
        if wrapped_input and any(_overlaps(src, b.leaves) for b in pending_bvhs):
            _resolve_outstanding(didx)
            pending_bvhs = ()

        # nodes: reuse from cache when the type matches (build.jl:256-263)
        num_nodes = self.tree.real_nodes - self.tree.real_leaves
        self.node_type = node_type
        if cache is not None:
            if cache.nodes.dtype != node_type.dtype:
                raise ArgumentError("eltype(cache.nodes) === N must hold")
            nodes_busy = any(_overlaps(cache.nodes, b.nodes) for b in pending_bvhs)
            if len(cache.nodes) == num_nodes and cache.nodes.device == src.device and not nodes_busy:
                self.nodes = cache.nodes
            else:
                self.nodes = DeviceArray.empty(num_nodes, node_type.dtype, src.device)
        else:
            self.nodes = DeviceArray.empty(num_nodes, node_type.dtype, src.device)

        if wrapped_input:
            self.leaves = src
            d_vol = None
        else:
            self.leaves = DeviceArray.empty(n, ldt, src.device)
            d_vol = src.ptr
        self._volumes_keepalive = src

        m = options.morton
        mins = (C.c_double * 3)(*[float(x) for x in m.mins])
        maxs = (C.c_double * 3)(*[float(x) for x in m.maxs])
        if reference_shaped:
            # measurement aid: the naive "reference-shaped" GPU build (proxy of the reference's CUDA.jl backend; same result)
            if d_vol is None or not m.compute_extrema:
                raise ArgumentError("the reference-shaped proxy build takes raw volumes and computes the extrema itself")
            rc = lib.ibvh_build_reference_shaped(self._handle, d_vol, self.leaves.ptr, n, C.byref(self.types),
                                                 self.nodes.ptr if num_nodes > 0 else None, self.built_level, _stream_ptr(didx))
            if rc != capi.OK:
                _raise(rc, self._handle, "ibvh_build_reference_shaped")
            return
        # (the library switches to the handle's device itself; the stream is torch's current stream on it)
        rc = lib.ibvh_build(self._handle, d_vol, self.leaves.ptr, n, C.byref(self.types),
                            self.nodes.ptr if num_nodes > 0 else None, self.built_level,
                            1 if m.compute_extrema else 0, mins, maxs, _stream_ptr(didx))
        if rc != capi.OK:
            _raise(rc, self._handle, "ibvh_build")
        self._build_id = int(lib.ibvh_last_build_id(self._handle))     # the build's sidecar (0 = none)

    # -- helpers ---------------------------------------------------------------------------------
    def _c_bvh(self) -> capi.Bvh:
        return capi.Bvh(self.leaves.ptr, self.nodes.ptr if len(self.nodes) else None, len(self.leaves), self.built_level, self.types,
                        getattr(self, "_build_id", 0))

    @property
    def index_dtype(self) -> np.dtype:
        return self.leaves.dtype["index"]

    def __repr__(self):
        return (f"BVH\n  built_level: {self.built_level}\n  tree:        {self.tree}\n  skips:       ({len(self.skips)},)\n"
                f"  nodes:       {self.nodes}\n  leaves:      {self.leaves}\n")


# ---------------------------------------------------------------------------------------------
# traversal results (traverse/traverse.jl:54-108)
# ---------------------------------------------------------------------------------------------
class BVHTraversal:
    def __init__(self, start_level1: int, start_level2: int, num_checks: int, num_contacts: int,
                 cache1: DeviceArray, cache2: DeviceArray):
        self.start_level1, self.start_level2 = int(start_level1), int(start_level2)
        self.num_checks = int(num_checks)
        self._num_contacts = int(num_contacts)
        self._resolve = None                  # deferred traversal (traverse(..., defer=True)): resolves on first use
        self._cancel = None                   # ... or is cancelled when the result is dropped unread
        self._bvhs = ()                       # the BVHs a pending deferred traversal still needs intact
        self.cache1, self.cache2 = cache1, cache2

    def __del__(self):
        cancel = getattr(self, "_cancel", None)
        if cancel is not None and getattr(self, "_resolve", None) is not None:
            try:
                cancel()
            except Exception:
                pass

    @property
    def num_contacts(self) -> int:
        if self._resolve is not None:
            resolve, self._resolve = self._resolve, None
            self._cancel = None
            try:
                resolve(self)
            finally:
                self._bvhs = ()
        return self._num_contacts

    @num_contacts.setter
    def num_contacts(self, v):
        self._num_contacts = int(v)

    @property
    def contacts(self) -> DeviceArray:
        """view(cache1, 1:num_contacts) — traverse/traverse.jl:98-104."""
        return self.cache1[: self.num_contacts]

    def __repr__(self):
        return (f"BVHTraversal\n  start_level1: {self.start_level1}\n  start_level2: {self.start_level2}\n"
                f"  num_checks:   {self.num_checks}\n  num_contacts: {self.num_contacts}\n"
                f"  cache1:       {self.cache1}\n  cache2:       {self.cache2}\n")


def default_start_level(bvh: BVH, alg=None) -> int:
    """leaf_vs_tree/leaf_vs_tree.jl:4-6; breadth_first/breadth_first.jl:4-6."""
    if isinstance(alg, BFSTraversal):
        return int(capi.lib().ibvh_bfs_default_start_level(bvh.tree.levels, bvh.built_level))
    if alg is not None and not isinstance(alg, LVTTraversal):
        raise ArgumentError(f"default_start_level not implemented for: {alg}")
    return max(1, bvh.built_level)


class _Pending:
    """A deferred traversal (IBVH_TRAVERSE_DEFER): resolved by ibvh_traverse_finish when num_contacts is first read;
    if the library asks for a repeat (scratch lists / contacts buffer too small) the synchronous call is made."""

    def __init__(self, handle):
        self.handle = handle

    def attach(self, tr: "BVHTraversal", rerun, bvhs=(), device_index: int = 0):
        handle = self.handle
        tr._bvhs = tuple(bvhs)
        tr._cancel = lambda: capi.lib().ibvh_traverse_cancel(handle)
        _outstanding[device_index] = weakref.ref(tr)

        def resolve(t: "BVHTraversal"):
            total = C.c_int64(0)
            rc = capi.lib().ibvh_traverse_finish(handle, C.byref(total))
            if rc == capi.OK:
                t._num_contacts = int(total.value)
                _finish_stats(t, handle)
                return
            if rc in (capi.ERR_AGAIN, capi.ERR_CAPACITY):
                again = rerun()
                t._num_contacts, t.cache1, t.cache2 = again.num_contacts, again.cache1, again.cache2
                return
            _raise(rc, handle, "ibvh_traverse_finish")

        tr._resolve = resolve
        return tr


def _finish_stats(tr: "BVHTraversal", handle) -> "BVHTraversal":
    """BVHTraversal.num_checks (traverse/traverse.jl:48,58): the reference's LVT leaves it at 0; here it carries the box-box +
    leaf-leaf tests of the traversal when the schedule that ran counts them (the pyramid schedule always does)."""
    st = (C.c_int64 * 4)()
    capi.lib().ibvh_last_traversal_stats(handle, st)
    if st[3] > 0:
        tr.num_checks = int(st[0]) + int(st[1])
    return tr


def _fused_call(call, peer, handle, what: str) -> C.c_int64:
    """One fused traversal + all-gather (collective). The ranks' regions of the gathered list are sized from the previous
    call's per-rank counts; if a region turns out too small (IBVH_ERR_CAPACITY: the same verdict on every rank, with the
    counts) they are re-sized from the counts just seen and the call is repeated once."""
    total = C.c_int64(0)
    rc = call(peer.next_fused(), total)
    if rc == capi.ERR_CAPACITY:
        counts = peer.last_counts()
        if not (peer.set_regions(counts) or peer.set_regions(counts, slack=0.01) or peer.set_regions(counts, slack=0.0, pad=2)):
            _raise(rc, handle, f"{what}: the peer list area is too small for {total.value} pairs")
        rc = call(peer.next_fused(), total)
    if rc != capi.OK:
        _raise(rc, handle, f"{what} (gathered total {total.value} pairs)")
    counts = peer.last_counts()
    if not (peer.set_regions(counts) or peer.set_regions(counts, slack=0.01) or peer.set_regions(counts, slack=0.0, pad=2)):
        peer.set_regions(None)
    return total


def _eval_narrow(narrow, *cols) -> np.ndarray:
    """Evaluate a user `narrow` predicate over candidate contacts. `cols` are equally long numpy arrays (BoundingVolume
    records; for rays also (k, 3) points / directions). The predicate is tried vectorised first (called once with the
    whole arrays, must return k booleans) and element by element otherwise, like the reference calls it."""
    k = len(cols[0])
    if k == 0:
        return np.zeros(0, bool)
    try:
        m = np.asarray(narrow(*cols))
        if m.shape == (k,) and m.dtype == np.bool_:
            return m
    except Exception:
        pass
    return np.fromiter((bool(narrow(*(c[i] for c in cols))) for i in range(k)), bool, k)


def _gather_records(arr: "DeviceArray", pos0: torch.Tensor) -> np.ndarray:
    """arr[pos0] (0-based positions on the device) as a host numpy structured array."""
    it = arr.dtype.itemsize
    rows = arr.tensor.view(len(arr), it).index_select(0, pos0.to(torch.int64))
    return rows.cpu().numpy().reshape(-1).view(arr.dtype)


def _narrow_pairs(narrow, tr: "BVHTraversal", leaves1: "DeviceArray", leaves2: "DeviceArray", single: bool, I: np.dtype,
                  nqueries: int, query_col: int, q_begin: int) -> "BVHTraversal":
    """Post-filter of a traversal run with IBVH_TRAVERSE_POSITIONS: the reference evaluates `narrow(bv1, bv2)` only after a
    positive leaf test (traverse_single.jl:170, traverse_pair.jl:206), so filtering the positive pairs with the same
    predicate over leaves[position] gives the identical list, order included. Here the predicate is a
    host Python callable, so the candidates travel to the host; a Julia caller would broadcast its closure on the device."""
    device = leaves1.device
    tI = torch.int32 if I.itemsize == 4 else torch.int64
    pos = tr.contacts.tensor.view(tI).reshape(-1, 2)
    b1 = _gather_records(leaves1, pos[:, 0] - 1)
    b2 = _gather_records(leaves2, pos[:, 1] - 1)
    keep = _eval_narrow(narrow, b1, b2)
    i1, i2 = b1["index"][keep], b2["index"][keep]
    out = np.zeros(int(keep.sum()), pair_dtype(I))
    if single:
        out["a"], out["b"] = np.minimum(i1, i2), np.maximum(i1, i2)
    else:
        out["a"], out["b"] = i1, i2
    c1 = DeviceArray.from_numpy(out, device=device)
    # cache2 = inclusive scan of the per-query counts that survive the predicate (the reference's thread_ncontacts)
    q = (pos[:, query_col].to(torch.int64) - 1 - q_begin)[torch.from_numpy(keep).to(device)]
    counts = torch.bincount(q, minlength=nqueries)[:nqueries].cumsum(0).to(tI)
    c2 = DeviceArray(counts.view(torch.uint8).reshape(-1).contiguous(), I)
    return BVHTraversal(tr.start_level1, tr.start_level2, tr.num_checks, len(out), c1, c2)


def _run_two_phase(call, handle, device, I: np.dtype, nqueries: int, cache: Optional[BVHTraversal], ordered: bool,
                   reference_shaped: bool, packet: bool = False, walk: bool = False, extra_flags: int = 0):
    """The reference's count -> accumulate -> allocate/grow -> write protocol (traverse_single.jl:23-78),
    with `cache1`/`cache2` reused and grown only when too small."""
    pdt = pair_dtype(I)
    flags = capi.TRAVERSE_ORDERED if ordered else capi.TRAVERSE_UNORDERED
    sched = (capi.TRAVERSE_REFERENCE_SHAPED if reference_shaped else 0) | (capi.TRAVERSE_PACKET if packet else 0) | \
            (capi.TRAVERSE_WALK if walk else 0) | extra_flags
    flags |= sched
    if cache is not None:
        if cache.cache2.dtype != I:
            raise ArgumentError("eltype(cache.cache2) === I must hold")
        if cache.cache1.dtype != pdt:
            raise ArgumentError("eltype(cache.cache1) === IndexPair{I} must hold")
        cache2 = cache.cache2 if (len(cache.cache2) >= nqueries and cache.cache2.device == device) else DeviceArray.empty(nqueries, I, device)
        cache1 = cache.cache1 if cache.cache1.device == device else DeviceArray.empty(0, pdt, device)
    else:
        cache2 = DeviceArray.empty(nqueries, I, device)
        cache1 = None
    total = C.c_int64(0)
    if cache1 is not None and len(cache1) > 0:
        rc = call(flags, cache2.ptr, cache1.ptr, len(cache1), total)
        if rc == capi.OK:
            return int(total.value), cache1, cache2
        if rc != capi.ERR_CAPACITY:
            _raise(rc, handle, "traverse")
        need = int(total.value)
        cache1 = DeviceArray.empty(need, pdt, device)
        second = (flags | capi.TRAVERSE_COUNTS_VALID) if ordered else flags
        rc = call(second, cache2.ptr, cache1.ptr, need, total)
        if rc != capi.OK:
            _raise(rc, handle, "traverse")
        return int(total.value), cache1, cache2
    # no usable contacts buffer yet: count pass, allocate exactly, write pass
    rc = call(capi.TRAVERSE_ORDERED | sched, cache2.ptr, None, 0, total)
    if rc != capi.OK:
        _raise(rc, handle, "traverse (count)")
    need = int(total.value)
    cache1 = DeviceArray.empty(need, pdt, device)
    if need == 0:
        return 0, cache1, cache2
    second = (flags | capi.TRAVERSE_COUNTS_VALID) if ordered else flags
    rc = call(second, cache2.ptr, cache1.ptr, need, total)
    if rc != capi.OK:
        _raise(rc, handle, "traverse (write)")
    return int(total.value), cache1, cache2


def _bfs_protocol(call, handle, device, I: np.dtype, cache: Optional[BVHTraversal], extra_flags: int = 0):
    """BFS output protocol: cache1 / cache2 are IndexPair vectors as in the reference (traverse_single.jl:88-101: both are
    BVTT buffers there; here the BVTT lives in library scratch and cache1 only receives contacts). A too-small cache1 is
    grown to the exact need and only the leaf level is repeated (IBVH_TRAVERSE_COUNTS_VALID)."""
    pdt = pair_dtype(I)
    if cache is not None:
        if cache.cache1.dtype != pdt:
            raise ArgumentError("eltype(cache.cache1) === IndexPair{I} must hold")
        if cache.cache2.dtype != pdt:
            raise ArgumentError("eltype(cache.cache2) === IndexPair{I} must hold")
    cache1 = cache.cache1 if (cache is not None and cache.cache1.device == device) else None
    cache2 = cache.cache2 if (cache is not None and cache.cache2.device == device) else DeviceArray.empty(0, pdt, device)
    total, checks = C.c_int64(0), C.c_int64(0)
    if cache1 is not None and len(cache1) > 0:
        rc = call(extra_flags, cache1.ptr, len(cache1), total, checks)
        if rc == capi.OK:
            return int(total.value), int(checks.value), cache1, cache2
        if rc != capi.ERR_CAPACITY:
            _raise(rc, handle, "traverse (BFS)")
    else:
        rc = call(extra_flags, None, 0, total, checks)
        if rc != capi.OK:
            _raise(rc, handle, "traverse (BFS, count)")
    need = int(total.value)
    cache1 = DeviceArray.empty(need, pdt, device)
    if need == 0:
        return 0, int(checks.value), cache1, cache2
    rc = call(extra_flags | capi.TRAVERSE_COUNTS_VALID, cache1.ptr, need, total, checks)
    if rc != capi.OK:
        _raise(rc, handle, "traverse (BFS, write)")
    return int(total.value), int(checks.value), cache1, cache2


def _bfs_narrow(narrow, tr: "BVHTraversal", leaves1: "DeviceArray", leaves2: "DeviceArray", single: bool, I: np.dtype) -> "BVHTraversal":
    """`narrow(leaf1, leaf2)` after a positive leaf test (traverse_single_gpu.jl:188, traverse_pair_gpu.jl): post-filter over
    the positions reported under IBVH_TRAVERSE_POSITIONS."""
    tI = torch.int32 if I.itemsize == 4 else torch.int64
    pos = tr.contacts.tensor.view(tI).reshape(-1, 2)
    b1 = _gather_records(leaves1, pos[:, 0] - 1)
    b2 = _gather_records(leaves2, pos[:, 1] - 1)
    keep = _eval_narrow(narrow, b1, b2)
    i1, i2 = b1["index"][keep], b2["index"][keep]
    out = np.zeros(int(keep.sum()), pair_dtype(I))
    if single:
        out["a"], out["b"] = np.minimum(i1, i2), np.maximum(i1, i2)
    else:
        out["a"], out["b"] = i1, i2
    return BVHTraversal(tr.start_level1, tr.start_level2, tr.num_checks, len(out), DeviceArray.from_numpy(out, device=leaves1.device), tr.cache2)


def _traverse_bfs(bvh: BVH, bvh2, alg, start_level, start_level1, start_level2, narrow, cache) -> BVHTraversal:
    """traverse(bvh[, bvh2], BFSTraversal(); ...) — breadth_first/traverse_single.jl:1-66, traverse_pair.jl:1-158."""
    lib = capi.lib()
    device = bvh.leaves.device
    _resolve_outstanding(device.index)
    I = bvh.index_dtype
    pos_flag = capi.TRAVERSE_POSITIONS if narrow is not None else 0
    use_cache = None if narrow is not None else cache
    if bvh2 is None:
        sl = default_start_level(bvh, alg) if start_level is None else int(start_level)
        if not (bvh.tree.levels >= sl >= bvh.built_level):                              # traverse_single.jl:10
            raise ArgumentError("bvh.tree.levels >= start_level >= bvh.built_level must hold")
        if bvh.tree.real_nodes <= 1:                                                    # :17-21
            return BVHTraversal(sl, 0, 0, 0, DeviceArray.empty(0, pair_dtype(I), device), DeviceArray.empty(0, pair_dtype(I), device))
        cb = bvh._c_bvh()

        def call(flags, p_contacts, capacity, total, checks):
            params = capi.TraverseParams(sl, 0, -1, flags, 0, 0, None)
            return lib.ibvh_traverse_bfs_single(bvh._handle, C.byref(cb), C.byref(params), p_contacts, capacity, C.byref(total),
                                                C.byref(checks), _stream_ptr(device.index))

        total, checks, c1, c2 = _bfs_protocol(call, bvh._handle, device, I, use_cache, pos_flag)
        out = BVHTraversal(sl, 0, checks, total, c1, c2)
        return _bfs_narrow(narrow, out, bvh.leaves, bvh.leaves, True, I) if narrow is not None else out
    sl1 = default_start_level(bvh, alg) if start_level1 is None else int(start_level1)
    sl2 = default_start_level(bvh2, alg) if start_level2 is None else int(start_level2)
    if not (bvh.tree.levels >= sl1 >= bvh.built_level):                                 # traverse_pair.jl:10-11
        raise ArgumentError("bvh1.tree.levels >= start_level1 >= bvh1.built_level must hold")
    if not (bvh2.tree.levels >= sl2 >= bvh2.built_level):
        raise ArgumentError("bvh2.tree.levels >= start_level2 >= bvh2.built_level must hold")
    if bvh.index_dtype != bvh2.index_dtype:
        raise ArgumentError("both BVHs must use one index type")
    if bvh.leaves.device != bvh2.leaves.device:
        raise ArgumentError("both BVHs must live on the same device")
    c1b, c2b = bvh._c_bvh(), bvh2._c_bvh()

    def call(flags, p_contacts, capacity, total, checks):
        return lib.ibvh_traverse_bfs_pair(bvh._handle, C.byref(c1b), C.byref(c2b), sl1, sl2, flags, p_contacts, capacity, C.byref(total),
                                          C.byref(checks), _stream_ptr(device.index))

    total, checks, c1, c2 = _bfs_protocol(call, bvh._handle, device, I, use_cache, pos_flag)
    out = BVHTraversal(sl1, sl2, checks, total, c1, c2)
    return _bfs_narrow(narrow, out, bvh.leaves, bvh2.leaves, False, I) if narrow is not None else out


def traverse(bvh: BVH, bvh2=None, alg=None, *, start_level: Optional[int] = None, start_level1: Optional[int] = None,
             start_level2: Optional[int] = None, narrow=None, cache: Optional[BVHTraversal] = None, options: BVHOptions = None,
             ordered: bool = True, reference_shaped: bool = False, packet: bool = False, walk: bool = False,
             query_range=None, peer=None, defer: bool = False) -> BVHTraversal:
    """`traverse(bvh[, bvh2], LVTTraversal(); start_level[1,2], narrow, cache, options)`.

    Extensions over the reference signature (all keyword-only, defaults reproduce the reference):
    `ordered=False` selects the unordered emission, `reference_shaped=True` the proxy of the reference's
    own GPU kernel, `packet=True` forces the warp-packet schedule (default: group walk + dense tiles for
    BBox nodes), `query_range=(begin, count)` restricts the query leaves (multi-GPU shard).
    `defer=True` (with `ordered=False` and a `cache` whose cache1 is allocated): the traversal is only enqueued;
    `num_contacts` waits for it on first use, so the next `BVH(...)` can be enqueued before the host blocks (at
    most one deferred traversal per device may be outstanding).
    `peer=dist.PeerGather(...)` (with `ordered=False`) fuses the sharded traversal with the all-gather of the
    contact shards: collective over the ranks, the returned (unordered) list holds the contacts of ALL ranks.
    """
    if isinstance(bvh2, (int, np.integer)) and not isinstance(bvh2, bool):   # old interface traverse(bvh, start_level[, cache]): BFS (traverse/traverse.jl:233-241)
        start_level, alg, bvh2 = int(bvh2), BFSTraversal(), None
    if bvh2 is not None and not isinstance(bvh2, BVH):       # traverse(bvh, alg)
        alg, bvh2 = bvh2, None
    if isinstance(alg, (int, np.integer)) and not isinstance(alg, bool):     # old interface traverse(bvh1, bvh2, start_level1[, start_level2]): BFS (traverse/traverse.jl:244-256)
        start_level1, alg = int(alg), BFSTraversal()
    if isinstance(alg, BFSTraversal):
        if defer or peer is not None or query_range is not None or reference_shaped or packet or walk:
            raise ArgumentError("BFSTraversal takes start_level[1,2], narrow and cache only")
        return _traverse_bfs(bvh, bvh2, alg, start_level, start_level1, start_level2, narrow, cache)
    if alg is not None and not isinstance(alg, LVTTraversal):
        raise ArgumentError(f"Traversal algorithm not implemented: {alg}")
    if narrow is not None and (defer or peer is not None):
        raise ArgumentError("a custom `narrow` predicate is a post-filter on the host: not with defer=True or peer=...")
    pos_flag = capi.TRAVERSE_POSITIONS if narrow is not None else 0
    lib = capi.lib()
    _resolve_outstanding(bvh.leaves.device.index)      # one deferred traversal per handle: a new one finishes the old one first
    qb, qc = (0, -1) if query_range is None else (int(query_range[0]), int(query_range[1]))
    if peer is not None and (ordered or reference_shaped or packet or walk):
        raise ArgumentError("the fused multi-GPU traversal is the unordered default schedule (ordered=False)")

    def run(call, handle, device, I, nq):
        if (defer and peer is None and not ordered and not (reference_shaped or packet or walk) and cache is not None
                and len(cache.cache1) > 0 and cache.cache1.device == device and cache.cache1.dtype == pair_dtype(I)):
            total = C.c_int64(0)
            rc = call(capi.TRAVERSE_UNORDERED | capi.TRAVERSE_DEFER, None, cache.cache1.ptr, len(cache.cache1), total)
            if rc == capi.OK and total.value == -1:
                return _Pending(handle), cache.cache1, cache.cache2
            if rc == capi.OK:
                return int(total.value), cache.cache1, cache.cache2
            # anything else (e.g. capacity): the synchronous protocol below sorts it out
        if peer is None:
            return _run_two_phase(call, handle, device, I, nq, None if narrow is not None else cache, ordered, reference_shaped, packet, walk,
                                  extra_flags=pos_flag)
        if peer.pair_bytes != pair_dtype(I).itemsize or peer.device != device:
            raise ArgumentError("PeerGather pair size / device do not match the BVH")
        total = _fused_call(lambda pref, tot: call(capi.TRAVERSE_UNORDERED, None, None, 0, tot, pref), peer, handle, "fused traverse")
        c2 = cache.cache2 if cache is not None else DeviceArray.empty(0, I, device)
        return int(total.value), DeviceArray(peer.list_area(), pair_dtype(I)), c2

    if bvh2 is None:
        sl = default_start_level(bvh) if start_level is None else int(start_level)
        if not (bvh.built_level <= sl <= bvh.tree.levels <= 32):                      # traverse_single.jl:9-11
            raise ArgumentError("bvh.built_level <= start_level <= bvh.tree.levels <= 32 must hold")
        I = bvh.index_dtype
        device = bvh.leaves.device
        if bvh.tree.real_nodes <= 1:                                                  # traverse_single.jl:17-21
            return BVHTraversal(sl, 0, 0, 0, DeviceArray.empty(0, pair_dtype(I), device), DeviceArray.empty(0, I, device))
        cb = bvh._c_bvh()
        nq = len(bvh.leaves) if qc < 0 else qc

        def call(flags, p_counts, p_contacts, capacity, total, peer_ref=None):
            params = capi.TraverseParams(sl, qb, qc, flags, 0, 0, peer_ref)
            return lib.ibvh_traverse_single(bvh._handle, C.byref(cb), C.byref(params), p_counts, p_contacts, capacity,
                                            C.byref(total), _stream_ptr(device.index))

        total, c1, c2 = run(call, bvh._handle, device, I, nq)
        if isinstance(total, _Pending):
            return total.attach(BVHTraversal(sl, 0, 0, 0, c1, c2),
                                lambda: traverse(bvh, start_level=sl, cache=BVHTraversal(sl, 0, 0, 0, c1, c2), ordered=False, query_range=query_range),
                                bvhs=(bvh,), device_index=device.index)
        out = _finish_stats(BVHTraversal(sl, 0, 0, total, c1, c2), bvh._handle)
        if narrow is not None:
            out = _narrow_pairs(narrow, out, bvh.leaves, bvh.leaves, True, I, nq, 0, max(qb, 0))
        return out

    # pair — traverse_pair.jl:1-116
    sl1 = default_start_level(bvh) if start_level1 is None else int(start_level1)
    sl2 = default_start_level(bvh2) if start_level2 is None else int(start_level2)
    if not (bvh.built_level <= sl1 <= bvh.tree.levels <= 32):
        raise ArgumentError("bvh1.built_level <= start_level1 <= bvh1.tree.levels <= 32 must hold")
    if not (bvh2.built_level <= sl2 <= bvh2.tree.levels <= 32):
        raise ArgumentError("bvh2.built_level <= start_level2 <= bvh2.tree.levels <= 32 must hold")
    if bvh.index_dtype != bvh2.index_dtype:
        raise ArgumentError("get_index_type(bvh2) === I must hold")
    if bvh.leaves.device != bvh2.leaves.device:
        raise ArgumentError("both BVHs must live on the same device")
    flip = not (len(bvh.leaves) >= len(bvh2.leaves))                                  # traverse_pair.jl:16-36
    queries, target, sl_t = (bvh2, bvh, sl1) if flip else (bvh, bvh2, sl2)
    I = bvh.index_dtype
    device = bvh.leaves.device
    cq, ct = queries._c_bvh(), target._c_bvh()
    nq = len(queries.leaves) if qc < 0 else qc

    def call(flags, p_counts, p_contacts, capacity, total, peer_ref=None):
        params = capi.TraverseParams(sl_t, qb, qc, flags, 1 if flip else 0, 0, peer_ref)
        return lib.ibvh_traverse_pair(bvh._handle, C.byref(cq), C.byref(ct), C.byref(params), p_counts, p_contacts, capacity,
                                      C.byref(total), _stream_ptr(device.index))

    total, c1, c2 = run(call, bvh._handle, device, I, nq)
    if isinstance(total, _Pending):
        return total.attach(BVHTraversal(sl1, sl2, 0, 0, c1, c2),
                            lambda: traverse(bvh, bvh2, start_level1=sl1, start_level2=sl2, cache=BVHTraversal(sl1, sl2, 0, 0, c1, c2),
                                             ordered=False, query_range=query_range),
                            bvhs=(bvh, bvh2), device_index=device.index)
    out = _finish_stats(BVHTraversal(sl1, sl2, 0, total, c1, c2), bvh._handle)
    if narrow is not None:
        out = _narrow_pairs(narrow, out, bvh.leaves, bvh2.leaves, False, I, nq, 1 if flip else 0, max(qb, 0))
    return out


def sort_contacts(traversal: BVHTraversal, unique: bool = False) -> BVHTraversal:
    """Sort a traversal's contact list ascending by (a, b) on the device, in place (optionally dropping repeated pairs) — what
    the reference's tests do on the host with `sort(traversal.contacts)` (test/gputests.jl:73-78) before comparing lists.
    For the unordered emission modes (`ordered=False`, `BFSTraversal()`), whose lists are sets. SURVEY.md §8f-2."""
    c1 = traversal.cache1
    n = traversal.num_contacts
    if not c1.is_cuda:
        raise ArgumentError("sort_contacts: the contacts must live on a CUDA device")
    handle = get_handle(c1.device)
    _resolve_outstanding(c1.device.index)
    kept = C.c_int64(n)
    rc = capi.lib().ibvh_sort_contacts(handle, c1.ptr if n else None, n, c1.dtype.itemsize // 2, 1 if unique else 0, C.byref(kept),
                                       _stream_ptr(c1.device.index))
    if rc != capi.OK:
        _raise(rc, handle, "sort_contacts")
    traversal.num_contacts = int(kept.value)
    return traversal


def traverse_rays(bvh: BVH, points, directions, alg=None, *, start_level: int = 1, narrow=None,
                  cache: Optional[BVHTraversal] = None, options: BVHOptions = None, ordered: bool = True,
                  id_base: int = 0, peer=None) -> BVHTraversal:
    """`traverse_rays(bvh, points, directions, LVTTraversal(); start_level=1, narrow, cache, options)`.

    `points` / `directions`: (3, R) numpy arrays (the reference's column-major 3xR matrices), or torch
    CUDA tensors of shape (R, 3) — the same memory layout — in the BVH float type.
    """
    bfs = isinstance(alg, BFSTraversal)
    if alg is not None and not bfs and not isinstance(alg, LVTTraversal):
        raise ArgumentError(f"Raytracing algorithm not implemented: {alg}")
    if bfs and peer is not None:
        raise ArgumentError("BFSTraversal takes start_level, narrow and cache only")
    if narrow is not None and peer is not None:
        raise ArgumentError("a custom `narrow` predicate is a post-filter on the host: not with peer=...")
    lib = capi.lib()
    if bvh.types.node_float_bytes not in (0, bvh.types.float_bytes):
        # isintersection(b::BBox{T}, p::...{T}, d::...{T}) where T (isintersection.jl:1-5): no method for mixed float types
        raise ArgumentError("traverse_rays needs leaves and nodes of one float type (the reference has no mixed-type ray test)")
    T = {4: np.float32, 8: np.float64}[bvh.types.float_bytes]
    device = bvh.leaves.device
    _resolve_outstanding(device.index)                 # (the ray traversal shares the handle's read-back slots)

    def to_dev(a):
        if isinstance(a, torch.Tensor):
            if a.dim() != 2 or a.shape[1] != 3:
                raise ArgumentError("device ray arrays must have shape (R, 3)")
            return a.to(device=device, dtype={4: torch.float32, 8: torch.float64}[bvh.types.float_bytes]).contiguous()
        a = np.asarray(a)
        if a.ndim != 2 or a.shape[0] != 3:                                           # leaf_vs_tree.jl:13
            raise ArgumentError("size(points, 1) == size(directions, 1) == 3 must hold")
        return torch.from_numpy(np.ascontiguousarray(a.T.astype(T))).to(device, non_blocking=True)

    p, d = to_dev(points), to_dev(directions)
    if p.shape != d.shape:                                                            # leaf_vs_tree.jl:14
        raise ArgumentError("size(points, 2) == size(directions, 2) must hold")
    if not (bvh.built_level <= start_level <= bvh.tree.levels <= 32):
        raise ArgumentError("bvh.built_level <= start_level <= bvh.tree.levels <= 32 must hold")
    I = bvh.index_dtype
    nrays = p.shape[0]
    if nrays == 0 and peer is None:                                                   # leaf_vs_tree.jl:22-26
        return BVHTraversal(start_level, 0, 0, 0, DeviceArray.empty(0, pair_dtype(I), device), DeviceArray.empty(0, pair_dtype(I), device))
    cb = bvh._c_bvh()

    if bfs:                                        # raytrace/breadth_first/breadth_first.jl:1-66
        def call_bfs(flags, p_contacts, capacity, total, checks):
            params = capi.TraverseParams(int(start_level), 0, -1, flags, 0, int(id_base), None)
            return lib.ibvh_traverse_bfs_rays(bvh._handle, C.byref(cb), p.data_ptr(), d.data_ptr(), nrays, C.byref(params), p_contacts, capacity,
                                              C.byref(total), C.byref(checks), _stream_ptr(device.index))

        total, checks, c1, c2 = _bfs_protocol(call_bfs, bvh._handle, device, I, None if narrow is not None else cache,
                                              capi.TRAVERSE_POSITIONS if narrow is not None else 0)
        out = BVHTraversal(start_level, 0, checks, total, c1, c2)
        if narrow is not None:
            tI = torch.int32 if I.itemsize == 4 else torch.int64
            pr = out.contacts.tensor.view(tI).reshape(-1, 2)
            lv = _gather_records(bvh.leaves, pr[:, 0] - 1)
            rid = (pr[:, 1].to(torch.int64) - 1 - int(id_base))
            keep = _eval_narrow(narrow, lv, p.index_select(0, rid).cpu().numpy(), d.index_select(0, rid).cpu().numpy())
            res = np.zeros(int(keep.sum()), pair_dtype(I))
            res["a"], res["b"] = lv["index"][keep], pr[:, 1].cpu().numpy()[keep]
            out = BVHTraversal(start_level, 0, checks, len(res), DeviceArray.from_numpy(res, device=device), c2)
        return out

    def call(flags, p_counts, p_contacts, capacity, total, peer_ref=None):
        params = capi.TraverseParams(int(start_level), 0, -1, flags, 0, int(id_base), peer_ref)
        return lib.ibvh_traverse_rays(bvh._handle, C.byref(cb), p.data_ptr() if nrays else None, d.data_ptr() if nrays else None, nrays,
                                      C.byref(params), p_counts, p_contacts, capacity, C.byref(total), _stream_ptr(device.index))

    if peer is not None:
        # fused ray traversal + all-gather (collective): the hits of ALL ranks' ray shards land in every rank's list
        if ordered:
            raise ArgumentError("the fused multi-GPU ray traversal is unordered (ordered=False)")
        if peer.pair_bytes != pair_dtype(I).itemsize or peer.device != device:
            raise ArgumentError("PeerGather pair size / device do not match the BVH")
        total = _fused_call(lambda pref, tot: call(capi.TRAVERSE_UNORDERED, None, None, 0, tot, pref), peer, bvh._handle, "fused traverse_rays")
        c2 = cache.cache2 if cache is not None else DeviceArray.empty(0, I, device)
        return BVHTraversal(start_level, 0, 0, int(total.value), DeviceArray(peer.list_area(), pair_dtype(I)), c2)
    total, c1, c2 = _run_two_phase(call, bvh._handle, device, I, nrays, None if narrow is not None else cache, ordered, False,
                                   extra_flags=capi.TRAVERSE_POSITIONS if narrow is not None else 0)
    out = BVHTraversal(start_level, 0, 0, total, c1, c2)
    if narrow is not None:
        # narrow(leaf, point, direction) after a positive ray / leaf test (raytrace/leaf_vs_tree/leaf_vs_tree.jl:194)
        tI = torch.int32 if I.itemsize == 4 else torch.int64
        pr = out.contacts.tensor.view(tI).reshape(-1, 2)
        lv = _gather_records(bvh.leaves, pr[:, 0] - 1)
        rid = (pr[:, 1].to(torch.int64) - 1 - int(id_base))
        keep = _eval_narrow(narrow, lv, p.index_select(0, rid).cpu().numpy(), d.index_select(0, rid).cpu().numpy())
        res = np.zeros(int(keep.sum()), pair_dtype(I))
        res["a"], res["b"] = lv["index"][keep], pr[:, 1].cpu().numpy()[keep]
        counts = torch.bincount(rid[torch.from_numpy(keep).to(device)], minlength=nrays)[:nrays].cumsum(0).to(tI)
        out = BVHTraversal(start_level, 0, 0, len(res), DeviceArray.from_numpy(res, device=device),
                           DeviceArray(counts.view(torch.uint8).reshape(-1).contiguous(), I))
    return out


# ---------------------------------------------------------------------------------------------
# stage-level entry points (for parity tests / profiling of one stage)
# ---------------------------------------------------------------------------------------------
def wrap_bounding_volumes(volumes: Union[np.ndarray, DeviceArray], options: BVHOptions = None, device=None) -> DeviceArray:
    """build.jl:328-352."""
    options = options or BVHOptions()
    src = DeviceArray.from_numpy(volumes, device=torch.device("cuda", _device_index(device))) if isinstance(volumes, np.ndarray) else volumes
    vol = _volume_of_dtype(src.dtype)
    if vol is None:
        raise ArgumentError("expected raw BSphere / BBox volumes")
    ldt = leaf_dtype(vol, options.index_dtype, options.morton.eltype)
    out = DeviceArray.empty(len(src), ldt, src.device)
    types = _types(vol, BBox({4: np.float32, 8: np.float64}[vol.float_bytes]), options.index_dtype, options.morton.eltype)
    h = get_handle(src.device)
    with torch.cuda.device(src.device.index):
        rc = capi.lib().ibvh_wrap(h, src.ptr, len(src), C.byref(types), out.ptr, _stream_ptr(src.device.index))
    if rc != capi.OK:
        _raise(rc, h, "ibvh_wrap")
    return out


def volumes_from_triangles(triangles, volume_type: VolumeType = None, device=None) -> DeviceArray:
    """`[BSphere{T}(tri) for tri in mesh]` / `BBox{T}(tri)` on the device (bsphere.jl:43-112, bbox.jl:59-70).
    `triangles`: (n, 3, 3) numpy array or CUDA tensor of vertex coordinates in the volume float type."""
    volume_type = volume_type or BSphere(np.float32)
    T = {4: np.float32, 8: np.float64}[volume_type.float_bytes]
    if isinstance(triangles, torch.Tensor):
        t = triangles.to(dtype={4: torch.float32, 8: torch.float64}[volume_type.float_bytes]).contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(triangles, T))).to(torch.device("cuda", _device_index(device)), non_blocking=True)
    if t.dim() != 3 or tuple(t.shape[1:]) != (3, 3):
        raise ArgumentError("triangles must have shape (n, 3, 3)")
    out = DeviceArray.empty(t.shape[0], volume_type.dtype, t.device)
    h = get_handle(t.device)
    with torch.cuda.device(t.device.index):
        rc = capi.lib().ibvh_volumes_from_triangles(h, t.data_ptr(), t.shape[0], volume_type.kind, volume_type.float_bytes, out.ptr,
                                                    _stream_ptr(t.device.index))
    if rc != capi.OK:
        _raise(rc, h, "ibvh_volumes_from_triangles")
    return out


def _leaf_types(leaves: DeviceArray, node: VolumeType = None) -> capi.Types:
    vol = _volume_of_dtype(leaves.dtype["volume"])
    node = node or BBox({4: np.float32, 8: np.float64}[vol.float_bytes])
    return _types(vol, node, leaves.dtype["index"], leaves.dtype["morton"])


def morton_encode(leaves: DeviceArray, options: BVHOptions = None):
    """`morton_encode!(bounding_volumes, options)` (morton/morton.jl:28-35): in place; returns (mins, maxs) used."""
    options = options or BVHOptions()
    if leaves.dtype["morton"] != options.morton.eltype:                               # morton.jl:38-42
        raise ArgumentError(f"Bounding volume Morton type {leaves.dtype['morton']} does not match options Morton type {options.morton.eltype}")
    types = _leaf_types(leaves)
    h = get_handle(leaves.device)
    m = options.morton
    mins = (C.c_double * 3)(*[float(x) for x in m.mins])
    maxs = (C.c_double * 3)(*[float(x) for x in m.maxs])
    omin, omax = (C.c_double * 3)(), (C.c_double * 3)()
    with torch.cuda.device(leaves.device.index):
        rc = capi.lib().ibvh_morton_encode(h, leaves.ptr, len(leaves), C.byref(types), 1 if m.compute_extrema else 0, mins, maxs,
                                           omin, omax, _stream_ptr(leaves.device.index))
    if rc != capi.OK:
        _raise(rc, h, "ibvh_morton_encode")
    return np.array(omin[:]), np.array(omax[:])


def sort_leaves(leaves: DeviceArray):
    types = _leaf_types(leaves)
    h = get_handle(leaves.device)
    with torch.cuda.device(leaves.device.index):
        rc = capi.lib().ibvh_sort_leaves(h, leaves.ptr, len(leaves), C.byref(types), _stream_ptr(leaves.device.index))
    if rc != capi.OK:
        _raise(rc, h, "ibvh_sort_leaves")


def aggregate(leaves: DeviceArray, node_type: VolumeType = None, built_level: int = 1) -> DeviceArray:
    types = _leaf_types(leaves, node_type)
    node_type = node_type or BBox({4: np.float32, 8: np.float64}[types.float_bytes])
    n = len(leaves)
    nodes = DeviceArray.empty(int(capi.lib().ibvh_num_nodes(n)), node_type.dtype, leaves.device)
    h = get_handle(leaves.device)
    with torch.cuda.device(leaves.device.index):
        rc = capi.lib().ibvh_aggregate(h, leaves.ptr, n, C.byref(types), nodes.ptr if len(nodes) else None, int(built_level),
                                       _stream_ptr(leaves.device.index))
    if rc != capi.OK:
        _raise(rc, h, "ibvh_aggregate")
    return nodes
