# ImplicitBVHB200Ext.jl — package extension of ImplicitBVH.jl (v0.7.1) that routes the hot path to libibvh_b200.so
# when the bounding volumes live in a `CuVector`: the methods below keep the reference's signatures
# (src/build.jl:198-271, src/traverse/traverse.jl:115-230, src/raytrace/raytrace.jl:71-81) and `ccall` the C ABI of
# include/ibvh.h. Add to the reference's Project.toml:
#
#     [weakdeps]
#     CUDA = "052768ef-5323-5732-b1bb-66c8b64840ba"
#     [extensions]
#     ImplicitBVHB200Ext = "CUDA"
#
# and drop this file into `ext/`. STATUS: written against the reference's sources and the header; it has NOT been
# executed (no Julia in the build image — the Python / ctypes mirror `implicitbvh.jl_b200/api.py` is the caller that
# exercises every entry point in the tests). Everything user-visible (leaves, nodes, skips, cache1, cache2, rays)
# stays a CuArray owned by Julia; the library owns only its per-device scratch.
module ImplicitBVHB200Ext

using ImplicitBVH
using ImplicitBVH: BVH, BVHOptions, BVHTraversal, BoundingVolume, BSphere, BBox, IndexPair, ImplicitTree,
                   LVTTraversal, BFSTraversal, get_index_type, compute_build_level, compute_skips!, default_start_level,
                   check_bounding_volume_types
using CUDA

const LIB = get(ENV, "IBVH_B200_LIB", "libibvh_b200.so")

# ---- mirrors of the C structs (include/ibvh.h) --------------------------------------------------------------
struct Types                      # ibvh_types_t
    leaf_kind::Int32; float_bytes::Int32; index_bytes::Int32; morton_bytes::Int32; node_kind::Int32; node_float_bytes::Int32
end
struct CBvh                       # ibvh_bvh_t
    d_leaves::CuPtr{Cvoid}; d_nodes::CuPtr{Cvoid}; n::Int64; built_level::Int64; types::Types; build_id::UInt64
end
struct Params                     # ibvh_traverse_params_t
    start_level::Int64; query_begin::Int64; query_count::Int64; flags::UInt32; flip::Int32; id_base::Int64; peer::Ptr{Cvoid}
end

const ORDERED = UInt32(0); const UNORDERED = UInt32(1); const COUNTS_VALID = UInt32(4)
const DEFER = UInt32(64); const POSITIONS = UInt32(128)
const ERR_CAPACITY = 5; const ERR_AGAIN = 8

kind(::Type{<:BSphere}) = Int32(0)
kind(::Type{<:BBox}) = Int32(1)
floattype(::Type{BSphere{T}}) where T = T
floattype(::Type{BBox{T}}) where T = T
function types(::Type{BoundingVolume{V, I, M}}, ::Type{N}) where {V, I, M, N}
    fl, fn = sizeof(floattype(V)), sizeof(floattype(N))
    Types(kind(V), fl, sizeof(I), sizeof(M), kind(N), fn == fl ? 0 : fn)     # 0 = node float type == leaf float type
end

# The reference's BVH struct has no room for the id of the build's sidecar (ibvh_bvh_t.build_id): it is kept beside the
# leaves array it belongs to. A BVH whose leaves are not in the table (built elsewhere, copied) simply passes 0.
const BUILD_IDS = WeakKeyDict{Any, UInt64}()
build_id(bvh::BVH) = get(BUILD_IDS, bvh.leaves, UInt64(0))
cbvh(bvh::BVH, t::Types) = CBvh(pointer(bvh.leaves), length(bvh.nodes) > 0 ? pointer(bvh.nodes) : CU_NULL, length(bvh.leaves),
                               Int64(bvh.built_level), t, build_id(bvh))

# ---- handle per device, errors ------------------------------------------------------------------------------
const HANDLES = Dict{Int, Ptr{Cvoid}}()
function handle()
    dev = CUDA.deviceid(CUDA.device())
    get!(HANDLES, dev) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ibvh_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, dev), C_NULL)
        h[]
    end
end

function check(rc, h)
    rc == 0 && return nothing
    rc == 1 && throw(ArgumentError("ibvh-b200: argument check failed (same conditions as the reference's @argcheck)"))
    rc == 2 && throw(DomainError(0, "must have at least one geometry!"))            # implicit_tree.jl:78-80
    msg = unsafe_string(ccall((:ibvh_status_string, LIB), Cstring, (Cint,), rc))
    h == C_NULL || (msg *= ": " * unsafe_string(ccall((:ibvh_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
    error("ibvh-b200 status $rc: $msg")
end

stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))      # the task-local CUDA.jl stream: all work is enqueued there

# ---- BVH(bounding_volumes::CuVector, node_type; built_level, cache, options) — src/build.jl:198-271 ------------
function ImplicitBVH.BVH(
    bounding_volumes::CuVector{L},
    node_type::Type{N}=BBox{Float32};
    built_level::Union{Integer, AbstractFloat}=1,
    cache::Union{Nothing, BVH}=nothing,
    options=BVHOptions(),
) where {L, N}
    I = get_index_type(options)
    M = eltype(options.morton)
    wrapped = L <: BoundingVolume
    if wrapped
        check_bounding_volume_types(bounding_volumes, options)                      # ArgumentError on I / M mismatch (build.jl:355-361)
        leaves = bounding_volumes                                                   # sorted IN PLACE, as in the reference
    else
        leaves = CuVector{BoundingVolume{L, I, M}}(undef, length(bounding_volumes)) # the library wraps: index = position, morton = 0
    end
    numbv = length(bounding_volumes)
    tree = ImplicitTree{I}(numbv)                                                   # DomainError if numbv < 1

    # skips: reused from the cache exactly as the reference does (build.jl:232-239)
    if isnothing(cache)
        skips = CuVector{I}(undef, tree.levels)
    else
        eltype(cache.skips) === I || throw(ArgumentError("eltype(cache.skips) === I must hold"))
        skips = cache.skips
        length(skips) == tree.levels || resize!(skips, tree.levels)
    end
    compute_skips!(skips, tree)

    built_ilevel = compute_build_level(tree, built_level)

    # nodes: reused from the cache (build.jl:256-263)
    num_nodes = Int(tree.real_nodes - tree.real_leaves)
    if isnothing(cache)
        nodes = CuVector{N}(undef, num_nodes)
    else
        eltype(cache.nodes) === N || throw(ArgumentError("eltype(cache.nodes) === N must hold"))
        nodes = cache.nodes
        length(nodes) == num_nodes || resize!(nodes, num_nodes)
    end

    alg = options.morton
    mins = Float64[Float64(x) for x in alg.mins]
    maxs = Float64[Float64(x) for x in alg.maxs]
    h = handle()
    rc = ccall((:ibvh_build, LIB), Cint,
        (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{Types}, CuPtr{Cvoid}, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
        h, wrapped ? CU_NULL : pointer(bounding_volumes), pointer(leaves), numbv, types(eltype(leaves), N),
        num_nodes > 0 ? pointer(nodes) : CU_NULL, Int64(built_ilevel), alg.compute_extrema ? 1 : 0, mins, maxs, stream())
    check(rc, h)
    BUILD_IDS[leaves] = ccall((:ibvh_last_build_id, LIB), UInt64, (Ptr{Cvoid},), h)
    BVH(I(built_ilevel), tree, skips, nodes, leaves)
end

# ---- the reference's count -> scan -> (grow cache1) -> write protocol over one C entry point --------------------
# `call(flags, counts_ptr, contacts_ptr, capacity, total)` wraps the ccall. cache1 / cache2 are reused and grown only
# when too small (traverse_single.jl:23-67). Returns (num_contacts, cache1, cache2).
function two_phase(call, h, ::Type{I}, nqueries::Int, cache, flags::UInt32) where I
    cache2 = (isnothing(cache) || length(cache.cache2) < nqueries) ? CuVector{I}(undef, nqueries) : cache.cache2
    cache1 = isnothing(cache) ? CuVector{IndexPair{I}}(undef, 0) : cache.cache1
    total = Ref{Int64}(0)
    if length(cache1) > 0
        rc = call(flags, pointer(cache2), pointer(cache1), length(cache1), total)
        rc == 0 && return Int(total[]), cache1, cache2
        rc == ERR_CAPACITY || check(rc, h)
    else
        check(call(flags, pointer(cache2), CU_NULL, 0, total), h)                   # count only
        total[] == 0 && return 0, cache1, cache2
    end
    cache1 = CuVector{IndexPair{I}}(undef, Int(total[]))                            # grow: "resize only if too small"
    check(call(flags | COUNTS_VALID, pointer(cache2), pointer(cache1), length(cache1), total), h)   # the reference's second pass
    Int(total[]), cache1, cache2
end

num_checks(h) = (st = zeros(Int64, 4); ccall((:ibvh_last_traversal_stats, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h, st);
                 st[4] > 0 ? Int(st[1] + st[2]) : 0)

# `narrow` is a Julia closure: it cannot cross a C ABI, but the reference evaluates it only AFTER a positive leaf test
# (traverse_single.jl:170, traverse_pair.jl:206), so the same list results from asking the library for leaf POSITIONS
# (IBVH_TRAVERSE_POSITIONS) and filtering on the device with the closure broadcast over the leaves at those positions.
isdefault(narrow) = narrow === nothing
function apply_narrow(narrow, positions::CuVector{IndexPair{I}}, leaves1, leaves2, single::Bool) where I
    keep = map(p -> narrow(leaves1[p[1]], leaves2[p[2]]), positions)                # one thread per candidate pair
    kept = positions[keep]
    map(kept) do p
        a, b = leaves1[p[1]].index, leaves2[p[2]].index
        single ? (min(a, b), max(a, b)) : (a, b)
    end
end

# ---- traverse(bvh, ::LVTTraversal; start_level, narrow, cache, options) — leaf_vs_tree/traverse_single.jl:1-79 ----
function ImplicitBVH.traverse(
    bvh::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    alg::LVTTraversal;
    start_level::Int=default_start_level(bvh, alg),
    narrow=nothing,
    cache::Union{Nothing, BVHTraversal}=nothing,
    options=BVHOptions(),
) where {I, N, L}
    bvh.built_level <= start_level <= bvh.tree.levels <= 32 ||
        throw(ArgumentError("bvh.built_level <= start_level <= bvh.tree.levels <= 32 must hold"))
    n = length(bvh.leaves)
    if bvh.tree.real_nodes <= 1                                                     # traverse_single.jl:17-21
        return BVHTraversal(start_level, 0, 0, CuVector{IndexPair{I}}(undef, 0), CuVector{I}(undef, 0))
    end
    h = handle()
    cb = cbvh(bvh, types(L, N))
    flags = isdefault(narrow) ? ORDERED : (ORDERED | POSITIONS)
    call(fl, pcounts, pcontacts, cap, total) = ccall((:ibvh_traverse_single, LIB), Cint,
        (Ptr{Cvoid}, Ref{CBvh}, Ref{Params}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{Int64}, Ptr{Cvoid}),
        h, cb, Params(start_level, 0, -1, fl, 0, 0, C_NULL), pcounts, pcontacts, cap, total, stream())
    total, cache1, cache2 = two_phase(call, h, I, n, isdefault(narrow) ? cache : nothing, flags)
    if !isdefault(narrow)
        cache1 = apply_narrow(narrow, view(cache1, 1:total), bvh.leaves, bvh.leaves, true)
        total = length(cache1)
    end
    BVHTraversal(start_level, 0, num_checks(h), total, cache1, cache2)
end

# ---- traverse(bvh1, bvh2, ::LVTTraversal; start_level1, start_level2, ...) — leaf_vs_tree/traverse_pair.jl:1-116 ----
function ImplicitBVH.traverse(
    bvh1::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    bvh2::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    alg::LVTTraversal;
    start_level1::Int=default_start_level(bvh1, alg),
    start_level2::Int=default_start_level(bvh2, alg),
    narrow=nothing,
    cache::Union{Nothing, BVHTraversal}=nothing,
    options=BVHOptions(),
) where {I, N, L}
    bvh1.built_level <= start_level1 <= bvh1.tree.levels <= 32 ||
        throw(ArgumentError("bvh1.built_level <= start_level1 <= bvh1.tree.levels <= 32 must hold"))
    bvh2.built_level <= start_level2 <= bvh2.tree.levels <= 32 ||
        throw(ArgumentError("bvh2.built_level <= start_level2 <= bvh2.tree.levels <= 32 must hold"))
    # the tree with more leaves provides the queries; FLIP when that is bvh2 (traverse_pair.jl:16-36)
    flip = !(length(bvh1.leaves) >= length(bvh2.leaves))
    queries, target, sl_t = flip ? (bvh2, bvh1, start_level1) : (bvh1, bvh2, start_level2)
    h = handle()
    t = types(L, N)
    cq = cbvh(queries, t)
    ct = cbvh(target, t)
    flags = isdefault(narrow) ? ORDERED : (ORDERED | POSITIONS)
    call(fl, pcounts, pcontacts, cap, total) = ccall((:ibvh_traverse_pair, LIB), Cint,
        (Ptr{Cvoid}, Ref{CBvh}, Ref{CBvh}, Ref{Params}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{Int64}, Ptr{Cvoid}),
        h, cq, ct, Params(sl_t, 0, -1, fl, flip ? 1 : 0, 0, C_NULL), pcounts, pcontacts, cap, total, stream())
    total, cache1, cache2 = two_phase(call, h, I, length(queries.leaves), isdefault(narrow) ? cache : nothing, flags)
    if !isdefault(narrow)           # positions come back as (position in bvh1.leaves, position in bvh2.leaves), flipped or not
        cache1 = apply_narrow(narrow, view(cache1, 1:total), bvh1.leaves, bvh2.leaves, false)
        total = length(cache1)
    end
    BVHTraversal(start_level1, start_level2, num_checks(h), total, cache1, cache2)
end

# ---- traverse_rays(bvh, points, directions, ::LVTTraversal; start_level, narrow, cache, options) ------------------
# raytrace/leaf_vs_tree/leaf_vs_tree.jl:1-90. points / directions: 3 x R CuMatrix (column-major == T xyz[R][3]).
function ImplicitBVH.traverse_rays(
    bvh::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    points::CuMatrix,
    directions::CuMatrix,
    alg::LVTTraversal;
    start_level::Int=1,
    narrow=nothing,
    cache::Union{Nothing, BVHTraversal}=nothing,
    options=BVHOptions(),
) where {I, N, L}
    size(points, 1) == size(directions, 1) == 3 || throw(ArgumentError("size(points, 1) == size(directions, 1) == 3 must hold"))
    size(points, 2) == size(directions, 2) || throw(ArgumentError("size(points, 2) == size(directions, 2) must hold"))
    bvh.built_level <= start_level <= bvh.tree.levels <= 32 ||
        throw(ArgumentError("bvh.built_level <= start_level <= bvh.tree.levels <= 32 must hold"))
    T = floattype(L.parameters[1])                                                  # the leaf float type (leaf_vs_tree.jl:116-125)
    nrays = size(points, 2)
    if nrays == 0                                                                   # leaf_vs_tree.jl:22-26
        return BVHTraversal(start_level, 0, 0, CuVector{IndexPair{I}}(undef, 0), CuVector{I}(undef, 0))
    end
    p = eltype(points) === T ? points : T.(points)
    d = eltype(directions) === T ? directions : T.(directions)
    h = handle()
    cb = cbvh(bvh, types(L, N))
    flags = isdefault(narrow) ? ORDERED : (ORDERED | POSITIONS)
    call(fl, pcounts, pcontacts, cap, total) = ccall((:ibvh_traverse_rays, LIB), Cint,
        (Ptr{Cvoid}, Ref{CBvh}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{Params}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{Int64}, Ptr{Cvoid}),
        h, cb, pointer(p), pointer(d), nrays, Params(start_level, 0, -1, fl, 0, 0, C_NULL), pcounts, pcontacts, cap, total, stream())
    total, cache1, cache2 = two_phase(call, h, I, nrays, isdefault(narrow) ? cache : nothing, flags)
    if !isdefault(narrow)           # narrow(leaf, point, direction) after a positive ray / leaf test (leaf_vs_tree.jl:194)
        hits = view(cache1, 1:total)                                                # (leaf position, ray id)
        keep = map(hp -> narrow(bvh.leaves[hp[1]], (p[1, hp[2]], p[2, hp[2]], p[3, hp[2]]), (d[1, hp[2]], d[2, hp[2]], d[3, hp[2]])), hits)
        cache1 = map(hp -> (bvh.leaves[hp[1]].index, hp[2]), hits[keep])
        total = length(cache1)
    end
    BVHTraversal(start_level, 0, num_checks(h), total, cache1, cache2)
end

# ---- BFSTraversal — src/traverse/breadth_first/*.jl, src/raytrace/breadth_first/*.jl ----------------------------------
# cache1 / cache2 are both Vector{IndexPair{I}} in the reference (its two BVTT buffers, traverse_single.jl:88-101); here the
# BVTT lives in library scratch, cache1 receives the contacts and cache2 is handed through. A cache1 that is too small is
# grown to the exact need and only the leaf level is repeated (COUNTS_VALID).
function bfs_protocol(call, h, ::Type{I}, cache, flags::UInt32) where I
    if !isnothing(cache)
        eltype(cache.cache1) === IndexPair{I} || throw(ArgumentError("eltype(cache.cache1) === IndexPair{I} must hold"))
        eltype(cache.cache2) === IndexPair{I} || throw(ArgumentError("eltype(cache.cache2) === IndexPair{I} must hold"))
    end
    cache1 = isnothing(cache) ? CuVector{IndexPair{I}}(undef, 0) : cache.cache1
    cache2 = isnothing(cache) ? CuVector{IndexPair{I}}(undef, 0) : cache.cache2
    total = Ref{Int64}(0); checks = Ref{Int64}(0)
    if length(cache1) > 0
        rc = call(flags, pointer(cache1), length(cache1), total, checks)
        rc == 0 && return Int(total[]), Int(checks[]), cache1, cache2
        rc == ERR_CAPACITY || check(rc, h)
    else
        check(call(flags, CU_NULL, 0, total, checks), h)                            # count only
        total[] == 0 && return 0, Int(checks[]), cache1, cache2
    end
    cache1 = CuVector{IndexPair{I}}(undef, Int(total[]))
    check(call(flags | COUNTS_VALID, pointer(cache1), length(cache1), total, checks), h)
    Int(total[]), Int(checks[]), cache1, cache2
end

ImplicitBVH.default_start_level(bvh::BVH{I, <:CuVector}, ::BFSTraversal) where I =
    Int(ccall((:ibvh_bfs_default_start_level, LIB), Int64, (Int64, Int64), bvh.tree.levels, bvh.built_level))

function ImplicitBVH.traverse(
    bvh::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    alg::BFSTraversal;
    start_level::Int=default_start_level(bvh, alg),
    narrow=nothing,
    cache::Union{Nothing, BVHTraversal}=nothing,
    options=BVHOptions(),
) where {I, N, L}
    bvh.tree.levels >= start_level >= bvh.built_level ||
        throw(ArgumentError("bvh.tree.levels >= start_level >= bvh.built_level must hold"))    # breadth_first/traverse_single.jl:10
    if bvh.tree.real_nodes <= 1
        return BVHTraversal(start_level, 0, 0, CuVector{IndexPair{I}}(undef, 0), CuVector{IndexPair{I}}(undef, 0))
    end
    h = handle()
    cb = cbvh(bvh, types(L, N))
    flags = isdefault(narrow) ? UInt32(0) : POSITIONS
    call(fl, pcontacts, cap, total, checks) = ccall((:ibvh_traverse_bfs_single, LIB), Cint,
        (Ptr{Cvoid}, Ref{CBvh}, Ref{Params}, CuPtr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}, Ptr{Cvoid}),
        h, cb, Params(start_level, 0, -1, fl, 0, 0, C_NULL), pcontacts, cap, total, checks, stream())
    total, checks, cache1, cache2 = bfs_protocol(call, h, I, isdefault(narrow) ? cache : nothing, flags)
    if !isdefault(narrow)
        cache1 = apply_narrow(narrow, view(cache1, 1:total), bvh.leaves, bvh.leaves, true)
        total = length(cache1)
    end
    BVHTraversal(start_level, checks, total, cache1, cache2)
end

function ImplicitBVH.traverse(
    bvh1::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    bvh2::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    alg::BFSTraversal;
    start_level1::Int=default_start_level(bvh1, alg),
    start_level2::Int=default_start_level(bvh2, alg),
    narrow=nothing,
    cache::Union{Nothing, BVHTraversal}=nothing,
    options=BVHOptions(),
) where {I, N, L}
    bvh1.tree.levels >= start_level1 >= bvh1.built_level ||
        throw(ArgumentError("bvh1.tree.levels >= start_level1 >= bvh1.built_level must hold"))  # breadth_first/traverse_pair.jl:10-11
    bvh2.tree.levels >= start_level2 >= bvh2.built_level ||
        throw(ArgumentError("bvh2.tree.levels >= start_level2 >= bvh2.built_level must hold"))
    h = handle()
    t = types(L, N)
    c1 = cbvh(bvh1, t)
    c2 = cbvh(bvh2, t)
    flags = isdefault(narrow) ? UInt32(0) : POSITIONS
    call(fl, pcontacts, cap, total, checks) = ccall((:ibvh_traverse_bfs_pair, LIB), Cint,
        (Ptr{Cvoid}, Ref{CBvh}, Ref{CBvh}, Int64, Int64, UInt32, CuPtr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}, Ptr{Cvoid}),
        h, c1, c2, start_level1, start_level2, fl, pcontacts, cap, total, checks, stream())
    total, checks, cache1, cache2 = bfs_protocol(call, h, I, isdefault(narrow) ? cache : nothing, flags)
    if !isdefault(narrow)
        cache1 = apply_narrow(narrow, view(cache1, 1:total), bvh1.leaves, bvh2.leaves, false)
        total = length(cache1)
    end
    BVHTraversal(start_level1, start_level2, checks, total, cache1, cache2)
end

function ImplicitBVH.traverse_rays(
    bvh::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}},
    points::CuMatrix,
    directions::CuMatrix,
    alg::BFSTraversal;
    start_level::Int=1,
    narrow=nothing,
    cache::Union{Nothing, BVHTraversal}=nothing,
    options=BVHOptions(),
) where {I, N, L}
    bvh.tree.levels >= start_level >= bvh.built_level ||
        throw(ArgumentError("bvh.tree.levels >= start_level >= bvh.built_level must hold"))    # raytrace/breadth_first/breadth_first.jl:12
    size(points, 1) == size(directions, 1) == 3 || throw(ArgumentError("size(points, 1) == size(directions, 1) == 3 must hold"))
    size(points, 2) == size(directions, 2) || throw(ArgumentError("size(points, 2) == size(directions, 2) must hold"))
    T = floattype(L.parameters[1])
    nrays = size(points, 2)
    if nrays == 0                                                                   # :25-29
        return BVHTraversal(start_level, 0, 0, CuVector{IndexPair{I}}(undef, 0), CuVector{IndexPair{I}}(undef, 0))
    end
    p = eltype(points) === T ? points : T.(points)
    d = eltype(directions) === T ? directions : T.(directions)
    h = handle()
    cb = cbvh(bvh, types(L, N))
    flags = isdefault(narrow) ? UInt32(0) : POSITIONS
    call(fl, pcontacts, cap, total, checks) = ccall((:ibvh_traverse_bfs_rays, LIB), Cint,
        (Ptr{Cvoid}, Ref{CBvh}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{Params}, CuPtr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}, Ptr{Cvoid}),
        h, cb, pointer(p), pointer(d), nrays, Params(start_level, 0, -1, fl, 0, 0, C_NULL), pcontacts, cap, total, checks, stream())
    total, checks, cache1, cache2 = bfs_protocol(call, h, I, isdefault(narrow) ? cache : nothing, flags)
    if !isdefault(narrow)
        hits = view(cache1, 1:total)
        keep = map(hp -> narrow(bvh.leaves[hp[1]], (p[1, hp[2]], p[2, hp[2]], p[3, hp[2]]), (d[1, hp[2]], d[2, hp[2]], d[3, hp[2]])), hits)
        cache1 = map(hp -> (bvh.leaves[hp[1]].index, hp[2]), hits[keep])
        total = length(cache1)
    end
    BVHTraversal(start_level, checks, total, cache1, cache2)
end

# ---- opt-in extension: sort (and deduplicate) a contact list on the device — what `sort(traversal.contacts)` does on the host in
# the reference's tests (test/gputests.jl:73-78); for the unordered emission modes, whose lists are sets (SURVEY.md §8f-2)
function sort_contacts!(t::BVHTraversal{<:CuVector{IndexPair{I}}}; unique::Bool=false) where I
    kept = Ref{Int64}(t.num_contacts)
    h = handle()
    check(ccall((:ibvh_sort_contacts, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, Int32, Cint, Ref{Int64}, Ptr{Cvoid}),
                h, pointer(t.cache1), t.num_contacts, sizeof(I), unique ? 1 : 0, kept, stream()), h)
    BVHTraversal(t.start_level1, t.start_level2, t.num_checks, Int(kept[]), t.cache1, t.cache2)
end

# ---- opt-in extension: asynchronous contact detection (IBVH_TRAVERSE_DEFER) ----------------------------------------
# `pending = traverse_deferred(bvh; cache)` only enqueues the (unordered) traversal; `BVH(next...; cache=other)` can be
# enqueued behind it; `finish!(pending)` waits for the count. At most one may be outstanding per device; a pending
# traversal that is dropped unread is cancelled by its finalizer so that later traversals are not refused.
mutable struct PendingTraversal{B, C1, C2}
    bvh::B; start_level::Int; cache1::C1; cache2::C2; handle::Ptr{Cvoid}; done::Bool
end
function traverse_deferred(bvh::BVH{I, <:CuVector, <:CuVector{N}, <:CuVector{L}}; start_level::Int=Int(max(1, bvh.built_level)),
                           cache::BVHTraversal) where {I, N, L}
    h = handle()
    cb = cbvh(bvh, types(L, N))
    total = Ref{Int64}(0)
    rc = ccall((:ibvh_traverse_single, LIB), Cint,
        (Ptr{Cvoid}, Ref{CBvh}, Ref{Params}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{Int64}, Ptr{Cvoid}),
        h, cb, Params(start_level, 0, -1, UNORDERED | DEFER, 0, 0, C_NULL), CU_NULL, pointer(cache.cache1), length(cache.cache1), total, stream())
    check(rc, h)
    p = PendingTraversal(bvh, start_level, cache.cache1, cache.cache2, h, total[] != -1)
    finalizer(x -> x.done || ccall((:ibvh_traverse_cancel, LIB), Cint, (Ptr{Cvoid},), x.handle), p)
    p
end
function finish!(p::PendingTraversal)
    total = Ref{Int64}(0)
    rc = p.done ? 0 : ccall((:ibvh_traverse_finish, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), p.handle, total)
    p.done = true
    if rc == ERR_AGAIN || rc == ERR_CAPACITY       # scratch lists / cache1 too small: repeat synchronously (the BVH must still be intact)
        return ImplicitBVH.traverse(p.bvh, LVTTraversal(); start_level=p.start_level,
                                    cache=BVHTraversal(p.start_level, 0, 0, p.cache1, p.cache2))
    end
    check(rc, p.handle)
    BVHTraversal(p.start_level, 0, num_checks(p.handle), Int(total[]), p.cache1, p.cache2)
end

end # module
