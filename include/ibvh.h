/* ibvh.h — C ABI of libibvh_b200.so: the B200-native (sm_100a) replacement for the data-parallel
 * hot path of ImplicitBVH.jl (Morton encode + scene bounds, Morton sort, bottom-up implicit-tree
 * merge, LVT and BFS single / pair / ray traversal).
 *
 * The reference has no FFI: its extension points are Julia multiple dispatch on
 * `MortonAlgorithm` (src/morton/morton.jl:4-15), `TraversalAlgorithm`
 * (src/traverse/traverse.jl:36,210-230; src/raytrace/raytrace.jl:71-81) and on the array type
 * (`AbstractGPUVector`, src/traverse/leaf_vs_tree/traverse_single.jl:24,93). A Julia package
 * extension specialising those methods on `CuVector` would `ccall` exactly the entry points below
 * (INTEGRATION.md shows the binding). Every entry point cites the reference code it replaces.
 *
 * Conventions
 *  - Plain pointers and sizes only. Every `d_*` pointer is DEVICE memory owned by the caller
 *    (CuArray / torch tensor); the library owns only the per-handle scratch workspace.
 *  - Layouts are the reference's isbits structs (natural alignment):
 *      BSphere{T}  { T x[3]; T r; }                       src/bounding_volumes/bsphere.jl:26-29
 *      BBox{T}     { T lo[3]; T up[3]; }                  src/bounding_volumes/bbox.jl:35-38
 *      BoundingVolume{V,I,M} { V volume; I index; M morton; }  bounding_volumes.jl:55-59
 *      IndexPair{I} { I a; I b; }                         src/traverse/traverse.jl:6
 *      points/directions: column-major 3 x R  ==  T xyz[R][3]    src/raytrace/raytrace.jl:12-13
 *  - Indices stored in leaves and reported in contacts are the caller's (1-based in Julia); ray ids
 *    are 1-based (raytrace/leaf_vs_tree/leaf_vs_tree.jl:127,157,200). Levels are 1-based.
 *    `query_begin` (sharding) is a 0-based offset.
 *  - All device work is enqueued on the caller's `stream` (a cudaStream_t passed as void*). Calls
 *    that must report a contact total synchronise that stream once, like the reference's scalar
 *    read-back (leaf_vs_tree/traverse_single.jl:60).
 *  - No exceptions cross the ABI: every call returns an ibvh_status. The Julia shim maps
 *    IBVH_ERR_ARGUMENT -> ArgumentError, IBVH_ERR_DOMAIN -> DomainError (build.jl:207,235,260,
 *    355-361; traverse_single.jl:9-11; implicit_tree.jl:78-80).
 *  - There is NO CPU fallback: without a CUDA device every device entry point returns
 *    IBVH_ERR_CUDA.
 */
#ifndef IBVH_H
#define IBVH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBVH_VERSION 100

#if defined(__GNUC__)
#define IBVH_API __attribute__((visibility("default")))
#else
#define IBVH_API
#endif

typedef enum ibvh_status {
    IBVH_OK = 0,
    IBVH_ERR_ARGUMENT = 1,     /* Julia ArgumentError / failed @argcheck                    */
    IBVH_ERR_DOMAIN = 2,       /* Julia DomainError (n < 1)                                 */
    IBVH_ERR_UNSUPPORTED = 3,  /* type combination not compiled into this build             */
    IBVH_ERR_CUDA = 4,         /* CUDA runtime error; text via ibvh_last_error              */
    IBVH_ERR_CAPACITY = 5,     /* contacts buffer too small; *num_contacts holds the need   */
    IBVH_ERR_ALLOC = 6,        /* workspace allocation failed                               */
    IBVH_ERR_PEER = 7,         /* a peer GPU did not arrive at the shard exchange in time   */
    IBVH_ERR_AGAIN = 8         /* deferred traversal: internal pair lists were too small — call the traversal again */
} ibvh_status;

typedef enum ibvh_volume_kind { IBVH_BSPHERE = 0, IBVH_BBOX = 1 } ibvh_volume_kind;

/* Static types of one BVH: BVH{I}(leaves::Vector{BoundingVolume{Leaf{T},I,M}}, nodes::Vector{Node{T}})
 * — build.jl:155-166, utils.jl:34-51 (index_exemplar, morton exemplar). */
typedef struct ibvh_types {
    int32_t leaf_kind;        /* ibvh_volume_kind of the leaf volumes                        */
    int32_t float_bytes;      /* 4 (Float32) or 8 (Float64): float type of the LEAF volumes  */
    int32_t index_bytes;      /* 4 (Int32) or 8 (Int64)   — BVHOptions.index_exemplar        */
    int32_t morton_bytes;     /* 2, 4, 8 (UInt16/32/64)   — DefaultMortonAlgorithm exemplar  */
    int32_t node_kind;        /* ibvh_volume_kind of the nodes (BBox leaves need BBox nodes) */
    int32_t node_float_bytes; /* float type of the NODE volumes: 0 = the leaf float type; 4 over 8-byte leaves = the
                                 reference's default call BVH(::Vector{BSphere{Float64}}) with BBox{Float32} nodes
                                 (build.jl:198-205, README.md:38-46; converting merges of merge.jl:47-81). Build and
                                 contact traversals only: the reference's own ray test needs one float type
                                 (isintersection.jl:1-5) */
} ibvh_types_t;

/* ImplicitTree{I} — implicit_tree.jl:52-67 */
typedef struct ibvh_tree {
    int64_t levels, real_leaves, real_nodes, virtual_leaves, virtual_nodes;
} ibvh_tree_t;

/* A built BVH as the kernels see it — build.jl:155-166 (skips are recomputed from n). */
typedef struct ibvh_bvh {
    const void* d_leaves;     /* BoundingVolume[n], Morton-sorted                            */
    const void* d_nodes;      /* Node[real_nodes - real_leaves]; levels < built_level unset  */
    int64_t n;                /* number of leaves                                            */
    int64_t built_level;      /* level up to which nodes are valid                           */
    ibvh_types_t types;
    uint64_t build_id;        /* 0, or the value ibvh_last_build_id returned right after the ibvh_build that produced
                                 these arrays: the library then reuses what that build left in its own memory for the
                                 traversal (the leaf volumes as aligned records, aligned copies of the node levels, the
                                 finest query-pyramid levels) instead of re-deriving it from the arrays on every call.
                                 Purely an accelerator: with 0, a stale id or arrays from elsewhere the traversal packs
                                 on the fly and returns the same result. The caller must not modify leaves / nodes
                                 between the build and a traversal that passes its id. */
} ibvh_bvh_t;

/* Traversal flags */
#define IBVH_TRAVERSE_ORDERED 0u     /* reference order: ascending query, then DFS order       */
#define IBVH_TRAVERSE_UNORDERED 1u   /* one pass, warp-aggregated atomic append (same set)     */
#define IBVH_TRAVERSE_REFERENCE_SHAPED 2u /* proxy of the reference GPU kernel: one thread per  */
                                          /* query, local stack, two passes (for comparison)   */
#define IBVH_TRAVERSE_PACKET 16u      /* force the warp-packet schedule (default for BSphere nodes); the */
                                      /* default for BBox nodes is the "group walk + dense tiles" one    */
#define IBVH_TRAVERSE_WALK 32u        /* force the "group walk + dense tiles" schedule instead of the     */
                                      /* default pyramid refinement (both BBox nodes only)               */
#define IBVH_TRAVERSE_DEFER 64u       /* single / pair, UNORDERED, default schedule: enqueue the whole traversal and  */
                                      /* return at once with *num_contacts = -1; ibvh_traverse_finish waits for it    */
                                      /* and reports the total. The host can enqueue the next build meanwhile, so the */
                                      /* GPU does not idle across the traversal's one host round trip.                */
#define IBVH_TRAVERSE_POSITIONS 128u  /* report 1-based POSITIONS in the (sorted) leaf arrays instead of the leaves' .index:   */
                                      /* single (query position, target position) with query < target; pair (position in the */
                                      /* queries' leaves, position in the target's leaves) (swapped under flip); rays (leaf    */
                                      /* position, ray id). Same pairs, same order. This is how a caller applies the         */
                                      /* reference's `narrow` predicate (traverse_single.jl:170, traverse_pair.jl:206,        */
                                      /* raytrace/leaf_vs_tree/leaf_vs_tree.jl:194: evaluated only AFTER a positive leaf test) */
                                      /* as a post-filter over leaves[position] with identical results.     */
#define IBVH_TRAVERSE_STATS 8u        /* fill the device counters read by ibvh_last_traversal_stats */
#define IBVH_TRAVERSE_COUNTS_VALID 4u /* ORDERED only: d_counts already holds the inclusive scan */
                                      /* left by a previous count-only call on the same queries: */
                                      /* skip the count pass and write (the reference's 2nd pass) */

/* Peer-memory descriptor of the multi-GPU exchange (see "multi-GPU" at the end of this header). */
#define IBVH_MAX_PEERS 16
typedef struct ibvh_peer {
    int32_t rank, world;                 /* world <= IBVH_MAX_PEERS                                  */
    uint64_t buffers[IBVH_MAX_PEERS];    /* device address of rank r's buffer as mapped in THIS process */
    uint64_t multicast;                  /* multicast alias of the buffers (NVLS), 0 = use peer stores */
    int64_t header_bytes;                /* >= 512 * 8, multiple of 256                               */
    int64_t capacity_bytes;              /* size of the list area that follows the header             */
    uint64_t epoch;                      /* collective-call counter shared by all ranks (> 0, increasing) */
    uint64_t fused_seq;                  /* 1, 2, 3, ... over the FUSED traversals issued on this buffer */
    int64_t region_begin[IBVH_MAX_PEERS + 1]; /* FUSED traversals: rank r appends its contacts to the entries
                                            [region_begin[r], region_begin[r+1]) of the list area (same values on every rank,
                                            e.g. the previous step's per-rank counts + 5 %); all zero = equal split */
} ibvh_peer_t;

typedef struct ibvh_traverse_params {
    int64_t start_level;      /* level of the descended tree the traversal starts from         */
    int64_t query_begin;      /* 0-based first query of this shard (leaf position / ray)       */
    int64_t query_count;      /* queries in this shard; < 0 means "all from query_begin"       */
    uint32_t flags;           /* IBVH_TRAVERSE_*                                               */
    int32_t flip;             /* pair only: emit (leaf.index, query.index) — traverse_pair.jl:212-216 */
    int64_t id_base;          /* rays only: ray r of the passed arrays is reported as id_base + r + 1
                                 (0 for a whole problem; the global offset of a per-GPU ray shard)     */
    const ibvh_peer_t* peer;  /* NULL, or (single / pair with BBox nodes, or rays; UNORDERED): fused traversal + all-gather —
                                 every rank traverses its query shard and the contacts of ALL ranks land in
                                 EVERY rank's list area (peer->buffers[r] + header_bytes) while the traversal
                                 runs: rank r reserves slots inside its own region (peer->region_begin) with local
                                 atomics and writes them with multimem.st through peer->multicast; the ranks exchange
                                 their counts at the end and the gaps between the regions are closed (entries beyond
                                 the total move into them), so the list area holds *num_contacts contiguous pairs,
                                 the same on every rank. d_contacts is ignored. ibvh_peer_last_counts gives the
                                 per-rank counts (to size the next call's regions). IBVH_ERR_CAPACITY: some rank's
                                 region was too small (nothing valid; *num_contacts = total need). Collective;
                                 IBVH_ERR_UNSUPPORTED if the combination cannot run fused (then: traverse locally +
                                 ibvh_allgather_pairs). */
} ibvh_traverse_params_t;

typedef struct ibvh_handle ibvh_handle_t;

/* ---- library ---------------------------------------------------------------------------- */
IBVH_API int ibvh_version(void);
IBVH_API const char* ibvh_status_string(int status);
IBVH_API const char* ibvh_last_error(const ibvh_handle_t* h);

/* ---- host-only integer math (replaces implicit_tree.jl:77-199, build.jl:309-325) ---------- */
/* ImplicitTree{I}(n) + compute_skips!: `skips` receives tree->levels entries (may be NULL).  */
IBVH_API int ibvh_tree_shape(int64_t n, ibvh_tree_t* tree, int64_t* skips);
IBVH_API int64_t ibvh_memory_index(const ibvh_tree_t* tree, int64_t implicit_index);
IBVH_API int ibvh_level_indices(const ibvh_tree_t* tree, int64_t level, int64_t* start, int64_t* stop);
IBVH_API int ibvh_isvirtual(const ibvh_tree_t* tree, int64_t implicit_index);
/* compute_build_level: is_float ? round(levels + (1-levels)*f) : check 1 <= ilevel <= levels   */
IBVH_API int ibvh_compute_build_level(int64_t levels, int is_float, int64_t ilevel, double flevel, int64_t* out);
/* sizeof(BoundingVolume{...}) / sizeof(Node) / sizeof(IndexPair{I}) for a type set; <0 if unsupported */
IBVH_API int64_t ibvh_leaf_bytes(const ibvh_types_t* types);
IBVH_API int64_t ibvh_volume_bytes(int32_t kind, int32_t float_bytes);
IBVH_API int64_t ibvh_num_nodes(int64_t n);   /* real_nodes - real_leaves, build.jl:256 */

/* ---- handle / workspace (replaces AK's temporary allocations) -----------------------------
 * Memory the library owns per handle (all grow-only, freed by ibvh_release_workspace / ibvh_destroy):
 *   build scratch     ~40 B / leaf for 32-bit keys (ibvh_workspace_query gives the exact figure): radix keys x2,
 *                     permutation x2, look-back words, and a copy of the leaves on the in-place path;
 *   sidecars          two slots of ~30 B / leaf (16-byte volume records, aligned node levels, pyramid levels, index array)
 *                     left by the two most recent pyramid-eligible builds (see ibvh_bvh_t.build_id);
 *   BFS lists         the two ping-pong BVTT lists of the BFS traversals (8 bytes per entry; the destination of a level is
 *                     sized 4x / 2x its source like the reference's bvtt2, traverse_single.jl:42);
 *   traversal scratch pair lists + quantised boxes (~70 B / leaf at 10 M uniformly random spheres; grows with the
 *                     density of the scene) and, for ORDERED traversals with a contacts buffer, a 16-byte-per-contact hit
 *                     stash sized from the contact totals this handle has SEEN (last total + 25 %, or 6 per query before the
 *                     first), never from the caller's `capacity`: a generously pre-sized cache1 costs no extra scratch. */
IBVH_API int ibvh_create(ibvh_handle_t** out, int device);
IBVH_API int ibvh_destroy(ibvh_handle_t* h);
/* Bytes of scratch a build of n leaves will hold on to (sort ping-pong, keys, look-back).    */
IBVH_API int64_t ibvh_workspace_query(const ibvh_types_t* types, int64_t n);
IBVH_API int64_t ibvh_workspace_bytes(const ibvh_handle_t* h);   /* currently allocated               */
IBVH_API int ibvh_release_workspace(ibvh_handle_t* h);

/* ---- build stages (each separately callable for parity tests and ncu) --------------------- */
/* wrap_bounding_volumes, build.jl:328-352: leaves[i] = BoundingVolume(volumes[i], I(i+1), M(0)) */
IBVH_API int ibvh_wrap(ibvh_handle_t* h, const void* d_volumes, int64_t n, const ibvh_types_t* types,
              void* d_leaves, void* stream);

/* Leaf volumes from triangles — BSphere{T}(p1, p2, p3), src/bounding_volumes/bsphere.jl:43-112, and
 * BBox{T}(p1, p2, p3), bbox.jl:59-70: the step right before the hot path (README "compute bounding
 * volumes"). d_triangles: T[n][3][3]; d_volumes: BSphere{T}[n] or BBox{T}[n]. */
IBVH_API int ibvh_volumes_from_triangles(ibvh_handle_t* h, const void* d_triangles, int64_t n, int32_t kind, int32_t float_bytes,
                                void* d_volumes, void* stream);

/* morton_encode!, morton/default.jl:43-82 + bounding_volumes_extrema, morton/utils.jl:1-72.
 * compute_extrema != 0: scene bounds reduced on device and padded exactly like the reference;
 * otherwise `mins/maxs` (3 doubles each, converted to the leaf float type) are used unpadded.
 * out_mins/out_maxs (host, 3 doubles each, may be NULL): the bounds used; non-NULL forces a sync. */
IBVH_API int ibvh_morton_encode(ibvh_handle_t* h, void* d_leaves, int64_t n, const ibvh_types_t* types,
                       int compute_extrema, const double* mins, const double* maxs,
                       double* out_mins, double* out_maxs, void* stream);

/* AK.sort!(leaves, by = morton), build.jl:248-253: stable ascending, in place. */
IBVH_API int ibvh_sort_leaves(ibvh_handle_t* h, void* d_leaves, int64_t n, const ibvh_types_t* types, void* stream);

/* aggregate_oibvh!, build.jl:366-523: nodes for levels built_level .. levels-1. */
IBVH_API int ibvh_aggregate(ibvh_handle_t* h, const void* d_leaves, int64_t n, const ibvh_types_t* types,
                   void* d_nodes, int64_t built_level, void* stream);

/* BVH(bounding_volumes, node_type; built_level, cache, options), build.jl:198-271 — the fused
 * pipeline: bounds+encode (+digit histograms) -> onesweep radix sort of (key, perm) -> gather of the
 * sorted leaves back into d_leaves fused with the bottom levels of the merge -> upper levels.
 * d_volumes != NULL: wrap path (raw volumes in, wrapped sorted leaves out; build.jl:220-225);
 * d_volumes == NULL: d_leaves already holds BoundingVolume structs and is sorted in place. */
IBVH_API int ibvh_build(ibvh_handle_t* h, const void* d_volumes, void* d_leaves, int64_t n,
               const ibvh_types_t* types, void* d_nodes, int64_t built_level,
               int compute_extrema, const double* mins, const double* maxs, void* stream);

/* Reference-SHAPED proxy of the same constructor (measurement aid, NOT the product path): launches what BVH(...) launches
 * through KernelAbstractions / AcceleratedKernels on a GPU — wrap (build.jl:340), two mapreduce passes with a scalar
 * read-back each (morton/utils.jl:24-44), a whole-struct encode pass (default.jl:66), a comparison merge sort that moves
 * the 24-byte structs (build.jl:248-253) and one merge launch per tree level (build.jl:413,492). Same results, bit for
 * bit. Stands in for the reference's CUDA.jl backend, which cannot run where there is no Julia; pair it with
 * IBVH_TRAVERSE_REFERENCE_SHAPED. BSphere{Float32} / Int32 / UInt32 / BBox{Float32} only (else IBVH_ERR_UNSUPPORTED). */
IBVH_API int ibvh_build_reference_shaped(ibvh_handle_t* h, const void* d_volumes, void* d_leaves, int64_t n,
                                const ibvh_types_t* types, void* d_nodes, int64_t built_level, void* stream);

/* Id of the sidecar the last successful ibvh_build on this handle left behind (0 = none: not a pyramid-eligible tree,
 * or sidecars disabled). Store it with the BVH and pass it in ibvh_bvh_t.build_id. The handle keeps the two newest. */
IBVH_API uint64_t ibvh_last_build_id(ibvh_handle_t* h);

/* ---- LVT traversals ------------------------------------------------------------------------ */
/* Common output protocol (replaces the count -> accumulate -> allocate -> write sequence of
 * leaf_vs_tree/traverse_single.jl:52-78):
 *   d_counts   : I[query_count] or NULL. ORDERED mode fills it with the inclusive scan of
 *                per-query contact counts (the GPU form of BVHTraversal.cache2).
 *   d_contacts : IndexPair{I}[capacity] or NULL (NULL = count only). Aligned to sizeof(IndexPair{I}) = 2 * index_bytes
 *                (any element of a cudaMalloc'ed / CuArray allocation is); a misaligned pointer is IBVH_ERR_ARGUMENT.
 *   num_contacts (host): total found. If it exceeds `capacity` the call returns
 *                IBVH_ERR_CAPACITY and the caller grows cache1 and calls again
 *                ("resize only if too small", traverse_single.jl:61-67). */

/* traverse(bvh, LVTTraversal(); start_level), leaf_vs_tree/traverse_single.jl:1-208 */
IBVH_API int ibvh_traverse_single(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const ibvh_traverse_params_t* params,
                         void* d_counts, void* d_contacts, int64_t capacity, int64_t* num_contacts,
                         void* stream);

/* traverse_lvt(bvh1, bvh2, ...), leaf_vs_tree/traverse_pair.jl:40-244: `queries` is the tree whose
 * leaves are the query set (the reference picks the one with more leaves and sets flip, :16-36 —
 * that choice is made by the caller), `target` is the tree descended from params->start_level. */
IBVH_API int ibvh_traverse_pair(ibvh_handle_t* h, const ibvh_bvh_t* queries, const ibvh_bvh_t* target,
                       const ibvh_traverse_params_t* params, void* d_counts, void* d_contacts,
                       int64_t capacity, int64_t* num_contacts, void* stream);

/* traverse_rays(bvh, points, directions, LVTTraversal()), raytrace/leaf_vs_tree/leaf_vs_tree.jl:1-228.
 * d_points / d_directions: T[nrays][3] in the BVH float type. Emits (leaf.index, id_base + r + 1), r 0-based. */
IBVH_API int ibvh_traverse_rays(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const void* d_points, const void* d_directions,
                       int64_t nrays, const ibvh_traverse_params_t* params, void* d_counts, void* d_contacts,
                       int64_t capacity, int64_t* num_contacts, void* stream);

/* ---- BFS traversals ------------------------------------------------------------------------ */
/* BFSTraversal (src/traverse/traverse.jl:19-24): simultaneous breadth-first descent, node pairs level by level. The
 * BVTT lists (8 bytes per entry, 10-20x the contacts at their widest) live in two grow-only buffers the library owns;
 * `d_contacts` only ever receives contacts. Like the reference's GPU backend (atomic appends, traverse_single_gpu.jl:
 * 84-101) the contacts come back as a SET in unspecified order; *num_checks (may be NULL) is BVHTraversal.num_checks =
 * the sum of the BVTT list lengths over the levels (traverse_single.jl:28,51), which is deterministic.
 * Output protocol: d_contacts == NULL counts only; more contacts than `capacity` -> IBVH_ERR_CAPACITY with the need in
 * *num_contacts. Either way the library keeps the leaf-level list: repeating the SAME call with a large enough buffer and
 * IBVH_TRAVERSE_COUNTS_VALID only redoes the leaf level (the repeat must be the next BFS call on the handle, with the trees
 * untouched in between; anything else — other arguments, no kept list — simply runs the whole traversal again). Other flags
 * honoured: IBVH_TRAVERSE_POSITIONS.
 * One host read-back per level (the reference reads dst_offsets[level] the same way). */
/* default_start_level(bvh, ::BFSTraversal) = max(levels / 2, built_level), breadth_first/breadth_first.jl:4-6 */
IBVH_API int64_t ibvh_bfs_default_start_level(int64_t levels, int64_t built_level);

/* traverse(bvh, BFSTraversal(); start_level), breadth_first/traverse_single.jl:1-157 + traverse_single_gpu.jl:
 * every pair of the real nodes of params->start_level (self pairs included) is descended; contacts are emitted as
 * (smaller index, larger index). params: start_level and flags are read. */
IBVH_API int ibvh_traverse_bfs_single(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const ibvh_traverse_params_t* params,
                             void* d_contacts, int64_t capacity, int64_t* num_contacts, int64_t* num_checks, void* stream);

/* traverse(bvh1, bvh2, BFSTraversal(); start_level1, start_level2), breadth_first/traverse_pair.jl:1-221 +
 * traverse_pair_gpu.jl: both trees descend in step, then the deeper one alone, down to leaf pairs; contacts are
 * (index in bvh1, index in bvh2). Both BVHs must share leaf / index / Morton / node types. */
IBVH_API int ibvh_traverse_bfs_pair(ibvh_handle_t* h, const ibvh_bvh_t* bvh1, const ibvh_bvh_t* bvh2, int64_t start_level1,
                           int64_t start_level2, uint32_t flags, void* d_contacts, int64_t capacity, int64_t* num_contacts,
                           int64_t* num_checks, void* stream);

/* traverse_rays(bvh, points, directions, BFSTraversal(); start_level), raytrace/breadth_first/breadth_first.jl:1-140 +
 * raytrace_gpu.jl: entries are (node, ray); emits (leaf.index, id_base + r + 1). nrays < 2^32.
 * params: start_level, flags and id_base are read. */
IBVH_API int ibvh_traverse_bfs_rays(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const void* d_points, const void* d_directions,
                           int64_t nrays, const ibvh_traverse_params_t* params, void* d_contacts, int64_t capacity,
                           int64_t* num_contacts, int64_t* num_checks, void* stream);

/* ---- sorted / unique contact lists (opt-in post-processing on the device) ---------
 * Sorts the `count` pairs of a contact list ascending by (a, b) — the order in which the reference's tests compare lists
 * (`sort(traversal.contacts)`, test/gputests.jl:73-78, runtests.jl:1246-1250) — in place, with the library's onesweep radix
 * sort over the keys a << 32 | b; `unique != 0` also drops repeated pairs (*out_count = pairs kept; the tail of the
 * buffer is unspecified). Meant for the unordered emission modes (IBVH_TRAVERSE_UNORDERED, the BFS traversals), whose
 * lists are sets. Indices must lie in [0, 2^31) x [0, 2^32) — true for the 1..n that wrap_bounding_volumes assigns —
 * else IBVH_ERR_UNSUPPORTED and the list is untouched. count <= 2^31. Synchronises the stream. */
IBVH_API int ibvh_sort_contacts(ibvh_handle_t* h, void* d_contacts, int64_t count, int32_t index_bytes, int unique,
                       int64_t* out_count, void* stream);

/* ---- per-kernel timing (measurement aid; bench.py's roofline uses it) ----------------------
 * When enabled, every kernel launch of this handle is bracketed by CUDA events on the launching
 * stream. ibvh_profile_get(i) synchronises event i and returns the kernel family name and its
 * duration in ms; entries are in launch order since the last enable / reset. */
IBVH_API int ibvh_profile_enable(ibvh_handle_t* h, int on);
IBVH_API int ibvh_profile_count(ibvh_handle_t* h);
IBVH_API int ibvh_profile_get(ibvh_handle_t* h, int i, char* name, int name_cap, float* ms);
IBVH_API int ibvh_profile_reset(ibvh_handle_t* h);

/* Device counters of the last ORDERED traversal run with IBVH_TRAVERSE_STATS on this handle (count
 * pass of the default BSphere{Float32}/Int32/UInt32/BBox type set only):
 * out[0] = node tests (per query), out[1] = leaf tests (per query),
 * packet schedule: out[2] = warp steps, out[3] = warp-uniform node/leaf loads;
 * reference-shaped schedule: out[2] = per-query steps, out[3] = sum over warps of the slowest lane's steps.
 * Pyramid schedule (the default for BBox nodes; no flag needed, any mode): totals of the last traversal derived
 * from its pair-list sizes — out[0] = box-box tests, out[1] = leaf-leaf tests, out[2] = candidate pairs of
 * 4-leaf groups, out[3] = pyramid levels. */
IBVH_API int ibvh_last_traversal_stats(ibvh_handle_t* h, int64_t out[4]);

/* Completes the traversal started with IBVH_TRAVERSE_DEFER on this handle (at most one may be outstanding; work
 * enqueued on the stream AFTER it, e.g. the next ibvh_build, is not waited for). *num_contacts = the total.
 * IBVH_ERR_CAPACITY as for the synchronous call; IBVH_ERR_AGAIN: the scratch pair lists overflowed (their size is
 * learned from the previous call) — nothing valid was written, repeat the traversal (it will size them right). */
IBVH_API int ibvh_traverse_finish(ibvh_handle_t* h, int64_t* num_contacts);
/* Drops the outstanding deferred traversal of this handle without judging its result (waits for the enqueued work,
 * since it still writes the handle's read-back slots). No-op when nothing is outstanding. A caller that loses
 * interest in a deferred result (the Julia finalizer of a BVHTraversal never read) calls this, so that later
 * traversals on the handle are not refused. */
IBVH_API int ibvh_traverse_cancel(ibvh_handle_t* h);

/* ---- multi-GPU: all-gather of the contact / hit shards over NVLink peer memory ------
 * The reference has no multi-GPU path; this is the exchange step that follows a query-range sharded
 * traversal (traverse_single.jl:157, traverse_pair.jl:199, raytrace/leaf_vs_tree/leaf_vs_tree.jl:137:
 * every query is independent). One process per GPU. Each process owns one "symmetric" buffer of the
 * same size and hands this library the device addresses at which ALL ranks' buffers are mapped into its
 * own address space (CUDA VMM / IPC peer mappings; the Python stand-in gets them from
 * torch.distributed._symmetric_memory) and, when the fabric offers it, the NVSwitch multicast alias.
 * Buffer layout: [header_bytes of signal slots, zeroed once at creation][list area].
 *
 * ibvh_allgather_pairs: ONE kernel per rank that (1) publishes this rank's pair count into every peer's
 * header and waits for theirs, (2) writes its shard at its rank-order offset into EVERY rank's list area
 * (multimem.st through the multicast alias, else plain peer stores), (3) runs a release/acquire barrier
 * over the headers, so that when the kernel completes this rank's list area holds the concatenation of
 * all shards in rank order — for ORDERED shards of contiguous query ranges exactly the single-GPU list.
 * Collective: every rank must call it with the same epoch (> 0, increasing from call to call, shared with the
 * fused traversals' peer->epoch) and
 * stream-order its readers of the previous list before the call. A rank that does not arrive within
 * 10 s makes the others return IBVH_ERR_PEER instead of hanging.
 * Host outputs (the call synchronises the stream): *out_total = pairs in the gathered list,
 * *out_offset = first pair of this rank's shard inside it. IBVH_ERR_CAPACITY if the list area is too
 * small for out_total pairs (same verdict on every rank; nothing is written). */
/* Per-rank contact / hit counts of the last FUSED traversal on this handle (world entries). */
IBVH_API int ibvh_peer_last_counts(ibvh_handle_t* h, int64_t* counts, int32_t world);

/* Host-only: the moves (in list entries) with which a fused traversal closes the gaps of its segmented list — rank r holds
 * counts[r] entries from region_begin[r] on; entries at positions >= sum(counts) move into the uncovered positions below
 * it. Returns the number of (src, dst, len) moves written (<= max_moves), -1 on bad arguments. For testing the host logic. */
IBVH_API int ibvh_peer_compact_plan(int32_t world, const int64_t* region_begin, const int64_t* counts,
                                    int64_t* src, int64_t* dst, int64_t* len, int32_t max_moves);

IBVH_API int ibvh_allgather_pairs(ibvh_handle_t* h, const ibvh_peer_t* peer, const void* d_shard, int64_t count,
                                  int32_t pair_bytes, int64_t* out_total, int64_t* out_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IBVH_H */
