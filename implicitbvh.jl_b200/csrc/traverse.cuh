// traverse.cuh — leaf-vs-tree (LVT) traversal kernels: single tree, BVH-vs-BVH and rays.
// Replaces traverse_lvt_single! (src/traverse/leaf_vs_tree/traverse_single.jl:136-208),
// traverse_lvt_pair! (traverse_pair.jl:176-244) and traverse_ray_lvt!
// (src/raytrace/leaf_vs_tree/leaf_vs_tree.jl:170-228) and their two-pass drivers.
//
// Exactness rule shared by every kernel here: query q reports leaf j iff the leaf-level predicate
// holds AND every ancestor of j from start_level down passes the node-level predicate for q (and,
// for single-tree traversal, j lies strictly to the right of q). Any schedule that evaluates exactly
// this predicate produces the reference's contact set bit for bit; per query the hits come out in
// ascending j (DFS left-first order), which is also the reference's order.
//
// Two schedules are provided:
//  * lvt_thread_kernel — "reference-shaped": one thread per query, private stack, one node per
//    step. Used for rays (incoherent queries) and as the comparison proxy of the reference's GPU
//    kernel (IBVH_TRAVERSE_REFERENCE_SHAPED).
//  * lvt_packet_kernel — B200 schedule for leaf queries: a warp owns 32 Morton-consecutive query
//    leaves and walks the tree ONCE for all of them with a warp-uniform (node, lane-mask) stack in
//    shared memory. Node loads are warp-uniform broadcasts (one sector instead of 32 scattered
//    ones), control flow never diverges, and the per-lane mask keeps the ancestor predicate exact.
#pragma once
#include "common.cuh"
#include "peer.cuh"

namespace ibvh {

enum TraverseKind { kSingle = 0, kPair = 1, kRays = 2 };
enum EmitMode { kCount = 0, kWrite = 1, kAtomic = 2 };

template <class L, class N> struct DBvh {
    const L* leaves;
    const N* nodes;
    TreeInfo ti;
};

struct TraverseArgs {
    int64_t q_begin;        // first query (0-based) of this shard
    int64_t q_count;        // number of queries
    int32_t start_level;
    int32_t flip;
    int64_t capacity;       // contacts capacity (pairs)
    int64_t id_base;        // rays: reported id = id_base + q + 1
    unsigned long long* total;   // atomic mode: running total
    unsigned long long* stats;   // optional counters (node tests, leaf tests, steps) or nullptr
    // optional indirection (fallback of the tiled schedule): the queries are the members of the query
    // groups listed in qmap[0 .. *qmap_count), qmap_group leaves per group; grid-stride over them
    const uint32_t* qmap;
    const uint32_t* qmap_count;
    int32_t qmap_group;
    const ibvh_peer_t* peer;     // fused traversal + all-gather over peer memory (pyramid schedule, unordered), else nullptr
    int32_t positions;           // IBVH_TRAVERSE_POSITIONS: report 1-based leaf POSITIONS in the sorted arrays instead of .index
    unsigned long long t_build_id, q_build_id;   // ibvh_bvh_t.build_id of the target / query tree (0 = none): selects a build's sidecar
    int64_t t_built_level;       // built_level of the target tree
    void* wq_data;               // rays: queue of long rays for rays_wide_kernel (RayWideEntry[wq_cap], workspace) or nullptr
    uint32_t wq_cap, wq_after;   //       its capacity; node steps after which a lane exports its ray
};

// 8-byte vectorised struct loads (volumes are 8-byte aligned by layout; see common.cuh)
template <class S> IBVH_D S load_struct(const S* p) {
    static_assert(sizeof(S) % 8 == 0, "8-byte multiple");
    alignas(8) S out;
    const uint2* s = reinterpret_cast<const uint2*>(p);
    uint2* d = reinterpret_cast<uint2*>(&out);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(S) / 8); ++k) d[k] = __ldg(s + k);
    return out;
}

IBVH_D bool tree_isvirtual(const TreeInfo& ti, uint32_t idx, int level) {     // implicit_tree.jl:191-199
    return (int64_t)(idx - (1u << (level - 1))) + 1 > ti.level_nreal[level];
}

// ---- emission ------------------------------------------------------------------------------------------
template <class I, int MODE> struct Emitter {
    IndexPair<I>* contacts;
    int64_t pos;            // kWrite: next slot; kCount: running count
    int64_t capacity;
    unsigned long long* total;
    IBVH_D void emit(I a, I b) {
        if constexpr (MODE == kCount) { pos += 1; }
        else if constexpr (MODE == kWrite) { contacts[pos] = IndexPair<I>{a, b}; pos += 1; }
        else {
            // warp-aggregated atomic append among the lanes that are emitting right now
            unsigned m = __activemask();
            int lane = threadIdx.x & 31;
            int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(total, (unsigned long long)__popc(m));
            base = __shfl_sync(m, base, leader);
            unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
            if ((int64_t)slot < capacity) contacts[slot] = IndexPair<I>{a, b};
        }
    }
};

// ---- query state ------------------------------------------------------------------------------------------
template <int KIND, class LQ, class N> struct QueryState;
// leaf queries (single / pair): the leaf volume, its node-typed twin (traverse_single.jl:154-155), index
template <class LQ, class N> struct QueryLeaf {
    typename LQ::vol_t vol;
    N bvn;
    typename LQ::idx_t index;
};
template <class T> struct QueryRay { T p[3]; T d[3]; };

// ===========================================================================================================
// Reference-shaped schedule: one thread per query
// ===========================================================================================================
template <int KIND, int MODE, class LQ, class LT, class N, class I, bool STATS = false>
__global__ void __launch_bounds__(128) lvt_thread_kernel(const LQ* __restrict__ qleaves,
                                                        const typename LT::value_type* __restrict__ points,
                                                        const typename LT::value_type* __restrict__ dirs,
                                                        DBvh<LT, N> bvh, TraverseArgs a,
                                                        I* counts, IndexPair<I>* contacts) {
    using T = typename LT::value_type;
    const TreeInfo& ti = bvh.ti;
    const int levels = ti.levels;
    const int64_t t_stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t_limit = a.qmap ? (int64_t)(*a.qmap_count) * a.qmap_group : a.q_count;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < t_limit; t += t_stride) {
    int64_t qi = t;                                                       // query within the shard
    if (a.qmap) qi = (int64_t)a.qmap[t / a.qmap_group] * a.qmap_group + (t % a.qmap_group);
    if (qi >= a.q_count) continue;
    const int64_t q = a.q_begin + qi;

    QueryLeaf<LQ, N> ql;
    QueryRay<T> qr;
    if constexpr (KIND == kRays) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { qr.p[k] = points[3 * q + k]; qr.d[k] = dirs[3 * q + k]; }
    } else {
        LQ leaf = load_struct(qleaves + q);
        ql.vol = leaf.volume;
        ql.index = a.positions ? (decltype(leaf.index))(q + 1) : leaf.index;
        ql.bvn = NodeOps<N>::convert(leaf.volume);
    }

    Emitter<I, MODE> em;
    em.contacts = contacts;
    em.capacity = a.capacity;
    em.total = a.total;
    em.pos = 0;
    if constexpr (MODE == kWrite) em.pos = (qi == 0) ? 0 : (int64_t)counts[qi - 1];

    const uint32_t inode_start = 1u << (a.start_level - 1);
    const uint32_t inode_end = inode_start + (uint32_t)ti.level_nreal[a.start_level] - 1u;
    const uint64_t q_impl = (uint64_t)q + (uint64_t(1) << (levels - 1));      // implicit index of the query leaf
    uint32_t stack[32];
    unsigned long long st_node = 0, st_leaf = 0, st_steps = 0;

    for (uint32_t root = inode_start; root <= inode_end; ++root) {
        int sp = 0;
        uint32_t inode = root;
        while (true) {
            const int level = 32 - __clz(inode);
            bool descend = false;
            bool skip = false;
            if constexpr (STATS) st_steps += 1;
            if constexpr (KIND == kSingle) {
                // traverse_single.jl:165-167 — subtree entirely at or left of the query leaf
                uint64_t rightmost = (((uint64_t)inode + 1u) << (levels - level)) - 1u;
                skip = rightmost <= q_impl;
            }
            if (!skip) {
                if (level == levels) {
                    if constexpr (STATS) st_leaf += 1;
                    LT leaf = load_struct(bvh.leaves + (inode - (1u << (levels - 1))));
                    if (a.positions) leaf.index = (decltype(leaf.index))(inode - (1u << (levels - 1)) + 1u);
                    if constexpr (KIND == kRays) {
                        if (isintersection(leaf.volume, qr.p, qr.d)) em.emit((I)leaf.index, (I)(a.id_base + q + 1));
                    } else {
                        if (iscontact(ql.vol, leaf.volume)) {
                            if constexpr (KIND == kSingle) {
                                if (ql.index > leaf.index) em.emit((I)leaf.index, (I)ql.index);
                                else em.emit((I)ql.index, (I)leaf.index);
                            } else {
                                if (a.flip) em.emit((I)leaf.index, (I)ql.index);
                                else em.emit((I)ql.index, (I)leaf.index);
                            }
                        }
                    }
                } else {
                    if constexpr (STATS) st_node += 1;
                    N node = load_struct(bvh.nodes + ((int64_t)inode - ti.skips[level] - 1));
                    bool hit;
                    if constexpr (KIND == kRays) hit = isintersection(node, qr.p, qr.d);
                    else hit = iscontact(ql.bvn, node);
                    if (hit) {
                        uint32_t right = 2u * inode + 1u;
                        if (!tree_isvirtual(ti, right, level + 1)) stack[sp++] = right;
                        inode = 2u * inode;
                        descend = true;
                    }
                }
            }
            if (descend) continue;
            if (sp == 0) break;
            inode = stack[--sp];
        }
    }
    if constexpr (MODE == kCount) counts[qi] = (I)em.pos;
    if constexpr (STATS) {
        if (a.stats) {
            atomicAdd(a.stats + 0, st_node);
            atomicAdd(a.stats + 1, st_leaf);
            atomicAdd(a.stats + 2, st_steps);
            unsigned m = __activemask();
            unsigned long long mx = st_steps;
            for (int off = 16; off > 0; off >>= 1) { unsigned long long o = __shfl_xor_sync(m, mx, off); mx = o > mx ? o : mx; }
            if ((threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(a.stats + 3, mx);     // sum over warps of the slowest lane's steps
        }
    }
    }   // grid-stride loop over queries
}

// ===========================================================================================================
// Ray schedule: one thread per ray, stackless, two children per step
// ===========================================================================================================
// Same predicate and same hit order as traverse_ray_lvt! (raytrace/leaf_vs_tree/leaf_vs_tree.jl:170-228): a leaf
// is reported iff its own test and the slab test of every ancestor from start_level down pass, hits come
// out left to right. Differences are only in the schedule: (1) the two children of a node are adjacent in
// memory (48 contiguous bytes) and are tested in the same step, which halves the dependent-load chain;
// (2) the implicit tree makes the stack redundant: a 32-bit "pending right sibling" mask indexed by level
// replaces it (the ancestor at level k of node i at level l is i >> (l - k)), so there is no local memory.
template <int MODE, class LT, class N, class I>
__global__ void __launch_bounds__(128) rays_kernel(const typename LT::value_type* __restrict__ points,
                                                  const typename LT::value_type* __restrict__ dirs,
                                                  DBvh<LT, N> bvh, TraverseArgs a, I* counts, IndexPair<I>* contacts) {
    using T = typename LT::value_type;
    using V = typename LT::vol_t;
    __shared__ uint32_t s_skip[34];
    __shared__ uint32_t s_nreal[34];
    for (int i = threadIdx.x; i < 34; i += blockDim.x) { s_skip[i] = (uint32_t)bvh.ti.skips[i]; s_nreal[i] = (uint32_t)bvh.ti.level_nreal[i]; }
    __syncthreads();
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= a.q_count) return;
    const int64_t q = a.q_begin + qi;
    const int levels = bvh.ti.levels;
    T p[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] = points[3 * q + k]; d[k] = dirs[3 * q + k]; }
    const I ray_id = (I)(a.id_base + q + 1);

    Emitter<I, MODE> em;
    em.contacts = contacts;
    em.capacity = a.capacity;
    em.total = a.total;
    em.pos = 0;
    if constexpr (MODE == kWrite) em.pos = (qi == 0) ? 0 : (int64_t)counts[qi - 1];

    const uint32_t leaf0 = 1u << (levels - 1);
    auto test_leaf = [&](uint32_t inode) {
        const LT* lp = bvh.leaves + (inode - leaf0);
        V v;
        const uint2* sp = reinterpret_cast<const uint2*>(lp);
        uint2* dp = reinterpret_cast<uint2*>(&v);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / 8); ++k) dp[k] = __ldg(sp + k);
        if (isintersection(v, p, d)) em.emit(a.positions ? (I)(inode - leaf0 + 1u) : (I)lp->index, ray_id);
    };

    const uint32_t inode_start = 1u << (a.start_level - 1);
    const uint32_t inode_end = inode_start + s_nreal[a.start_level] - 1u;
    for (uint32_t root = inode_start; root <= inode_end; ++root) {
        if (a.start_level == levels) { test_leaf(root); continue; }
        {
            N rb = load_struct(bvh.nodes + (root - s_skip[a.start_level] - 1u));
            if (!isintersection(rb, p, d)) continue;
        }
        uint32_t inode = root;
        int level = a.start_level;
        uint32_t pending = 0;             // bit k: the right child of my ancestor at level k is still to be visited
        while (true) {
            const uint32_t c0 = 2u * inode, c1 = c0 + 1u;
            const int cl = level + 1;
            const bool c1_real = (c1 - (1u << level)) < s_nreal[cl];
            bool descended = false;
            if (cl == levels) {
                test_leaf(c0);
                if (c1_real) test_leaf(c1);
            } else {
                const N* cp = bvh.nodes + (c0 - s_skip[cl] - 1u);
                const N b0 = load_struct(cp);
                const bool h0 = isintersection(b0, p, d);
                bool h1 = false;
                if (c1_real) { const N b1 = load_struct(cp + 1); h1 = isintersection(b1, p, d); }
                if (h0) { if (h1) pending |= 1u << level; inode = c0; level = cl; descended = true; }
                else if (h1) { inode = c1; level = cl; descended = true; }
            }
            if (descended) continue;
            if (pending == 0) break;
            const int k = 31 - __clz(pending);            // deepest level with a pending right child
            pending &= ~(1u << k);
            inode = 2u * (inode >> (level - k)) + 1u;
            level = k + 1;
            // the popped node's box was already tested (h1) when it was marked pending: expand it directly,
            // unless it is a leaf
            if (level == levels) {                        // cannot happen (leaves are never marked pending), kept for safety
                test_leaf(inode);
                if (pending == 0) break;
            }
        }
    }
    if constexpr (MODE == kCount) counts[qi] = (I)em.pos;
}

// ---- long rays --------------------------------------------------------------------------------------------------
// A ray grazing the 1 M-sphere shell of configs[3] hits up to ~2300 leaves, and one lane walks ~10 dependent L2-latency
// node steps per hit (a ray with 94 hits took 0.65 ms on its own). Such rays set a fixed ~2 ms tail per call — which is
// what kept the 8-GPU ray scaling at 6.8x (tools/rays_scaling_probe.py: t(R) = 0.091 ms / M rays + 2.1 ms).
// A lane of rays_persistent_kernel whose ray is still running after
// `after` node steps EXPORTS it — ray number, current node, pending right children, hits so far: 32 bytes — to a queue
// and takes the next ray; rays_wide_kernel then finishes each queued ray with a whole warp: the exported nodes (all
// already box-tested) seed a shared-memory stack, and every step the 32 lanes pop the 32 newest nodes, test their
// children and push the hit ones back (ballot-compacted) until the stack is empty. Popping the newest nodes keeps the
// walk depth-first, 32 wide: the stack grows by <= 32 per level; above kRaysWideFull entries one node is popped per
// step (plain depth-first, +1 per level), so it cannot overflow. A full queue just leaves the ray with its lane.
// (Finishing the ray inside the persistent kernel, inlined or as a called function, cost the per-lane loop its register
// budget: 47 -> 56 registers + spills, 100 M rays 93 -> 133 ms, with a 48-register cap 216 ms. Hence the second kernel.)
struct alignas(16) RayWideEntry { uint32_t q_lo, q_hi, inode, pending, level, root, count, count_hi; };   // count: hits so far (kCount) / next output slot (kWrite)
struct RayWideQueue {
    RayWideEntry* data;              // nullptr = no export
    unsigned long long* count;       // entries requested (may exceed cap: only the first cap were stored)
    uint32_t cap;
    uint32_t after;                  // node steps before a ray is exported
};
constexpr int kRaysWideFull = 768;                         // above this many stacked nodes: pop one per step
constexpr int kRaysWideCap = kRaysWideFull + 64 + 64;      // + one full step's pushes + a depth-first descent

// Hit buffer of one warp (unordered mode): `hit` holds *nhit_slot buffered (leaf, ray) pairs. Flushes whole 32-entry
// rounds (all = everything) with one atomic on `total` + one coalesced store per round; the fused multi-GPU variant
// (HB > 128) flushes once >= HB - 128 hits are buffered, two hits per 16-byte multimem.st. Warp-uniform.
template <int MODE, class I, int HB>
IBVH_D void rays_flush_hits(IndexPair<I>* hit, unsigned int* nhit_slot, int lane, unsigned long long* total, int64_t capacity,
                            IndexPair<I>* contacts, bool all) {
    constexpr bool kFused = HB > 128;
    if constexpr (MODE == kAtomic) {
        __syncwarp();
        unsigned int n = *nhit_slot;
        if constexpr (kFused) {
            if (n >= (unsigned)(HB - 128) || (all && n > 0u)) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(total, (unsigned long long)n);      // local counter: slots index this rank's region
                base = __shfl_sync(0xffffffffu, base, 0);
                if ((int64_t)(base + n) <= capacity) {
                    if constexpr (sizeof(IndexPair<I>) == 8) {
                        const unsigned head = (unsigned)(base & 1ull), npair = (n - head) >> 1;
                        for (unsigned k = lane; k < npair; k += 32) {
                            const IndexPair<I> p0 = hit[head + 2 * k], p1 = hit[head + 2 * k + 1];
                            uint4 v;
                            v.x = (uint32_t)p0.a; v.y = (uint32_t)p0.b; v.z = (uint32_t)p1.a; v.w = (uint32_t)p1.b;
                            multimem_st_v4(contacts + base + head + 2 * k, v);
                        }
                        if (lane == 0 && head) multimem_store_pair(contacts + base, hit[0]);
                        if (lane == 1 && ((n - head) & 1u)) multimem_store_pair(contacts + base + n - 1, hit[n - 1]);
                    } else {
                        for (unsigned k = lane; k < n; k += 32) multimem_store_pair(contacts + base + k, hit[k]);
                    }
                }
                n = 0;
            }
        }
        while (!kFused && (n >= 32u || (all && n > 0u))) {
            const unsigned int take = n >= 32u ? 32u : n;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(total, (unsigned long long)take);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((unsigned)lane < take && (int64_t)(base + lane) < capacity) contacts[base + lane] = hit[n - take + lane];
            n -= take;
        }
        __syncwarp();
        if (lane == 0) *nhit_slot = n;
        __syncwarp();
    }
}

// One warp per exported ray (ticket order). MODE kAtomic: hits appended to the list; kCount: counts[ray] = hits so far
// (as exported) + the hits found here; kWrite: the remaining hits written from the exported slot on, IN THE REFERENCE'S
// ORDER. That order is depth-first, left child first = ascending leaf position, and the stack keeps it: it is sorted
// with the leftmost node on top, the popped nodes' children are pushed back in order (they lie left of everything
// below), and a popped leaf parent emits its hits only if no internal node was popped left of it in the same step —
// otherwise it is pushed back as it is and comes up again after that node's subtree.
template <int MODE, class LT, class N, class I, int HB = 128>
__global__ void __launch_bounds__(128) rays_wide_kernel(const typename LT::value_type* __restrict__ points,
                                                       const typename LT::value_type* __restrict__ dirs,
                                                       DBvh<LT, N> bvh, TraverseArgs a, I* counts, IndexPair<I>* contacts,
                                                       RayWideQueue wq, unsigned long long* ticket) {
    constexpr bool kOrdered = MODE == kWrite;
    using T = typename LT::value_type;
    using V = typename LT::vol_t;
    __shared__ uint32_t s_skip[34];
    __shared__ uint32_t s_nreal[34];
    __shared__ IndexPair<I> s_hit[4][MODE == kAtomic ? HB : 1];
    __shared__ unsigned int s_nhit[4];
    __shared__ uint32_t s_stack[4][kRaysWideCap];
    for (int i = threadIdx.x; i < 34; i += blockDim.x) { s_skip[i] = (uint32_t)bvh.ti.skips[i]; s_nreal[i] = (uint32_t)bvh.ti.level_nreal[i]; }
    if (threadIdx.x < 4) s_nhit[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int levels = bvh.ti.levels;
    const uint32_t leaf0 = 1u << (levels - 1);
    const uint32_t inode_end = (1u << (a.start_level - 1)) + s_nreal[a.start_level] - 1u;
    unsigned long long nq = *wq.count;
    if (nq > wq.cap) nq = wq.cap;
    uint32_t* stack = s_stack[w];
    const unsigned lt = (1u << lane) - 1u;
    while (true) {
        unsigned long long e = 0;
        if (lane == 0) e = atomicAdd(ticket, 1ull);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= nq) break;
        const RayWideEntry en = wq.data[e];
        const int64_t qi = (int64_t)(((unsigned long long)en.q_hi << 32) | en.q_lo);
        const int64_t q = a.q_begin + qi;
        T p[3], d[3], inv[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { p[k] = points[3 * q + k]; d[k] = dirs[3 * q + k]; inv[k] = T(1) / d[k]; }
        const I ray_id = (I)(a.id_base + q + 1);
        uint32_t nhit = 0;                                                 // kCount: hits found by this lane
        int64_t pos = (int64_t)(((unsigned long long)en.count_hi << 32) | en.count);      // kWrite: next slot (warp-uniform)
        uint32_t n = 0;                                                    // stacked nodes (warp-uniform)
        auto leaf_test = [&](uint32_t node) -> bool {
            V v;
            const uint2* sp = reinterpret_cast<const uint2*>(bvh.leaves + (node - leaf0));
            uint2* dp = reinterpret_cast<uint2*>(&v);
#pragma unroll
            for (int k = 0; k < (int)(sizeof(V) / 8); ++k) dp[k] = __ldg(sp + k);
            return isintersection(v, p, d);
        };
        auto reported = [&](uint32_t node) -> I { return a.positions ? (I)(node - leaf0 + 1u) : (I)bvh.leaves[node - leaf0].index; };
        auto run = [&]() {
            __syncwarp();
            while (n > 0u) {
                const uint32_t take = n > (uint32_t)kRaysWideFull ? 1u : (n < 32u ? n : 32u);
                const bool have = (uint32_t)lane < take;
                uint32_t node = 1u;
                if (have) node = stack[n - 1u - (uint32_t)lane];         // lane 0 = top = leftmost
                n -= take;
                __syncwarp();
                const int lv = 32 - __clz(node);
                const int cl = lv + 1;
                const uint32_t c0 = 2u * node;
                const bool c1_real = (c0 + 1u - (1u << lv)) < s_nreal[cl];
                const bool leafpar = have && cl == levels;
                const bool inner = have && cl != levels;
                bool do_leaves = leafpar, repush = false;
                if constexpr (kOrdered) {
                    const unsigned im = __ballot_sync(0xffffffffu, inner);
                    const int first_inner = im ? __ffs(im) - 1 : 32;
                    do_leaves = leafpar && lane < first_inner;
                    repush = leafpar && lane > first_inner;
                }
                bool h0 = false, h1 = false;
                if (do_leaves) {
                    h0 = leaf_test(c0);
                    h1 = c1_real && leaf_test(c0 + 1u);
                    if constexpr (MODE == kCount) nhit += (uint32_t)h0 + (uint32_t)h1;
                    if constexpr (MODE == kAtomic) {
                        if (h0) s_hit[w][atomicAdd(&s_nhit[w], 1u)] = IndexPair<I>{reported(c0), ray_id};
                        if (h1) s_hit[w][atomicAdd(&s_nhit[w], 1u)] = IndexPair<I>{reported(c0 + 1u), ray_id};
                    }
                } else if (inner) {
                    const N* cp = bvh.nodes + (c0 - s_skip[cl] - 1u);
                    const N b0 = load_struct(cp);
                    h0 = ray_hits_node(b0, p, d, inv);
                    if (c1_real) { const N b1 = load_struct(cp + 1); h1 = ray_hits_node(b1, p, d, inv); }
                }
                if constexpr (kOrdered) {
                    const unsigned e0 = __ballot_sync(0xffffffffu, do_leaves && h0), e1 = __ballot_sync(0xffffffffu, do_leaves && h1);
                    if (do_leaves) {
                        const int64_t at = pos + __popc(e0 & lt) + __popc(e1 & lt);
                        if (h0) contacts[at] = IndexPair<I>{reported(c0), ray_id};
                        if (h1) contacts[at + (h0 ? 1 : 0)] = IndexPair<I>{reported(c0 + 1u), ray_id};
                    }
                    pos += __popc(e0) + __popc(e1);
                }
                // push back, leftmost on top: lane order = left to right, a lane's left child above its right child
                const unsigned m0 = __ballot_sync(0xffffffffu, inner && h0), m1 = __ballot_sync(0xffffffffu, inner && h1);
                const unsigned mr = __ballot_sync(0xffffffffu, repush);
                const uint32_t total = (uint32_t)(__popc(m0) + __popc(m1) + __popc(mr));
                uint32_t idx = n + total - 1u - (uint32_t)(__popc(m0 & lt) + __popc(m1 & lt) + __popc(mr & lt));
                if (repush) stack[idx] = node;
                else if (inner) {
                    if (h0) { stack[idx] = c0; idx -= 1u; }
                    if (h1) stack[idx] = c0 + 1u;
                }
                n += total;
                __syncwarp();
                rays_flush_hits<MODE, I, HB>(s_hit[w], &s_nhit[w], lane, a.total, a.capacity, contacts, false);
            }
        };
        // seeds, all already box-tested, left to right: the node the lane was about to expand, then the pending right
        // children of its ancestors from the deepest ancestor up (bit k of pending = the ancestor at level k)
        const uint32_t npend = (uint32_t)__popc(en.pending);
        n = 1u + npend;
        if (lane == 0) stack[npend] = en.inode;
        if ((en.pending >> lane) & 1u) stack[__popc(en.pending & lt)] = 2u * (en.inode >> ((int)en.level - lane)) + 1u;
        run();
        // the roots right of the exported one (start_level > 1): 32 box tests at a time, leftmost on top
        for (uint32_t r0 = en.root + 1u; r0 <= inode_end && r0 != 0u; r0 += 32u) {
            const uint32_t r = r0 + (uint32_t)lane;
            bool h = false;
            if (r <= inode_end && r >= r0) h = ray_hits_node(load_struct(bvh.nodes + (r - s_skip[a.start_level] - 1u)), p, d, inv);
            const unsigned m = __ballot_sync(0xffffffffu, h);
            n = (uint32_t)__popc(m);
            if (h) stack[n - 1u - (uint32_t)__popc(m & lt)] = r;
            run();
        }
        if constexpr (MODE == kCount) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) nhit += __shfl_xor_sync(0xffffffffu, nhit, off);
            if (lane == 0) counts[qi] = (I)(en.count + nhit);
        }
    }
    rays_flush_hits<MODE, I, HB>(s_hit[w], &s_nhit[w], lane, a.total, a.capacity, contacts, true);
}

// Persistent variant of rays_kernel. Random rays are incoherent: with one ray per thread a warp runs until
// its longest ray is done with a handful of lanes active (measured: 4.2 of 32). Here every warp keeps
// pulling rays from a global ticket counter and a lane that finishes its ray is refilled, so the lanes
// stay busy. Which lane handles a ray is irrelevant for the results: counts / offsets are per ray id.
// HB > 128 = the fused multi-GPU variant (ibvh_traverse_params_t.peer, unordered): `contacts` is the multicast alias
// of this rank's region of every rank's hit list, `a.total` a local slot counter; a warp reserves slots for
// >= HB - 128 hits at a time and writes two hits per 16-byte multimem.st (multicast stores do not coalesce).
template <int MODE, class LT, class N, class I, int HB = 128>
__global__ void __launch_bounds__(128) rays_persistent_kernel(const typename LT::value_type* __restrict__ points,
                                                             const typename LT::value_type* __restrict__ dirs,
                                                             DBvh<LT, N> bvh, TraverseArgs a, I* counts, IndexPair<I>* contacts,
                                                             unsigned long long* ticket, RayWideQueue wq) {
    constexpr bool kFused = HB > 128;
    constexpr bool kWide = true;
    using T = typename LT::value_type;
    using V = typename LT::vol_t;
    __shared__ uint32_t s_skip[34];
    __shared__ uint32_t s_nreal[34];
    // unordered mode: hits are pushed into a per-warp buffer (shared-memory atomic slot counter) and flushed
    // 32 at a time at the warp-uniform top of the loop: one global atomic + one coalesced store per 32 hits
    __shared__ IndexPair<I> s_hit[4][MODE == kAtomic ? HB : 1];
    __shared__ unsigned int s_nhit[4];
    for (int i = threadIdx.x; i < 34; i += blockDim.x) { s_skip[i] = (uint32_t)bvh.ti.skips[i]; s_nreal[i] = (uint32_t)bvh.ti.level_nreal[i]; }
    if (threadIdx.x < 4) s_nhit[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int levels = bvh.ti.levels;
    const uint32_t leaf0 = 1u << (levels - 1);
    const uint32_t inode_start = 1u << (a.start_level - 1);
    const uint32_t inode_end = inode_start + s_nreal[a.start_level] - 1u;

    int64_t qi = -1;                      // ray handled by this lane (index within the shard), -1 = idle
    T p[3] = {0, 0, 0}, d[3] = {0, 0, 0}, inv[3] = {0, 0, 0};      // inv = 1 / d, once per ray (isintersection.jl:6-8 computes it per test)
    I ray_id = 0;
    uint32_t root = 0, inode = 0, pending = 0;
    int level = 0;
    bool need_root = true;
    Emitter<I, MODE> em;
    em.contacts = contacts;
    em.capacity = a.capacity;
    em.total = a.total;
    em.pos = 0;
    bool exhausted = false;

    auto test_leaf = [&](uint32_t node) {
        const LT* lp = bvh.leaves + (node - leaf0);
        V v;
        const uint2* sp = reinterpret_cast<const uint2*>(lp);
        uint2* dp = reinterpret_cast<uint2*>(&v);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / 8); ++k) dp[k] = __ldg(sp + k);
        if (isintersection(v, p, d)) {
            const I li = a.positions ? (I)(node - leaf0 + 1u) : (I)lp->index;
            if constexpr (MODE == kAtomic) s_hit[w][atomicAdd(&s_nhit[w], 1u)] = IndexPair<I>{li, ray_id};
            else em.emit(li, ray_id);
        }
    };
    auto flush_hits = [&](bool all) { rays_flush_hits<MODE, I, HB>(s_hit[w], &s_nhit[w], lane, a.total, a.capacity, contacts, all); };
    uint32_t steps = 0;                   // node steps of this lane's ray
    const uint32_t export_after = (kWide && wq.data) ? wq.after : 0xffffffffu;

    while (true) {
        flush_hits(false);
        // ---- refill idle lanes (warp-uniform decision) ------------------------------------------------
        const unsigned idle = __ballot_sync(0xffffffffu, qi < 0);
        if (idle && !exhausted && (__popc(idle) >= 8 || idle == 0xffffffffu)) {
            const int nidle = __popc(idle);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(ticket, (unsigned long long)nidle);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((int64_t)base + nidle >= a.q_count) exhausted = true;
            if (qi < 0) {
                const int64_t r = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
                if (r < a.q_count) {
                    qi = r;
                    const int64_t q = a.q_begin + r;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { p[k] = points[3 * q + k]; d[k] = dirs[3 * q + k]; inv[k] = T(1) / d[k]; }
                    ray_id = (I)(a.id_base + q + 1);
                    root = inode_start;
                    need_root = true;
                    steps = 0;
                    em.pos = 0;
                    if constexpr (MODE == kWrite) em.pos = (r == 0) ? 0 : (int64_t)counts[r - 1];
                }
            }
        }
        if (__ballot_sync(0xffffffffu, qi >= 0) == 0) break;
        if (qi < 0) continue;                  // idle lanes rejoin at the top (all warp-level calls sit there)
        // ---- one step of this lane's ray ------------------------------------------------------------------
        if (need_root) {
            if (root > inode_end) {                                   // ray finished
                if constexpr (MODE == kCount) counts[qi] = (I)em.pos;
                qi = -1;
            } else if (a.start_level == levels) {
                test_leaf(root);
                root += 1;
            } else {
                N rb = load_struct(bvh.nodes + (root - s_skip[a.start_level] - 1u));
                if (ray_hits_node(rb, p, d, inv)) { inode = root; level = a.start_level; pending = 0; need_root = false; }
                else root += 1;
            }
            continue;
        }
        if constexpr (kWide) {
            if (++steps > export_after) {                              // a long ray: hand it to rays_wide_kernel
                const unsigned long long slot = atomicAdd(wq.count, 1ull);
                if (slot < wq.cap) {
                    uint4* dst = reinterpret_cast<uint4*>(wq.data + slot);
                    dst[0] = make_uint4((uint32_t)qi, (uint32_t)((unsigned long long)qi >> 32), inode, pending);
                    dst[1] = make_uint4((uint32_t)level, root, (uint32_t)em.pos, (uint32_t)((unsigned long long)em.pos >> 32));
                    qi = -1;
                    continue;
                }
                steps = 0;                                               // queue full: the ray stays with this lane
            }
        }
        const uint32_t c0 = 2u * inode, c1 = c0 + 1u;
        const int cl = level + 1;
        const bool c1_real = (c1 - (1u << level)) < s_nreal[cl];
        bool descended = false;
        if (cl == levels) {
            test_leaf(c0);
            if (c1_real) test_leaf(c1);
        } else {
            const N* cp = bvh.nodes + (c0 - s_skip[cl] - 1u);
            const N b0 = load_struct(cp);
            const bool h0 = ray_hits_node(b0, p, d, inv);
            bool h1 = false;
            if (c1_real) { const N b1 = load_struct(cp + 1); h1 = ray_hits_node(b1, p, d, inv); }
            if (h0) { if (h1) pending |= 1u << level; inode = c0; level = cl; descended = true; }
            else if (h1) { inode = c1; level = cl; descended = true; }
        }
        if (!descended) {
            if (pending == 0) { root += 1; need_root = true; }
            else {
                const int k = 31 - __clz(pending);
                pending &= ~(1u << k);
                inode = 2u * (inode >> (level - k)) + 1u;
                level = k + 1;
            }
        }
    }
    flush_hits(true);
}

// ===========================================================================================================
// Packet schedule: one warp per 32 consecutive query leaves (single / pair)
// ===========================================================================================================
constexpr int kPacketWarps = 8;

template <int KIND, int MODE, class LQ, class LT, class N, class I, bool STATS = false>
__global__ void __launch_bounds__(kPacketWarps * 32) lvt_packet_kernel(const LQ* __restrict__ qleaves, DBvh<LT, N> bvh,
                                                                      TraverseArgs a, I* counts, IndexPair<I>* contacts) {
    __shared__ uint2 s_stack[kPacketWarps][34];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * kPacketWarps + w;
    const int64_t qi = warp_global * 32 + lane;
    if (warp_global * 32 >= a.q_count) return;          // whole warp out of range (uniform)
    const bool valid = qi < a.q_count;
    const int64_t q = a.q_begin + qi;
    const TreeInfo& ti = bvh.ti;
    const int levels = ti.levels;

    QueryLeaf<LQ, N> ql;
    if (valid) {
        LQ leaf = load_struct(qleaves + q);
        ql.vol = leaf.volume;
        ql.index = a.positions ? (decltype(leaf.index))(q + 1) : leaf.index;
        ql.bvn = NodeOps<N>::convert(leaf.volume);
    }
    int64_t pos = 0;                                       // kCount: count, kWrite: next slot
    if constexpr (MODE == kWrite) pos = (valid && qi > 0) ? (int64_t)counts[qi - 1] : 0;

    const uint32_t inode_start = 1u << (a.start_level - 1);
    const uint32_t inode_end = inode_start + (uint32_t)ti.level_nreal[a.start_level] - 1u;
    const uint64_t q_impl = (uint64_t)q + (uint64_t(1) << (levels - 1));
    const unsigned valid_mask = __ballot_sync(0xffffffffu, valid);
    uint2* stack = s_stack[w];
    unsigned long long st_node = 0, st_leaf = 0, st_steps = 0, st_loads = 0;

    for (uint32_t root = inode_start; root <= inode_end; ++root) {
        int sp = 0;
        uint32_t inode = root;
        unsigned mask = valid_mask;
        while (true) {
            const int level = 32 - __clz(inode);
            bool act = (mask >> lane) & 1u;
            if constexpr (STATS) st_steps += 1;
            if constexpr (KIND == kSingle) {
                uint64_t rightmost = (((uint64_t)inode + 1u) << (levels - level)) - 1u;
                act = act && (rightmost > q_impl);
            }
            bool descend = false;
            if (level == levels) {
                if (__any_sync(0xffffffffu, act)) {
                    if constexpr (STATS) { st_leaf += act ? 1 : 0; st_loads += 1; }
                    LT leaf = load_struct(bvh.leaves + (inode - (1u << (levels - 1))));
                    if (a.positions) leaf.index = (decltype(leaf.index))(inode - (1u << (levels - 1)) + 1u);
                    bool hit = act && iscontact(ql.vol, leaf.volume);
                    I ea, eb;
                    if constexpr (KIND == kSingle) {
                        if (ql.index > leaf.index) { ea = (I)leaf.index; eb = (I)ql.index; } else { ea = (I)ql.index; eb = (I)leaf.index; }
                    } else {
                        if (a.flip) { ea = (I)leaf.index; eb = (I)ql.index; } else { ea = (I)ql.index; eb = (I)leaf.index; }
                    }
                    if constexpr (MODE == kCount) { pos += hit ? 1 : 0; }
                    else if constexpr (MODE == kWrite) { if (hit) { contacts[pos] = IndexPair<I>{ea, eb}; pos += 1; } }
                    else {
                        unsigned hm = __ballot_sync(0xffffffffu, hit);
                        if (hm) {
                            unsigned long long base = 0;
                            if (lane == 0) base = atomicAdd(a.total, (unsigned long long)__popc(hm));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            unsigned long long slot = base + __popc(hm & ((1u << lane) - 1u));
                            if (hit && (int64_t)slot < a.capacity) contacts[slot] = IndexPair<I>{ea, eb};
                        }
                    }
                }
            } else {
                unsigned am = __ballot_sync(0xffffffffu, act);
                if (am) {
                    if constexpr (STATS) { st_node += act ? 1 : 0; st_loads += 1; }
                    N node = load_struct(bvh.nodes + ((int64_t)inode - ti.skips[level] - 1));
                    bool hit = act && iscontact(ql.bvn, node);
                    unsigned hm = __ballot_sync(0xffffffffu, hit);
                    if (hm) {
                        uint32_t right = 2u * inode + 1u;
                        if (!tree_isvirtual(ti, right, level + 1)) {
                            if (lane == 0) stack[sp] = make_uint2(right, hm);
                            sp += 1;
                        }
                        inode = 2u * inode;
                        mask = hm;
                        descend = true;
                    }
                }
            }
            if (descend) continue;
            if (sp == 0) break;
            __syncwarp();
            uint2 e = stack[--sp];
            inode = e.x;
            mask = e.y;
            __syncwarp();
        }
    }
    if constexpr (MODE == kCount) { if (valid) counts[qi] = (I)pos; }
    if constexpr (STATS) {
        if (a.stats) {
            atomicAdd(a.stats + 0, st_node);
            atomicAdd(a.stats + 1, st_leaf);
            if (lane == 0) { atomicAdd(a.stats + 2, st_steps); atomicAdd(a.stats + 3, st_loads); }
        }
    }
}

// ===========================================================================================================
// Inclusive scan of per-query counts (AK.accumulate!(+), traverse_single.jl:57): reduce / scan / apply
// ===========================================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class I>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const I* __restrict__ in, int64_t n, long long* __restrict__ block_sums) {
    __shared__ long long ws[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    long long s = 0;
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k * kScanThreads + threadIdx.x;
        if (i < n) s += (long long)in[i];
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int j = 0; j < kScanThreads / 32; ++j) t += ws[j];
        block_sums[blockIdx.x] = t;
    }
}
// single block: exclusive scan of block_sums in place; total -> *total
static __global__ void __launch_bounds__(1024) scan_block_sums_kernel(long long* block_sums, int64_t nblocks, unsigned long long* total) {
    __shared__ long long ws[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += 1024) {
        int64_t i = base + threadIdx.x;
        long long v = i < nblocks ? block_sums[i] : 0;
        long long incl = v;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int off = 1; off < 32; off <<= 1) { long long o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
        if (lane == 31) ws[w] = incl;
        __syncthreads();
        long long wb = 0;
        for (int j = 0; j < w; ++j) wb += ws[j];
        long long c = carry;
        if (i < nblocks) block_sums[i] = c + wb + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wb + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (unsigned long long)carry;
}
template <class I>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(I* __restrict__ data, int64_t n, const long long* __restrict__ block_excl) {
    __shared__ long long ws[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems)
    long long v[kScanItems];
    long long s = 0;
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        v[k] = i < n ? (long long)data[i] : 0;
        s += v[k];
    }
    long long incl = s;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int off = 1; off < 32; off <<= 1) { long long o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    long long wb = 0;
    for (int j = 0; j < w; ++j) wb += ws[j];
    long long run = block_excl[blockIdx.x] + wb + incl - s;
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        run += v[k];
        if (i < n) data[i] = (I)run;
    }
}

}  // namespace ibvh
