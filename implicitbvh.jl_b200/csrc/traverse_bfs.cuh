// traverse_bfs.cuh — BFSTraversal: simultaneous breadth-first descent over the bounding-volume test tree (BVTT).
// Replaces src/traverse/breadth_first/{traverse_single,traverse_pair}*.jl and src/raytrace/breadth_first/*.jl.
//
// The reference's GPU backend runs one kernel per level: a thread takes one BVTT entry (two implicit node indices, or a
// node and a ray id), tests the two volumes, and on contact appends the 1..4 pairs of children to the next level's list
// through a per-block buffer and one global atomic per block; the last level tests leaf volumes and appends contacts the
// same way. Its lists are therefore sets (the append order depends on block scheduling), and so are ours: same entries at
// every level (hence the same `num_checks`) and the same contacts, in unspecified order.
//
// What is different here:
//   * entries are 8 bytes ((u32, u32): implicit indices are < 2^32 because levels <= 32) whatever the index type, and live
//     in two library-owned ping-pong buffers — the caller's cache1 only ever receives contacts;
//   * a thread owns FOUR consecutive entries (two 16-byte loads): consecutive entries are mostly the four child pairs of
//     one parent pair, so their node loads hit the same sectors; the block reserves its output with ONE atomic per 1024
//     entries (shuffle scan + one shared-memory round), and every thread's children land contiguously;
//   * the node kernels depend on the NODE type only and the leaf kernels on (leaf volume, index type) only — leaves are
//     read through a byte stride — so the Morton type of the leaves does not multiply the instantiations;
//   * "is the right child virtual" is one compare against the child level's real-node count (passed by the host) instead
//     of the ilog2 + shifts of implicit_tree.jl:191-199 per entry.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace ibvh {

constexpr int kBfsThreads = 256;
constexpr int kBfsItems = 4;                      // entries per thread
constexpr int kBfsTile = kBfsThreads * kBfsItems;

// One side of a BVTT level: where its volumes are and how its children behave.
struct BfsSide {
    const void* base;          // nodes of this level's tree (whole node array) or leaves
    uint32_t sub;              // memory position (0-based) = implicit - sub: num_skips + 1 for a node level, 2^(levels-1) for leaves
    uint32_t stride;           // bytes between elements (leaf arrays: sizeof(BoundingVolume))
    uint32_t child_first;      // implicit index of the first node of the CHILD level (2^level)
    uint32_t child_nreal;      // real nodes on the child level: child 2i+1 is virtual iff 2i+1 - child_first >= child_nreal
    uint32_t vec;              // widest load every element address allows: 16, 8 or 4 bytes (from the base pointer and the stride)
};
IBVH_D bool bfs_right_real(const BfsSide& s, uint32_t implicit) { return 2u * implicit + 1u - s.child_first < s.child_nreal; }
template <class V> IBVH_D V bfs_load(const BfsSide& s, uint32_t implicit) {
    const char* p = (const char*)s.base + (size_t)(implicit - s.sub) * s.stride;
    alignas(16) V v;
    using T = typename V::value_type;
    // (warp-uniform choice: a BBox{Float32} is three 8-byte loads instead of six 4-byte ones, a BSphere{Float32} node one 16-byte load)
    if (sizeof(V) % 16 == 0 && s.vec >= 16u) {
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / 16); ++k) reinterpret_cast<uint4*>(&v)[k] = __ldg(reinterpret_cast<const uint4*>(p) + k);
    } else if (sizeof(V) % 8 == 0 && s.vec >= 8u) {
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / 8); ++k) reinterpret_cast<uint2*>(&v)[k] = __ldg(reinterpret_cast<const uint2*>(p) + k);
    } else {
        const T* f = reinterpret_cast<const T*>(p);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / sizeof(T)); ++k) reinterpret_cast<T*>(&v)[k] = __ldg(f + k);
    }
    return v;
}

// ---- tests between volumes of possibly different float types (pair traversal started at / reaching a leaf level on one
// side only: traverse_pair_gpu.jl `_traverse_nodes_leaves_*`; Julia promotes operation by operation == convert first) ----
template <class C, class T> IBVH_D BBox<C> bfs_box_as(const BBox<T>& b) {
    BBox<C> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = (C)b.lo[k]; o.up[k] = (C)b.up[k]; }
    return o;
}
template <class C, class T> IBVH_D BSphere<C> bfs_sphere_as(const BSphere<T>& s) {
    BSphere<C> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) o.x[k] = (C)s.x[k];
    o.r = (C)s.r;
    return o;
}
template <class TA, class TB> struct BfsCommon { using type = float; };
template <> struct BfsCommon<double, float> { using type = double; };
template <> struct BfsCommon<float, double> { using type = double; };
template <> struct BfsCommon<double, double> { using type = double; };
template <class TA, class TB> IBVH_D bool bfs_contact(const BBox<TA>& a, const BBox<TB>& b) {
    using C = typename BfsCommon<TA, TB>::type;
    return iscontact(bfs_box_as<C>(a), bfs_box_as<C>(b));
}
template <class TA, class TB> IBVH_D bool bfs_contact(const BSphere<TA>& a, const BSphere<TB>& b) {
    using C = typename BfsCommon<TA, TB>::type;
    return iscontact(bfs_sphere_as<C>(a), bfs_sphere_as<C>(b));
}
template <class TA, class TB> IBVH_D bool bfs_contact(const BSphere<TA>& a, const BBox<TB>& b) {     // iscontact.jl:17-23
    return bfs_contact(to_box(a), b);                                                                // box of the sphere in ITS type
}
template <class TA, class TB> IBVH_D bool bfs_contact(const BBox<TA>& a, const BSphere<TB>& b) { return bfs_contact(b, a); }

// ---- block-level reservation: every thread has `k` outputs; returns where its first one goes --------------------------------
// One shuffle scan per warp, one shared-memory round over the warp totals, ONE global atomic per block call.
IBVH_D unsigned long long bfs_reserve(uint32_t k, unsigned long long* counter, uint32_t* s_warp, unsigned long long* s_base) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = k;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < kBfsThreads / 32 ? s_warp[lane] : 0u;
        uint32_t ti = t;
#pragma unroll
        for (int d = 1; d < kBfsThreads / 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += o;
        }
        if (lane < kBfsThreads / 32) s_warp[lane] = ti - t;                 // exclusive offset of each warp
        if (lane == kBfsThreads / 32 - 1) *s_base = ti ? atomicAdd(counter, (unsigned long long)ti) : 0ull;
    }
    __syncthreads();
    // No third barrier: the callers alternate between two sets of (s_warp, s_base). A thread can only write a set again two
    // calls later, i.e. after it has passed the second barrier of the call in between — which every thread reaches only
    // after its reads of this call.
    return *s_base + s_warp[w] + (incl - k);
}

// Loads the (up to) four consecutive entries of a thread. Buffers are 256-byte aligned and a thread's first entry index is
// a multiple of 4, so the two 16-byte loads are aligned; entries past `count` read as zero (implicit index 0 = "none").
IBVH_D void bfs_load_entries(const uint2* src, unsigned long long first, unsigned long long count, uint2 e[kBfsItems]) {
    if (first + kBfsItems <= count) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src + first));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(src + first) + 1);
        e[0] = make_uint2(a.x, a.y); e[1] = make_uint2(a.z, a.w); e[2] = make_uint2(b.x, b.y); e[3] = make_uint2(b.z, b.w);
    } else {
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) e[j] = first + j < count ? __ldg(src + first + j) : make_uint2(0u, 0u);
    }
}

// ---- node levels ------------------------------------------------------------------------------------------------------------
// MODE: which sides sprout, as in traverse_pair.jl:39-140.
enum { kBfsSingle = 0, kBfsBoth = 1, kBfsLeft = 2, kBfsRight = 3 };

// VA / VB: volume types read on the two sides (node types; a leaf volume type on the side that already is at its leaves).
// (float nodes: 34 registers left 6 of the 8 CTAs launched per SM resident — profiles/r2_ncu_bfs.csv; capped at 32 for 8)
template <class VA, class VB> constexpr int bfs_nodes_minb() { return (sizeof(typename VA::value_type) == 4 && sizeof(typename VB::value_type) == 4) ? 8 : 1; }
template <int MODE, class VA, class VB>
__global__ void __launch_bounds__(kBfsThreads, bfs_nodes_minb<VA, VB>()) bfs_nodes_kernel(const uint2* __restrict__ src, unsigned long long count, BfsSide sa, BfsSide sb,
                                                                int self_checks, uint2* __restrict__ dst, unsigned long long* counter) {
    __shared__ uint32_t s_warp[2][kBfsThreads / 32];
    __shared__ unsigned long long s_base[2];
    int par = 0;                                 // which set this tile's reservation uses (see bfs_reserve)
    const unsigned long long tiles = (count + kBfsTile - 1) / kBfsTile;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const unsigned long long first = tile * kBfsTile + (unsigned long long)threadIdx.x * kBfsItems;
        uint2 e[kBfsItems];
        bfs_load_entries(src, first, count, e);
        // 4 bits per entry: which of (left,left) (left,right) (right,left) (right,right) [or the one-sided forms] to emit
        uint32_t mask = 0, k = 0;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            if (e[j].x == 0u) continue;
            uint32_t m = 0;
            if (MODE == kBfsSingle && e[j].x == e[j].y) {
                // self-check (traverse_single_gpu.jl:63-80): (l,l) (l,r) (r,r); below the second-to-last level only (l,r)
                const bool rr = bfs_right_real(sa, e[j].x);
                if (self_checks) m = rr ? 0xBu : 0x1u;          // bits: 0 = (2a,2b), 1 = (2a,2b+1), 2 = (2a+1,2b), 3 = (2a+1,2b+1)
                else m = rr ? 0x2u : 0x0u;
            } else {
                const VA va = bfs_load<VA>(sa, e[j].x);
                const VB vb = bfs_load<VB>(sb, e[j].y);
                if (bfs_contact(va, vb)) {
                    if (MODE == kBfsSingle) m = bfs_right_real(sb, e[j].y) ? 0xFu : 0x5u;       // node 1 is left of node 2: its children are real
                    else if (MODE == kBfsBoth) {
                        const bool r1 = bfs_right_real(sa, e[j].x), r2 = bfs_right_real(sb, e[j].y);
                        m = 0x1u | (r2 ? 0x2u : 0u) | (r1 ? 0x4u : 0u) | (r1 && r2 ? 0x8u : 0u);
                    } else if (MODE == kBfsLeft) m = 0x1u | (bfs_right_real(sa, e[j].x) ? 0x4u : 0u);   // (2a, b), (2a+1, b)
                    else m = 0x1u | (bfs_right_real(sb, e[j].y) ? 0x2u : 0u);                           // (a, 2b), (a, 2b+1)
                }
            }
            mask |= m << (4 * j);
            k += __popc(m);
        }
        unsigned long long pos = bfs_reserve(k, counter, s_warp[par], &s_base[par]);
        par ^= 1;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            const uint32_t m = (mask >> (4 * j)) & 0xFu;
            if (!m) continue;
            const uint32_t a2 = (MODE == kBfsRight) ? e[j].x : 2u * e[j].x, b2 = (MODE == kBfsLeft) ? e[j].y : 2u * e[j].y;
            if (m & 1u) dst[pos++] = make_uint2(a2, b2);
            if (m & 2u) dst[pos++] = make_uint2(a2, b2 + 1u);
            if (m & 4u) dst[pos++] = make_uint2(a2 + 1u, b2);
            if (m & 8u) dst[pos++] = make_uint2(a2 + 1u, b2 + 1u);
        }
    }
}

// ---- leaf level: contacts (traverse_single_gpu.jl:143-211, traverse_pair_gpu.jl `_traverse_leaves_pair_gpu!`) --------------------
// SORT_PAIR: single tree — the indices of a contact are emitted in ascending order; pair — (index in bvh1, index in bvh2).
// positions != 0: leaf positions (1-based) instead of the leaves' indices (how the caller applies `narrow`).
template <bool SORT_PAIR, class V, class I>
__global__ void __launch_bounds__(kBfsThreads) bfs_leaves_kernel(const uint2* __restrict__ src, unsigned long long count, BfsSide sa, BfsSide sb,
                                                                 uint32_t index_offset, int positions, IndexPair<I>* __restrict__ out,
                                                                 unsigned long long capacity, unsigned long long* counter) {
    __shared__ uint32_t s_warp[2][kBfsThreads / 32];
    __shared__ unsigned long long s_base[2];
    int par = 0;                                 // which set this tile's reservation uses (see bfs_reserve)
    const unsigned long long tiles = (count + kBfsTile - 1) / kBfsTile;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const unsigned long long first = tile * kBfsTile + (unsigned long long)threadIdx.x * kBfsItems;
        uint2 e[kBfsItems];
        bfs_load_entries(src, first, count, e);
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            if (e[j].x == 0u) continue;
            const V va = bfs_load<V>(sa, e[j].x);
            const V vb = bfs_load<V>(sb, e[j].y);
            if (iscontact(va, vb)) mask |= 1u << j;
        }
        unsigned long long pos = bfs_reserve(__popc(mask), counter, s_warp[par], &s_base[par]);
        par ^= 1;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            if (!((mask >> j) & 1u)) continue;
            I ia, ib;
            if (positions) { ia = (I)(e[j].x - sa.sub + 1u); ib = (I)(e[j].y - sb.sub + 1u); }
            else {
                ia = *reinterpret_cast<const I*>((const char*)sa.base + (size_t)(e[j].x - sa.sub) * sa.stride + index_offset);
                ib = *reinterpret_cast<const I*>((const char*)sb.base + (size_t)(e[j].y - sb.sub) * sb.stride + index_offset);
                if (SORT_PAIR && ia > ib) { const I t = ia; ia = ib; ib = t; }
            }
            if (pos < capacity) out[pos] = IndexPair<I>{ia, ib};
            ++pos;
        }
    }
}

// ---- last node level fused with the leaf level ---------------------------------------------------------------------------------
// The children of a contacting pair of leaf parents are leaf pairs. The reference writes them to the next list — the widest of
// the whole descent, ~40 % of all entries — and a second kernel reads them back; here the thread that sprouts them tests the
// leaf volumes at once and appends the contacts. The children are still counted (*children): they are part of num_checks.
template <int MODE, bool SORT_PAIR, class N, class V, class I>
__global__ void __launch_bounds__(kBfsThreads) bfs_last_kernel(const uint2* __restrict__ src, unsigned long long count, BfsSide na, BfsSide nb,
                                                               BfsSide la, BfsSide lb, uint32_t index_offset, int positions,
                                                               IndexPair<I>* __restrict__ out, unsigned long long capacity,
                                                               unsigned long long* counter, unsigned long long* children) {
    __shared__ uint32_t s_warp[2][kBfsThreads / 32];
    __shared__ unsigned long long s_base[2];
    int par = 0;                                 // which set this tile's reservation uses (see bfs_reserve)
    unsigned long long nchild = 0;
    const unsigned long long tiles = (count + kBfsTile - 1) / kBfsTile;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const unsigned long long first = tile * kBfsTile + (unsigned long long)threadIdx.x * kBfsItems;
        uint2 e[kBfsItems];
        bfs_load_entries(src, first, count, e);
        uint32_t hitmask = 0;                       // 4 bits per entry, bit layout of bfs_nodes_kernel
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            if (e[j].x == 0u) continue;
            uint32_t m = 0;
            if (MODE == kBfsSingle && e[j].x == e[j].y) m = bfs_right_real(na, e[j].x) ? 0x2u : 0x0u;     // leaf self-checks are pointless: only (l, r)
            else if (bfs_contact(bfs_load<N>(na, e[j].x), bfs_load<N>(nb, e[j].y))) {
                if (MODE == kBfsSingle) m = bfs_right_real(nb, e[j].y) ? 0xFu : 0x5u;
                else {
                    const bool r1 = bfs_right_real(na, e[j].x), r2 = bfs_right_real(nb, e[j].y);
                    m = 0x1u | (r2 ? 0x2u : 0u) | (r1 ? 0x4u : 0u) | (r1 && r2 ? 0x8u : 0u);
                }
            }
            if (!m) continue;
            nchild += __popc(m);
            // the (up to) four leaves involved, loaded once
            const uint32_t a2 = 2u * e[j].x, b2 = 2u * e[j].y;
            const V va0 = bfs_load<V>(la, a2), vb0 = bfs_load<V>(lb, (m & 0x5u) ? b2 : b2 + 1u);
            const V va1 = (m & 0xCu) ? bfs_load<V>(la, a2 + 1u) : va0;
            const V vb1 = ((m & 0xAu) && (m & 0x5u)) ? bfs_load<V>(lb, b2 + 1u) : vb0;
            uint32_t hm = 0;
            if ((m & 0x1u) && iscontact(va0, vb0)) hm |= 0x1u;
            if ((m & 0x2u) && iscontact(va0, vb1)) hm |= 0x2u;
            if ((m & 0x4u) && iscontact(va1, vb0)) hm |= 0x4u;
            if ((m & 0x8u) && iscontact(va1, vb1)) hm |= 0x8u;
            hitmask |= hm << (4 * j);
        }
        unsigned long long pos = bfs_reserve(__popc(hitmask), counter, s_warp[par], &s_base[par]);
        par ^= 1;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            uint32_t hm = (hitmask >> (4 * j)) & 0xFu;
            while (hm) {
                const int b = __ffs(hm) - 1;
                hm &= hm - 1;
                const uint32_t ca = 2u * e[j].x + (uint32_t)(b >> 1), cb = 2u * e[j].y + (uint32_t)(b & 1);
                I ia, ib;
                if (positions) { ia = (I)(ca - la.sub + 1u); ib = (I)(cb - lb.sub + 1u); }
                else {
                    ia = *reinterpret_cast<const I*>((const char*)la.base + (size_t)(ca - la.sub) * la.stride + index_offset);
                    ib = *reinterpret_cast<const I*>((const char*)lb.base + (size_t)(cb - lb.sub) * lb.stride + index_offset);
                    if (SORT_PAIR && ia > ib) { const I t = ia; ia = ib; ib = t; }
                }
                if (pos < capacity) out[pos] = IndexPair<I>{ia, ib};
                ++pos;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nchild += __shfl_xor_sync(0xffffffffu, nchild, d);
    if ((threadIdx.x & 31) == 0 && nchild) atomicAdd(children, nchild);
}

// ---- rays (raytrace/breadth_first/raytrace_gpu.jl): entries are (implicit node index, 1-based ray id) ------------------------
template <class T> IBVH_D void bfs_load_ray(const T* points, const T* dirs, uint32_t iray, T p[3], T d[3]) {
    const size_t o = (size_t)(iray - 1u) * 3u;
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] = __ldg(points + o + k); d[k] = __ldg(dirs + o + k); }
}
template <class N>
__global__ void __launch_bounds__(kBfsThreads) bfs_rays_nodes_kernel(const uint2* __restrict__ src, unsigned long long count, BfsSide sa,
                                                                     const typename N::value_type* __restrict__ points,
                                                                     const typename N::value_type* __restrict__ dirs,
                                                                     uint2* __restrict__ dst, unsigned long long* counter) {
    using T = typename N::value_type;
    __shared__ uint32_t s_warp[2][kBfsThreads / 32];
    __shared__ unsigned long long s_base[2];
    int par = 0;                                 // which set this tile's reservation uses (see bfs_reserve)
    const unsigned long long tiles = (count + kBfsTile - 1) / kBfsTile;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const unsigned long long first = tile * kBfsTile + (unsigned long long)threadIdx.x * kBfsItems;
        uint2 e[kBfsItems];
        bfs_load_entries(src, first, count, e);
        uint32_t mask = 0, k = 0;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            if (e[j].x == 0u) continue;
            T p[3], d[3];
            bfs_load_ray(points, dirs, e[j].y, p, d);
            if (isintersection(bfs_load<N>(sa, e[j].x), p, d)) {
                const uint32_t m = bfs_right_real(sa, e[j].x) ? 3u : 1u;
                mask |= m << (2 * j);
                k += __popc(m);
            }
        }
        unsigned long long pos = bfs_reserve(k, counter, s_warp[par], &s_base[par]);
        par ^= 1;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            const uint32_t m = (mask >> (2 * j)) & 3u;
            if (m & 1u) dst[pos++] = make_uint2(2u * e[j].x, e[j].y);
            if (m & 2u) dst[pos++] = make_uint2(2u * e[j].x + 1u, e[j].y);
        }
    }
}
template <class V, class I>
__global__ void __launch_bounds__(kBfsThreads) bfs_rays_leaves_kernel(const uint2* __restrict__ src, unsigned long long count, BfsSide sa,
                                                                      const typename V::value_type* __restrict__ points,
                                                                      const typename V::value_type* __restrict__ dirs,
                                                                      uint32_t index_offset, int positions, long long id_base,
                                                                      IndexPair<I>* __restrict__ out, unsigned long long capacity,
                                                                      unsigned long long* counter) {
    using T = typename V::value_type;
    __shared__ uint32_t s_warp[2][kBfsThreads / 32];
    __shared__ unsigned long long s_base[2];
    int par = 0;                                 // which set this tile's reservation uses (see bfs_reserve)
    const unsigned long long tiles = (count + kBfsTile - 1) / kBfsTile;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const unsigned long long first = tile * kBfsTile + (unsigned long long)threadIdx.x * kBfsItems;
        uint2 e[kBfsItems];
        bfs_load_entries(src, first, count, e);
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            if (e[j].x == 0u) continue;
            T p[3], d[3];
            bfs_load_ray(points, dirs, e[j].y, p, d);
            if (isintersection(bfs_load<V>(sa, e[j].x), p, d)) mask |= 1u << j;
        }
        unsigned long long pos = bfs_reserve(__popc(mask), counter, s_warp[par], &s_base[par]);
        par ^= 1;
#pragma unroll
        for (int j = 0; j < kBfsItems; ++j) {
            if (!((mask >> j) & 1u)) continue;
            I ia;
            if (positions) ia = (I)(e[j].x - sa.sub + 1u);
            else ia = *reinterpret_cast<const I*>((const char*)sa.base + (size_t)(e[j].x - sa.sub) * sa.stride + index_offset);
            if (pos < capacity) out[pos] = IndexPair<I>{ia, (I)(id_base + (long long)e[j].y)};
            ++pos;
        }
    }
}

// ---- initial BVTT (traverse_single.jl:69-157, traverse_pair.jl:161-221, raytrace/breadth_first/breadth_first.jl:69-140) --------
// Single tree: all pairs (i, j), i <= j (i < j when starting at the leaf level), of the real nodes of the start level, row-major
// as in the reference's CPU loop; entry positions are computed in closed form (no float sqrt as in the reference's tri_ij).
static __global__ void bfs_init_single_kernel(uint2* dst, uint32_t first, uint32_t nreal, int with_self) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long total = (unsigned long long)nreal * nreal;
    if (t >= total) return;
    const uint32_t i = (uint32_t)(t / nreal), j = (uint32_t)(t % nreal);
    if (j < i || (!with_self && j == i)) return;
    // rows before i hold (nreal - r) entries each with the diagonal, (nreal - 1 - r) without
    const unsigned long long n = nreal, r = i;
    const unsigned long long row0 = with_self ? r * n - r * (r - 1) / 2 : r * (n - 1) - r * (r - 1) / 2;
    dst[row0 + (j - i) - (with_self ? 0 : 1)] = make_uint2(first + i, first + j);
}
// Pair / rays: the full product, row-major: entry k = (first_a + k / nb, first_b + k % nb)
static __global__ void bfs_init_product_kernel(uint2* dst, uint32_t first_a, unsigned long long na, uint32_t first_b, unsigned long long nb) {
    const unsigned long long total = na * nb;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x)
        dst[t] = make_uint2(first_a + (uint32_t)(t / nb), first_b + (uint32_t)(t % nb));
}

// ---- sorted / unique contact lists ------------------------------------------------------------------------
// The reference's tests compare `sort(traversal.contacts)` (test/gputests.jl:73-78, runtests.jl:1246-1250): lexicographic by
// (a, b). On the device that is the onesweep radix sort of radix_sort.cuh over the 64-bit keys (a << 32 | b).
// pairs -> keys (+ the digit histograms of all passes; *bad is set if an index does not fit 32 unsigned bits)
template <class I>
__global__ void __launch_bounds__(256) contact_keys_kernel(const IndexPair<I>* __restrict__ pairs, int64_t n, uint64_t* __restrict__ keys,
                                                          uint32_t* __restrict__ hist, uint32_t* bad) {
    constexpr int P = radix_passes<uint64_t>();
    constexpr int RB = radix_bits<uint64_t>(), BINS = radix_bins<uint64_t>();
    __shared__ uint32_t sh[P][BINS];
    for (int i = threadIdx.x; i < P * BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const IndexPair<I> pr = pairs[i];
        const uint64_t a = (uint64_t)(long long)pr.a, b = (uint64_t)(long long)pr.b;
        if ((a >> 31) != 0 || (b >> 32) != 0) *bad = 1u;          // (a < 2^31: the key keeps bit 63 clear, as the sort's sentinel needs)
        const uint64_t m = (a << 32) | (b & 0xffffffffull);
        keys[i] = m;
#pragma unroll
        for (int p = 0; p < P; ++p) atomicAdd(&sh[p][(uint32_t)(m >> (p * RB)) & (BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * BINS; i += blockDim.x) {
        const uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}
// flags[i] = 1 if sorted key i starts a run of equal keys
static __global__ void __launch_bounds__(256) contact_flags_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// sorted keys -> pairs; slots != nullptr: the inclusive scan of the run-start flags (unique: key i goes to slot slots[i] - 1 if it starts a run)
template <class I>
__global__ void __launch_bounds__(256) contact_unpack_kernel(const uint64_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ slots,
                                                            IndexPair<I>* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t m = keys[i];
        int64_t dst = i;
        if (slots) {
            if (i > 0 && keys[i - 1] == m) continue;
            dst = (int64_t)slots[i] - 1;
        }
        out[dst] = IndexPair<I>{(I)(m >> 32), (I)(m & 0xffffffffull)};
    }
}

}  // namespace ibvh
