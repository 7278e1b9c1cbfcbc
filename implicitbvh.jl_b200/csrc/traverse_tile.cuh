// traverse_tile.cuh — the B200 schedule for LVT contact traversal over BBox nodes (single tree and
// BVH-vs-BVH): "group walk + dense tiles". Same results as traverse_lvt_single! / traverse_lvt_pair!
// (src/traverse/leaf_vs_tree/traverse_single.jl:136-208, traverse_pair.jl:176-244), bit for bit.
//
// Why it is exact. The reference reports (q, j) iff the leaf test passes and every ancestor of j from
// start_level down to the level above the leaves passes iscontact(BBox(q), node). BBox nodes are merged
// with exact min / max (src/bounding_volumes/merge.jl:30-43) and a virtual right child copies the left
// one (src/build.jl:513-515), so every node's box CONTAINS its children's boxes exactly and the
// closed-interval box test is monotone under containment: if BBox(q) touches a node it touches all of
// that node's ancestors. For levels above the leaf parents the predicate therefore collapses to
//        P(q, j)  =  leaf_test(q, j)  AND  iscontact(BBox(q), parent_{levels-1}(j))          (*)
// (the leaf-parent box itself need not contain the leaf's box in floating point — merge.jl:62-68 — so
// that one test is kept). Any candidate search that does not lose a pair satisfying (*) gives the
// reference's contact set. The same containment makes a GROUP query exact-conservative: if U is the
// exact min/max union of the boxes of G consecutive query leaves, then U touches every node that any
// member touches.
//
// Schedule.
//   phase 1  group_walk_kernel: one thread per group of G consecutive query leaves walks the tree with
//            U down to level `levels - log2(G)` only (1/G of the queries, log2(G) fewer levels) and
//            records the target groups B it reaches (count -> scan -> write, ascending B).
//   phase 2  tile_kernel: a warp owns 32/G query groups; lane = (group slot, member). For each recorded
//            B the slot stages B's G leaves and their G/2 leaf-parent boxes in shared memory and every
//            lane evaluates (*) for its own query against the G leaves: dense, divergence-free, coalesced.
// Per query the hits still come out in ascending target position, so the ordered mode reproduces the
// reference order exactly.
#pragma once
#include "common.cuh"
#include "traverse.cuh"

namespace ibvh {

constexpr int kWalkThreads = 128;
constexpr int kTileWarps = 4;

// leaf-level predicate of the reference for two leaf volumes (iscontact.jl:2-14)
template <class V> IBVH_D bool leaf_contact(const V& a, const V& b) { return iscontact(a, b); }
// ---- packed FP32 pairs (sm_100: FADD2 / FMUL2; every half rounds like the scalar instruction) --------------------------------
IBVH_D unsigned long long f32x2_pack(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
IBVH_D float f32x2_lo(unsigned long long v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
IBVH_D float f32x2_hi(unsigned long long v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
IBVH_D unsigned long long f32x2_sub(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
IBVH_D unsigned long long f32x2_add(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
IBVH_D unsigned long long f32x2_mul(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// iscontact(a::BSphere{Float32}, b::BSphere{Float32}) (iscontact.jl:2-4) on the packed FP32 pipe: a record (x, y, z, r) is two
// aligned register pairs as it comes out of a 16-byte load, so
//   (dx, dy) = (b.x, b.y) - (a.x, a.y)     (dz, rs) = (b.z, b.r) - (a.z, -a.r)     2 FADD2   (the -a.r is an operand modifier)
//   (dx^2, dy^2), (dz^2, rs^2)                                                     2 FMUL2
//   (dx^2 + dy^2) + dz^2 <= rs^2                                                   2 FADD + FSETP
// = 7 instructions instead of 11. Bit for bit the reference's expression: (b - a)^2 == (a - b)^2, b.r - (-a.r) == a.r + b.r,
// and the sum keeps the reference's association as SCALAR adds — ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even
// under -fmad=false, so no packed add may follow a packed mul.
#ifndef IBVH_F32X2
#define IBVH_F32X2 1
#endif
#if IBVH_F32X2
template <> IBVH_D bool leaf_contact<BSphere<float>>(const BSphere<float>& a, const BSphere<float>& b) {
    const unsigned long long dxy = f32x2_sub(f32x2_pack(b.x[0], b.x[1]), f32x2_pack(a.x[0], a.x[1]));
    const unsigned long long dzr = f32x2_sub(f32x2_pack(b.x[2], b.r), f32x2_pack(a.x[2], -a.r));
    const unsigned long long mxy = f32x2_mul(dxy, dxy), mzr = f32x2_mul(dzr, dzr);
    return (f32x2_lo(mxy) + f32x2_hi(mxy)) + f32x2_lo(mzr) <= f32x2_hi(mzr);
}
#endif


struct GroupArgs {
    int64_t q_begin;        // first query leaf (0-based) of the shard
    int64_t q_count;        // query leaves in the shard
    int64_t n_groups;       // ceil(q_count / G)
    int32_t start_level;
    int32_t group_level;    // levels - log2(G): target level whose nodes are the target groups
    int32_t flip;
    int32_t positions;      // report 1-based leaf positions instead of .index (IBVH_TRAVERSE_POSITIONS)
    uint32_t seg_cap;       // list entries reserved per query group
    uint32_t step_cap;      // walk-step budget per query group
    int64_t capacity;       // contacts capacity (pairs)
    unsigned long long* total;
    unsigned long long* dbg;     // optional: [0] total walk steps, [1] max steps of one group, [2..] log2 histograms
};

constexpr uint32_t kGroupFlagged = 0xffffffffu;

// ---- phase 1 -------------------------------------------------------------------------------------------
// One pass. glist[A * seg_cap + k] = k-th target group reached by query group A (ascending), gcounts[A] = how
// many. A group whose walk exceeds the step budget or the segment (a group that straddles a coarse Morton
// boundary has a huge union box: a few per mille of the groups, but thousands of steps each) is FLAGGED
// instead: gcounts[A] = kGroupFlagged and A is appended to flist; its members are then traversed one by one
// by the per-query kernel, where each has its own small box. This bounds the tail of both phases.
template <int KIND, int G, class LQ, class LT>
__global__ void __launch_bounds__(kWalkThreads) group_walk_kernel(const LQ* __restrict__ qleaves, int64_t n_query_total,
                                                                 DBvh<LT, BBox<typename LT::value_type>> bvh, GroupArgs a,
                                                                 uint32_t* __restrict__ gcounts, uint32_t* __restrict__ glist,
                                                                 uint32_t* __restrict__ flist, uint32_t* fcount) {
    using T = typename LT::value_type;
    using N = BBox<T>;
    __shared__ int64_t s_skip[34];
    __shared__ int64_t s_nreal[34];
    for (int i = threadIdx.x; i < 34; i += blockDim.x) { s_skip[i] = bvh.ti.skips[i]; s_nreal[i] = bvh.ti.level_nreal[i]; }
    __syncthreads();
    const int64_t A = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (A >= a.n_groups) return;
    const int levels = bvh.ti.levels;
    const int glevel = a.group_level;

    // exact union of the member boxes (the member box is NodeType(leaf.volume), traverse_single.jl:154-155)
    const int64_t q0 = a.q_begin + A * G;
    N U;
    {
        LQ leaf = load_struct(qleaves + q0);
        U = NodeOps<N>::convert(leaf.volume);
#pragma unroll
        for (int k = 1; k < G; ++k) {
            if (q0 + k < a.q_begin + a.q_count && q0 + k < n_query_total) {
                LQ l2 = load_struct(qleaves + q0 + k);
                U = merge(U, NodeOps<N>::convert(l2.volume));
            }
        }
    }
    uint32_t cnt = 0;
    uint32_t* seg = glist + (size_t)A * a.seg_cap;
    const uint32_t inode_start = 1u << (a.start_level - 1);
    const uint32_t inode_end = inode_start + (uint32_t)s_nreal[a.start_level] - 1u;
    // single tree: a subtree is useful only if it holds a leaf strictly right of the group's first leaf
    const uint64_t q_impl = (uint64_t)q0 + (uint64_t(1) << (levels - 1));
    uint32_t stack[32];
    uint32_t steps = 0;
    bool flagged = false;
    for (uint32_t root = inode_start; root <= inode_end && !flagged; ++root) {
        int sp = 0;
        uint32_t inode = root;
        while (true) {
            const int level = 32 - __clz(inode);
            bool descend = false;
            bool skip = false;
            if (++steps > a.step_cap) { flagged = true; break; }
            if constexpr (KIND == kSingle) {
                uint64_t rightmost = (((uint64_t)inode + 1u) << (levels - level)) - 1u;
                skip = rightmost <= q_impl;
            }
            if (!skip) {
                N node = load_struct(bvh.nodes + ((int64_t)inode - s_skip[level] - 1));
                if (iscontact(U, node)) {
                    if (level == glevel) {
                        if (cnt == a.seg_cap) { flagged = true; break; }
                        seg[cnt++] = inode - (1u << (glevel - 1));
                    } else {
                        uint32_t right = 2u * inode + 1u;
                        if ((int64_t)(right - (1u << level)) < s_nreal[level + 1]) stack[sp++] = right;     // right child is real
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
    if (flagged) {
        gcounts[A] = kGroupFlagged;
        flist[atomicAdd(fcount, 1u)] = (uint32_t)A;
    } else {
        gcounts[A] = cnt;
    }
    if (a.dbg) {
        atomicAdd(a.dbg + 0, (unsigned long long)steps);
        atomicAdd(a.dbg + 2 + (31 - __clz((int)(steps | 1u))), 1ull);
        atomicAdd(a.dbg + 40 + (31 - __clz((int)(cnt | 1u))), 1ull);
        if (flagged) atomicAdd(a.dbg + 1, 1ull);
    }
}

// ---- phase 2 -------------------------------------------------------------------------------------------
// MODE kCount -> counts[q]; kWrite -> contacts at counts[q-1] (counts = inclusive scan of per-query
// counts); kAtomic -> unordered append. Members of flagged groups are left to the per-query kernel.
template <int KIND, int MODE, int G, class LQ, class LT, class I>
__global__ void __launch_bounds__(kTileWarps * 32) tile_kernel(const LQ* __restrict__ qleaves, int64_t n_query_total,
                                                              DBvh<LT, BBox<typename LT::value_type>> bvh, GroupArgs a,
                                                              const uint32_t* __restrict__ gcounts, const uint32_t* __restrict__ glist,
                                                              I* counts, IndexPair<I>* contacts) {
    using T = typename LT::value_type;
    using N = BBox<T>;
    using VT = typename LT::vol_t;
    using VQ = typename LQ::vol_t;
    constexpr int SLOTS = 32 / G;
    // staged target group per slot; 16-byte aligned elements (one LDS.128 per sphere) and one element of
    // padding per slot so that the four slots of a warp fall into different banks
    struct alignas(16) TVol { VT v; };
    struct alignas(8) TPar { N b; };
    __shared__ TVol s_vol[kTileWarps][SLOTS][G + 1];
    __shared__ typename LT::idx_t s_idx[kTileWarps][SLOTS][G];
    __shared__ TPar s_par[kTileWarps][SLOTS][(G + 1) / 2 + 1];

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / G, m = lane % G;
    const int64_t warp_global = (int64_t)blockIdx.x * kTileWarps + w;
    const int64_t A = warp_global * SLOTS + slot;                 // query group of this lane
    if (warp_global * SLOTS >= a.n_groups) return;                // warp-uniform
    const bool a_valid = A < a.n_groups;
    const int64_t qi = A * G + m;                                   // query index within the shard
    const int64_t q = a.q_begin + qi;                               // query leaf position
    const bool q_valid = a_valid && qi < a.q_count && q < n_query_total;
    const int levels = bvh.ti.levels;
    const int64_t n_target = bvh.ti.n;
    const int64_t par_base = bvh.ti.level_start[levels - 1];       // memory position of the first leaf-parent node
    const int64_t par_real = bvh.ti.level_nreal[levels - 1];

    VQ qvol{};
    typename LQ::idx_t qidx = 0;
    N qbox{};
    if (q_valid) {
        LQ leaf = load_struct(qleaves + q);
        qvol = leaf.volume;
        qidx = a.positions ? (typename LQ::idx_t)(q + 1) : leaf.index;
        qbox = NodeOps<N>::convert(leaf.volume);
    }
    uint32_t len = 0;
    bool flagged = false;
    if (a_valid) { len = gcounts[A]; if (len == kGroupFlagged) { len = 0; flagged = true; } }
    const uint32_t* seg = glist + (size_t)(a_valid ? A : 0) * a.seg_cap;
    uint32_t maxlen = len;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { uint32_t o = __shfl_xor_sync(0xffffffffu, maxlen, off); maxlen = o > maxlen ? o : maxlen; }

    int64_t pos = 0;                                                // kCount: count; kWrite: next slot
    if constexpr (MODE == kWrite) pos = (q_valid && qi > 0) ? (int64_t)counts[qi - 1] : 0;

    for (uint32_t k = 0; k < maxlen; ++k) {
        const bool have = k < len;
        int64_t B = 0;
        if (have) B = (int64_t)seg[k];
        // stage the target group's leaves and leaf-parent boxes
        const int64_t j0 = B * G;
        if (have) {
            if (j0 + m < n_target) {
                LT tl = load_struct(bvh.leaves + j0 + m);
                s_vol[w][slot][m].v = tl.volume;
                s_idx[w][slot][m] = a.positions ? (decltype(tl.index))(j0 + m + 1) : tl.index;
            }
            if (m < (G + 1) / 2) {
                int64_t pj = j0 / 2 + m;
                if (pj < par_real) s_par[w][slot][m].b = load_struct(bvh.nodes + par_base + pj);
            }
        }
        __syncwarp();
        uint32_t hits = 0;
        if (have && q_valid) {
            // which of the G target leaves are admissible for this query: inside the tree, and (single tree)
            // strictly right of the query leaf (traverse_single.jl:165-167)
            int64_t nval = n_target - j0;
            uint32_t allowed = nval >= G ? ((1u << G) - 1u) : ((1u << (int)nval) - 1u);
            if constexpr (KIND == kSingle) {
                int64_t lo = q - j0 + 1;                                  // first admissible j
                if (lo > 0) allowed &= lo >= G ? 0u : ~((1u << (int)lo) - 1u);
            }
            // leaf predicate for all G leaves (branch-free), the leaf-parent box only for the rare hits
#pragma unroll
            for (int j = 0; j < G; ++j) {
                VT tv = s_vol[w][slot][j].v;
                hits |= (leaf_contact(qvol, tv) ? 1u : 0u) << j;
            }
            hits &= allowed;
            uint32_t cand = hits;
            while (cand) {
                int j = __ffs(cand) - 1;
                cand &= cand - 1;
                if (!iscontact(qbox, s_par[w][slot][j >> 1].b)) hits &= ~(1u << j);
            }
        }
        if constexpr (MODE == kCount) {
            pos += __popc(hits);
        } else if constexpr (MODE == kWrite) {
            while (hits) {
                int j = __ffs(hits) - 1;
                hits &= hits - 1;
                I li = (I)s_idx[w][slot][j];
                I ea, eb;
                if constexpr (KIND == kSingle) { if ((I)qidx > li) { ea = li; eb = (I)qidx; } else { ea = (I)qidx; eb = li; } }
                else { if (a.flip) { ea = li; eb = (I)qidx; } else { ea = (I)qidx; eb = li; } }
                contacts[pos++] = IndexPair<I>{ea, eb};
            }
        } else {
            // unordered: one atomic per warp step, exclusive prefix of the per-lane hit counts
            int nh = __popc(hits);
            int incl = nh;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
            int tot = __shfl_sync(0xffffffffu, incl, 31);
            if (tot) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.total, (unsigned long long)tot);
                base = __shfl_sync(0xffffffffu, base, 0);
                unsigned long long wp = base + (unsigned long long)(incl - nh);
                while (hits) {
                    int j = __ffs(hits) - 1;
                    hits &= hits - 1;
                    I li = (I)s_idx[w][slot][j];
                    I ea, eb;
                    if constexpr (KIND == kSingle) { if ((I)qidx > li) { ea = li; eb = (I)qidx; } else { ea = (I)qidx; eb = li; } }
                    else { if (a.flip) { ea = li; eb = (I)qidx; } else { ea = (I)qidx; eb = li; } }
                    if ((int64_t)wp < a.capacity) contacts[wp] = IndexPair<I>{ea, eb};
                    ++wp;
                }
            }
        }
        __syncwarp();
    }
    if constexpr (MODE == kCount) { if (q_valid && !flagged) counts[qi] = (I)pos; }
}

// ---- phase 2, balanced variant (count / unordered modes) ---------------------------------------------------
// The per-group lists of the 32/G query groups of a warp are walked as ONE flat list, G lanes per entry, so a
// long list does not idle the other slots. The 32 query volumes live in shared memory; everything that is
// only needed for a hit (indices, the query box, the leaf-parent box) is fetched lazily from global memory.
template <int KIND, int MODE, int G, class LQ, class LT, class I>
__global__ void __launch_bounds__(kTileWarps * 32) tile_flat_kernel(const LQ* __restrict__ qleaves, int64_t n_query_total,
                                                                   DBvh<LT, BBox<typename LT::value_type>> bvh, GroupArgs a,
                                                                   const uint32_t* __restrict__ gcounts, const uint32_t* __restrict__ glist,
                                                                   I* counts, IndexPair<I>* contacts) {
    static_assert(MODE == kCount || MODE == kAtomic, "ordered write uses tile_kernel");
    using T = typename LT::value_type;
    using N = BBox<T>;
    using VT = typename LT::vol_t;
    using VQ = typename LQ::vol_t;
    constexpr int SLOTS = 32 / G;
    struct alignas(16) TVol { VT v; };
    struct alignas(16) QVol { VQ v; };
    __shared__ TVol s_vol[kTileWarps][SLOTS][G + 1];
    __shared__ QVol s_q[kTileWarps][32];
    __shared__ uint32_t s_cnt[kTileWarps][32];
    __shared__ uint2 s_buf[MODE == kAtomic ? kTileWarps : 1][MODE == kAtomic ? 32 * G + 32 : 1];

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / G, m = lane % G;
    const int64_t warp_global = (int64_t)blockIdx.x * kTileWarps + w;
    const int64_t A0 = warp_global * SLOTS;                         // first query group of the warp
    // flush up to 32 buffered hits: one atomic for the chunk, indices fetched now, coalesced store
    auto flush = [&](uint2 e, uint32_t n, bool mine) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.total, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (mine) {
            const I qidx = a.positions ? (I)(e.x + 1u) : (I)qleaves[e.x].index;
            const I li = a.positions ? (I)(e.y + 1u) : (I)bvh.leaves[e.y].index;
            I ea, eb;
            if constexpr (KIND == kSingle) { if (qidx > li) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
            else { if (a.flip) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
            const unsigned long long wp = base + (unsigned long long)lane;
            if ((int64_t)wp < a.capacity) contacts[wp] = IndexPair<I>{ea, eb};
        }
    };
    if (A0 >= a.n_groups) return;                                   // warp-uniform
    const uint32_t n_target = (uint32_t)bvh.ti.n;
    const int levels = bvh.ti.levels;

    // my own query (lane -> query A0*G + lane) goes to shared memory
    const int64_t my_qi = A0 * G + lane;
    const bool my_valid = my_qi < a.q_count && a.q_begin + my_qi < n_query_total;
    if (my_valid) {
        LQ leaf = load_struct(qleaves + a.q_begin + my_qi);
        s_q[w][lane].v = leaf.volume;
    }
    s_cnt[w][lane] = 0;
    // list lengths of the SLOTS groups -> exclusive prefix c[0..SLOTS]
    uint32_t len = 0;
    bool my_flagged = false;
    if (lane < SLOTS && A0 + lane < a.n_groups) { len = gcounts[A0 + lane]; if (len == kGroupFlagged) len = 0; }
    {   // every lane learns whether ITS group is flagged (for the final count store)
        int64_t Ag = A0 + lane / G;
        if (Ag < a.n_groups) my_flagged = gcounts[Ag] == kGroupFlagged;
    }
    uint32_t c[SLOTS + 1];
    c[0] = 0;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) c[k + 1] = c[k] + __shfl_sync(0xffffffffu, len, k);
    const uint32_t total_pairs = c[SLOTS];
    const uint32_t* seg0 = glist + (size_t)A0 * a.seg_cap;
    __syncwarp();

    // flat entry p -> (group al of the warp, target group B); the B of the NEXT step is prefetched
    auto entry = [&](uint32_t p, int& al_out) -> uint32_t {
        int al = 0;
#pragma unroll
        for (int k = 1; k < SLOTS; ++k) al += (p >= c[k]) ? 1 : 0;
        al_out = al;
        return p < total_pairs ? seg0[(size_t)al * a.seg_cap + (p - c[al])] : 0u;
    };
    int al_next = 0;
    uint32_t B_next = entry((uint32_t)slot, al_next);
    uint32_t nbuf = 0;                                                // pending hits in the warp's buffer (kAtomic)

    for (uint32_t p0 = 0; p0 < total_pairs; p0 += SLOTS) {
        const uint32_t p = p0 + slot;
        const bool have = p < total_pairs;
        const int al = al_next;                                       // which of the warp's groups this entry belongs to
        const uint32_t B = B_next;
        B_next = entry(p + SLOTS, al_next);
        const uint32_t j0 = B * G;
        if (have && j0 + m < n_target) {
            // stage only the volume (index / parent box are fetched on a hit)
            const LT* tp = bvh.leaves + j0 + m;
            VT tv;
            const uint2* sp = reinterpret_cast<const uint2*>(tp);
            uint2* dp = reinterpret_cast<uint2*>(&tv);
#pragma unroll
            for (int k = 0; k < (int)(sizeof(VT) / 8); ++k) dp[k] = __ldg(sp + k);
            s_vol[w][slot][m].v = tv;
        }
        __syncwarp();
        uint32_t hits = 0;
        const int ql = al * G + m;                                     // query slot in shared memory
        const uint32_t qi32 = (uint32_t)(A0 * G) + (uint32_t)ql;       // query index within the shard
        const uint32_t qpos = (uint32_t)a.q_begin + qi32;              // query leaf position
        const bool q_ok = have && qi32 < (uint32_t)a.q_count && qpos < (uint32_t)n_query_total;
        if (q_ok) {
            const VQ qv = s_q[w][ql].v;
            uint32_t nval = n_target - j0;
            uint32_t allowed = nval >= (uint32_t)G ? ((1u << G) - 1u) : ((1u << nval) - 1u);
            if constexpr (KIND == kSingle) {
                if (qpos >= j0) {                                      // only leaves strictly right of the query
                    uint32_t lo = qpos - j0 + 1u;
                    allowed &= lo >= (uint32_t)G ? 0u : ~((1u << lo) - 1u);
                }
            }
#pragma unroll
            for (int j = 0; j < G; ++j) hits |= (leaf_contact(qv, s_vol[w][slot][j].v) ? 1u : 0u) << j;
            hits &= allowed;
            if (hits) {
                // rare path: the exact leaf-parent test of (*) with lazily fetched data
                const N qbox = NodeOps<N>::convert(qv);
                const int64_t par_base = bvh.ti.level_start[levels - 1];
                uint32_t cand = hits;
                while (cand) {
                    int j = __ffs(cand) - 1;
                    cand &= cand - 1;
                    N par = load_struct(bvh.nodes + par_base + ((j0 + j) >> 1));
                    if (!iscontact(qbox, par)) hits &= ~(1u << j);
                }
            }
        }
        if constexpr (MODE == kCount) {
            if (hits) atomicAdd(&s_cnt[w][ql], (uint32_t)__popc(hits));
        } else {
            // buffer (query position, target position) of every hit in shared memory; full 32-entry chunks are
            // flushed with ONE atomic and one coalesced 256-byte store (indices are fetched at flush time)
            unsigned any = __ballot_sync(0xffffffffu, hits != 0);
            while (any) {
                const bool hv = hits != 0;
                if (hv) {
                    int j = __ffs(hits) - 1;
                    hits &= hits - 1;
                    s_buf[w][nbuf + __popc(any & ((1u << lane) - 1u))] = make_uint2(qpos, j0 + (uint32_t)j);
                }
                nbuf += __popc(any);
                any = __ballot_sync(0xffffffffu, hits != 0);
            }
            __syncwarp();
            while (nbuf >= 32) {
                nbuf -= 32;
                flush(s_buf[w][nbuf + lane], 32u, true);
            }
        }
        __syncwarp();
    }
    if constexpr (MODE == kCount) {
        __syncwarp();
        if (my_valid && !my_flagged) counts[my_qi] = (I)s_cnt[w][lane];
    } else {
        __syncwarp();
        if (nbuf) flush(lane < nbuf ? s_buf[w][lane] : make_uint2(0u, 0u), nbuf, lane < nbuf);
    }
}

// inclusive scan of uint32 group counts reuses the reduce / scan / apply kernels of traverse.cuh with I = uint32_t

}  // namespace ibvh
