// traverse_pyramid.cuh — "pyramid refinement": the fastest schedule for LVT contact traversal over BBox
// nodes (single tree and BVH-vs-BVH). Same contact set as traverse_lvt_single! / traverse_lvt_pair!
// (src/traverse/leaf_vs_tree/traverse_single.jl:136-208, traverse_pair.jl:176-244), bit for bit.
//
// Exactness (see traverse_tile.cuh for the containment argument): with BBox nodes the reference's
// predicate is   P(q, j) = leaf_test(q, j) AND iscontact(BBox(q), parent_{levels-1}(j))   (+ j right of q
// for the single tree), and if P(q, j) holds then for EVERY granularity 2^k the exact union box of the
// query group containing q touches the tree node that covers the 2^k-leaf group containing j. So a
// top-down refinement over *pairs of groups* that only drops pairs whose boxes are disjoint never
// loses a contact, and the final 4 x 4 leaf tiles evaluate P itself.
//
// Schedule (no per-thread tree walk, no stacks, every work item is the same size):
//   0. aligned records: leaf volumes packed to 16-byte records (NaN padding past the end), the target node
//      levels copied to 64-byte aligned runs, the query pyramid stored as 32-byte boxes — the hot kernels are
//      bound by the LSU data pipe, and AoS structs read as 8-byte pieces cost 3-7x the wavefronts.
//   1. query pyramid: exact union boxes of the query leaves at group sizes 4, 32, 256, ... (min / max).
//   2. top: all pairs (query group, tree node) at the coarsest size (<= 2048 groups per side).
//   3. refine x (levels): each surviving pair (A, B) expands to its 8 x 8 child pairs over conservatively
//      quantised 15-bit boxes (a test = three integer subtractions); a warp handles 8 pairs per step, a lane
//      owns TWO children of A; the 8 child boxes of B are fetched cooperatively as aligned 16-byte pieces into
//      shared memory and read back as LDS.128 broadcasts (pyr_refine_q2_kernel; the float and one-child
//      forms are kept for A/B and for trees the quantisation does not cover).
//   4. leaf tiles: pairs of 4-leaf groups, 16 pairs per warp step, a lane owns TWO query leaves (register
//      tile 2 x 4), targets travel global -> shared with cp.async (double buffer, conflict-free mapping),
//      the sphere tests run on the packed FP32 pipe (FADD2 / FMUL2: traverse_tile.cuh), lazy leaf-parent
//      test on a hit. Pair lists, contact list and hit stash are read / written with evict-first hints.
// Work distribution: warps draw chunks of consecutive list entries from a ticket counter, so the spatial
// order of the top-level list survives from level to level. Survivors are appended to flat (A, B) lists
// through a per-warp shared-memory buffer (one shared atomic per lane with hits) that is flushed with one
// global atomic per >= 256 entries and coalesced stores. List sizes stay on the device (the next phase reads
// the count), so the whole traversal needs a single host synchronisation: the read-back of the contact
// total, as in the reference — or none at all in the deferred form (IBVH_TRAVERSE_DEFER).
// Output modes of the leaf-tile kernel: unordered append; fused multi-GPU append (slots inside this rank's
// region of the gathered list from a LOCAL counter, multimem.st into every rank's list); ordered = count per
// query + stash, then scan, scatter and a per-query sort staged through shared memory (pyr_scatter_kernel,
// pyr_fixup_kernel).
#pragma once
#include "peer.cuh"
#include <type_traits>

#include "common.cuh"
#include "traverse.cuh"
#include "traverse_tile.cuh"

namespace ibvh {

constexpr int kPyrWarps = 4;                // warps per CTA in the refine / tile kernels
constexpr int kPyrLeafLog = 2;              // finest groups: 4 leaves
#ifndef IBVH_PYR_FAN
#define IBVH_PYR_FAN 3
#endif
constexpr int kPyrFan = IBVH_PYR_FAN;       // 2^3 = 8 children per refinement
constexpr int kPyrMaxLevels = 12;
#ifndef IBVH_PYR_FLUSH
#define IBVH_PYR_FLUSH 256
#endif
constexpr int kPyrFlush = IBVH_PYR_FLUSH;   // buffered entries per list atomic

struct PyrLevel {
    int32_t k;               // group size 2^k leaves
    int32_t tree_level;      // levels - k: tree level whose nodes are the target groups
    int64_t ntg;             // target groups = real nodes on tree_level
    int64_t tnode0;          // memory position of the first node of tree_level
    int64_t qg_first;        // first query group (absolute leaf position >> k)
    int64_t nqg;             // query groups covering the shard
    int64_t u_off;           // offset (in boxes) of this level inside the query-pyramid array
    int64_t t_off;           // offset (in boxes, multiple of 8) of this level's nodes inside the aligned target copy
};

struct PairList {
    uint2* data;
    unsigned long long* count;     // device counter (keeps counting past `cap`: tells the host the need)
    unsigned long long cap;
};

#ifndef IBVH_PYR_AHEAD
#define IBVH_PYR_AHEAD 0
#endif
#ifndef IBVH_PYR_CHUNK_MAX
#define IBVH_PYR_CHUNK_MAX 32
#endif
// (measured on the leaf-tile kernel, 10 M leaves: chunk maximum 4 / 8 / 32 / 64 / 128 steps -> 1.33 / 1.28 / 1.25 / 1.29 / 1.30 ms)
// steps per ticket chunk: up to 32, fewer for short lists so that every resident warp still gets ~4 chunks
IBVH_D uint32_t pyr_chunk_steps(uint32_t count, int slots) {
    const uint32_t warps = gridDim.x * (uint32_t)kPyrWarps;
    uint32_t s = count / ((uint32_t)slots * warps * 4u);
    return s < 1u ? 1u : (s > (uint32_t)IBVH_PYR_CHUNK_MAX ? (uint32_t)IBVH_PYR_CHUNK_MAX : s);
}

// The pair lists (475 MB at the last refinement level) and the contact list (318 MB) are written once and read once or
// never by the GPU, while the 160 MB of leaf records and the node levels are re-read many times and almost fit the 126 MB L2:
// IBVH_STREAM_HINTS marks the list traffic evict-first (ld.global.cs / st.global.cs) so that it does not push them out
// (A/B of two builds in one gpurun call, twice each: tile 1.133 -> 1.116 ms, refine 0.677 -> 0.673, step 2.651 -> 2.632 ms).
#ifndef IBVH_STREAM_HINTS
#define IBVH_STREAM_HINTS 1
#endif
IBVH_D uint2 pyr_stream_load(const uint2* p) {
#if IBVH_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}
IBVH_D void pyr_stream_store(uint2* p, uint2 v) {
#if IBVH_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}
template <class I> IBVH_D void pyr_stream_store(IndexPair<I>* p, const IndexPair<I>& v) {
#if IBVH_STREAM_HINTS
    if constexpr (sizeof(I) == 4) __stcs(reinterpret_cast<uint2*>(p), make_uint2((uint32_t)v.a, (uint32_t)v.b));
    else __stcs(reinterpret_cast<ulonglong2*>(p), make_ulonglong2((unsigned long long)v.a, (unsigned long long)v.b));
#else
    *p = v;
#endif
}
// EXPERIMENT -DIBVH_L2_KEEP=<percent>: the leaf records of the tile kernel are fetched with an L2 evict-last policy for that
// fraction of the accesses (createpolicy.fractional + ld / cp.async ... L2::cache_hint).
#ifndef IBVH_L2_KEEP
#define IBVH_L2_KEEP 0
#endif
IBVH_D unsigned long long pyr_keep_policy() {
    unsigned long long pol = 0;
#if IBVH_L2_KEEP > 0
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol) : "f"((float)IBVH_L2_KEEP / 100.0f));
#endif
    return pol;
}
template <class R> IBVH_D R load16_keep(const R* p, unsigned long long pol) {
#if IBVH_L2_KEEP > 0
    static_assert(sizeof(R) % 16 == 0, "16-byte records");
    alignas(16) R out;
    uint4* d = reinterpret_cast<uint4*>(&out);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(R) / 16); ++k)
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(d[k].x), "=r"(d[k].y), "=r"(d[k].z), "=r"(d[k].w)
                     : "l"(reinterpret_cast<const uint4*>(p) + k), "l"(pol));
    return out;
#else
    (void)pol;
    return load16(p);
#endif
}
IBVH_D void atomic_inc(int32_t* p) { atomicAdd(p, 1); }
IBVH_D void atomic_inc(int64_t* p) { atomicAdd(reinterpret_cast<unsigned long long*>(p), 1ull); }

// ---- TMA bulk copies + mbarrier (sm_90+ PTX; SASS: UBLKCP / SYNCS) -----------------------------------------------
// One lane arms the warp's mbarrier with the bytes it expects (arrive.expect_tx) and issues
// cp.async.bulk.shared::cluster.global: the copy engine moves whole 16-byte-aligned runs from global to shared memory
// and completes the transaction count on the barrier; the consumers wait on the barrier's phase parity. No registers,
// no LDG -> STS round trip through the LSU data pipe (which ncu named as the limiter of the refine kernel: 93 %).
// MEASURED (B200, 10 M leaves, gpurun_out/r2g -> profiles/r2_tma_ab.txt), same results bit for bit:
//   refine: LDG -> STS 1.009 ms, TMA (192 + 256-byte runs, 8 copies per warp step)      1.205 ms
//   tiles : cp.async   1.302 ms, TMA (64 + 64-byte runs, 32 copies per warp step)       2.717 ms
// The runs this algorithm needs are a few hundred bytes at scattered addresses: at that size a bulk copy costs more
// in the copy engine's per-request overhead than the LSU round trip it saves, so both TMA forms are kept as OPT-IN
// variants (refine: IBVH_PYR_TMA=1 in the environment at ibvh_create; tiles: compile with -DIBVH_PYR_TMA=1) and the
// LDG / cp.async forms stay the default. Large contiguous tiles (where TMA wins) do not occur on this path.
#ifndef IBVH_PYR_TMA
#define IBVH_PYR_TMA 0
#endif
#ifndef IBVH_PYR_QPL
#define IBVH_PYR_QPL 2            // query leaves per lane in the leaf-tile kernel (register tile QPL x 4)
#endif
#ifndef IBVH_PYR_TILE_MINB
#define IBVH_PYR_TILE_MINB 8      // resident CTAs per SM the tile kernel's registers are capped for (8 -> 64 registers)
#endif
// Float64 volumes: half as many resident CTAs, twice the registers (BBox<double> leaves spilled ~300 bytes per thread at 64)
template <class L> constexpr int pyr_tile_minb() { return sizeof(typename L::value_type) == 8 ? (IBVH_PYR_TILE_MINB + 1) / 2 : IBVH_PYR_TILE_MINB; }
IBVH_D void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
IBVH_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
IBVH_D void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
IBVH_D void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy; dst / src 16-byte aligned, bytes a multiple of 16
IBVH_D void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
IBVH_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// leaf volumes -> 16-byte aligned records (one pass per traversal; 0.07 ms for 10 M sphere leaves)
// Records [n, n_pad) are all-ones bit patterns = NaN volumes: every comparison with them is false, so the tile
// kernel needs no bounds masks on the target side.
template <class L>
__global__ void __launch_bounds__(256) pyr_pack_volumes_kernel(const L* __restrict__ leaves, int64_t n, int64_t n_pad, Packed<typename L::vol_t>* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    alignas(16) Packed<typename L::vol_t> r;
    if (i < n) {
        memset(&r, 0, sizeof(r));
        r.v = load_struct(leaves + i).volume;
    } else {
        memset(&r, 0xFF, sizeof(r));
    }
    store16(out + i, r);
}

// pack + finest query-pyramid level in one pass over the leaves (one thread per 4-leaf group): the leaves are read
// once instead of twice. Groups inside the query shard [lv.qg_first, lv.qg_first + lv.nqg) also get their union box.
template <class L, class T>
__global__ void __launch_bounds__(256) pyr_pack_groups_kernel(const L* __restrict__ leaves, int64_t n, int64_t n_pad, Packed<typename L::vol_t>* __restrict__ out,
                                                             int64_t q_begin, int64_t q_end, PyrLevel lv, UBox<T>* __restrict__ U) {
    constexpr int G = 1 << kPyrLeafLog;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = g * G;
    if (i0 >= n_pad) return;
    BBox<T> u = empty_box<T>();
#pragma unroll
    for (int m = 0; m < G; ++m) {
        const int64_t i = i0 + m;
        if (i >= n_pad) break;
        alignas(16) Packed<typename L::vol_t> r;
        if (i < n) {
            memset(&r, 0, sizeof(r));
            r.v = load_struct(leaves + i).volume;
            if (i >= q_begin && i < q_end) u = merge(u, NodeOps<BBox<T>>::convert(r.v));
        } else {
            memset(&r, 0xFF, sizeof(r));
        }
        store16(out + i, r);
    }
    const int64_t t = g - lv.qg_first;
    if (t >= 0 && t < lv.nqg) U[lv.u_off + t].b = u;
}

// ---- 1. query pyramid ------------------------------------------------------------------------------------
// Uf / Uc: the fine / coarse level's boxes, index = group - level.qg_first (the fine level may live in a build's sidecar)
template <class T>
__global__ void __launch_bounds__(256) pyr_up_kernel(PyrLevel fine, PyrLevel coarse, const UBox<T>* __restrict__ Uf, UBox<T>* __restrict__ Uc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= coarse.nqg) return;
    const int64_t c0 = (coarse.qg_first + t) << kPyrFan;
    BBox<T> u = empty_box<T>();
#pragma unroll
    for (int i = 0; i < (1 << kPyrFan); ++i) {
        int64_t c = c0 + i - fine.qg_first;
        if (c >= 0 && c < fine.nqg) u = merge(u, load16(Uf + c).b);
    }
    Uc[t].b = u;
}

// warp-aggregated direct append (top kernel)
IBVH_D void pyr_append_direct(const PairList& out, bool pred, uint2 e) {
    unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(out.count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pred) {
        unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < out.cap) out.data[slot] = e;
    }
}

// ---- 2. top: all pairs ---------------------------------------------------------------------------------------
template <int KIND, class T>
// U: the top level's boxes, index = group - lv.qg_first
__global__ void __launch_bounds__(256) pyr_top_kernel(PyrLevel lv, const UBox<T>* __restrict__ U, const BBox<T>* __restrict__ nodes, PairList out) {
    const int64_t total = lv.nqg * lv.ntg;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < total; t0 += stride) {       // warp-uniform trip count
        const int64_t t = t0 + threadIdx.x;
        bool pred = false;
        uint2 e = make_uint2(0u, 0u);
        if (t < total) {
            const int64_t a = t / lv.ntg, b = t % lv.ntg;
            const int64_t A = lv.qg_first + a;
            bool ok = true;
            if constexpr (KIND == kSingle) ok = b >= A;           // the target group must reach right of the query group
            if (ok) {
                BBox<T> u = load16(U + a).b;
                BBox<T> nb = load_struct(nodes + lv.tnode0 + b);
                pred = iscontact(u, nb);
                e = make_uint2((uint32_t)A, (uint32_t)b);
            }
        }
        pyr_append_direct(out, pred, e);
    }
}

// Shared-memory hit buffer of one warp: entries are pushed by the lanes that have one, full 32-entry chunks
// are flushed with one atomic + one coalesced store.
template <int CAP> struct WarpBuf {
    uint2* buf;       // shared memory, CAP + 32 entries
    uint32_t n;
};

// ---- 3. refine: (A, B) at group size 2^k -> child pairs at 2^(k-3) ------------------------------------------------
// Uf: the query-pyramid boxes of the fine level (32-byte records, index = group - f_first).
// Nf: the fine level's target nodes inside the 64-byte aligned copy (whole groups of F readable), so that the F
//     children of a target group are one aligned run of 16-byte pieces fetched cooperatively by the slot's lanes.
template <int KIND, class T>
__global__ void __launch_bounds__(kPyrWarps * 32) pyr_refine_kernel(const UBox<T>* __restrict__ Uf, const BBox<T>* __restrict__ Nf,
                                                                   uint32_t f_first, uint32_t f_nqg, uint32_t f_ntg,
                                                                   PairList in, PairList out, uint32_t* ticket) {
    using N = BBox<T>;
    constexpr int F = 1 << kPyrFan;          // 8
    constexpr int SLOTS = 32 / F;            // 4 pairs per warp step
    constexpr int PIECES = F * (int)sizeof(N) / 16;                        // 16-byte pieces of the F child boxes of one target group
    constexpr int ROUNDS = (PIECES + F - 1) / F;
    constexpr int SLOT_BYTES = F * (int)sizeof(N) + 16;                    // + 16: slots start in different banks
    static_assert((F * sizeof(N)) % 16 == 0, "child run is a whole number of 16-byte pieces");
    __shared__ __align__(16) unsigned char s_raw[kPyrWarps][SLOTS][SLOT_BYTES];
    // hit buffer: up to 32*F new entries per step on top of < kPyrFlush pending ones
    __shared__ uint2 s_buf[kPyrWarps][32 * F + kPyrFlush];
    __shared__ uint32_t s_n[kPyrWarps];                                    // entries buffered per warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / F, i = lane % F;
    if (lane == 0) s_n[w] = 0;
    __syncwarp();
    unsigned long long count64 = *in.count;
    if (count64 > in.cap) count64 = in.cap;
    const uint32_t count = (uint32_t)count64;                              // list sizes are < 2^32 (checked by the host)
    uint32_t nbuf = 0;
    // flush the n newest buffered entries [nbuf - n, nbuf) with ONE atomic (all warps share one list counter, and
    // same-address atomics are what limits this kernel otherwise) and coalesced 256-byte stores
    auto flush = [&](uint32_t n) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out.count, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t b0 = nbuf - n;
        for (uint32_t k = lane; k < n; k += 32) if (base + k < out.cap) pyr_stream_store(out.data + base + k, s_buf[w][b0 + k]);
        nbuf = b0;
    };
    // Software pipeline: while step t is computed, the boxes of step t+1 and the list entry of step t+2 are in flight
    // (the kernel is bound by issue slots: 2x unrolled ping-pong stages instead of copying a stage per step, loads
    // from clamped indices instead of predicated loads + selects — `a_ok` / `allowed` mask what is not real).
    struct Stage { uint32_t Ac, Bc0; bool a_ok; UBox<T> u; uint4 tp[ROUNDS]; };
    auto fetch = [&](uint2 pr, bool have) -> Stage {
        Stage sg;
        sg.Ac = (pr.x << kPyrFan) + (uint32_t)i;                           // my child of A (absolute group index, fine level)
        sg.Bc0 = pr.y << kPyrFan;                                          // first child of B
        const uint32_t ua = sg.Ac - f_first;                               // wraps to a huge value if Ac < f_first
        sg.a_ok = have && ua < f_nqg;
        sg.u = load16(Uf + min(ua, f_nqg - 1u));
        // the 8 lanes of a slot fetch the F child boxes of B as PIECES consecutive 16-byte pieces (the copy is padded
        // to whole groups, so the run is always readable; boxes past f_ntg are masked by `allowed`)
        const uint4* run = reinterpret_cast<const uint4*>(Nf + sg.Bc0);
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int pc = r * F + i;
            sg.tp[r] = make_uint4(0u, 0u, 0u, 0u);
            if (pc < PIECES) sg.tp[r] = __ldg(run + pc);
        }
        return sg;
    };
    volatile uint32_t* s_nv = s_n;
    auto process = [&](const Stage& cur) {
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int pc = r * F + i;
            if (pc < PIECES) reinterpret_cast<uint4*>(s_raw[w][slot])[pc] = cur.tp[r];
        }
        __syncwarp();
        uint32_t hits = 0;
        if (cur.a_ok) {
            // two boxes = three 16-byte shared-memory loads
            static_assert(F % 2 == 0, "boxes are read in pairs");
#pragma unroll
            for (int j = 0; j < F; j += 2) {
                struct alignas(16) Two { N a, b; };
                static_assert(sizeof(Two) % 16 == 0, "pair of boxes");
                Two two;
                const uint4* sp = reinterpret_cast<const uint4*>(s_raw[w][slot] + j * sizeof(N));
                uint4* dp = reinterpret_cast<uint4*>(&two);
#pragma unroll
                for (int k = 0; k < (int)(sizeof(Two) / 16); ++k) dp[k] = sp[k];
                hits |= box_contact_bit(cur.u.b, two.a, 1u << j) | box_contact_bit(cur.u.b, two.b, 1u << (j + 1));
            }
            bool edge = cur.Bc0 + (uint32_t)F > f_ntg;
            if constexpr (KIND == kSingle) edge = edge || cur.Ac > cur.Bc0;
            if (edge) {
                const uint32_t nval = f_ntg - cur.Bc0;
                uint32_t allowed = nval >= (uint32_t)F ? ((1u << F) - 1u) : ((1u << nval) - 1u);
                if constexpr (KIND == kSingle) {
                    if (cur.Ac > cur.Bc0) {                                  // child j admissible iff Bc0 + j >= Ac
                        const uint32_t lo = cur.Ac - cur.Bc0;
                        allowed &= lo >= (uint32_t)F ? 0u : ~((1u << lo) - 1u);
                    }
                }
                hits &= allowed;
            }
        }
        // append: one shared-memory atomic per lane with hits reserves its slots in the warp's buffer
        if (hits) {
            uint32_t wpos = atomicAdd(&s_n[w], (uint32_t)__popc(hits));
            do {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                s_buf[w][wpos++] = make_uint2(cur.Ac, cur.Bc0 + (uint32_t)j);
            } while (hits);
        }
        __syncwarp();
        nbuf = s_nv[w];
        if (nbuf >= (uint32_t)kPyrFlush) { flush(nbuf & ~31u); __syncwarp(); if (lane == 0) s_nv[w] = nbuf; }
        __syncwarp();
    };
    // Work distribution: warps draw chunks of consecutive list entries from a ticket counter. A warp's output
    // flushes then hold the children of CONSECUTIVE parents, so the spatial order of the top-level list survives
    // from level to level (at flush granularity) and the leaf loads of the tile kernel hit in L1 / L2.
    const uint32_t chunk = SLOTS * pyr_chunk_steps(count, SLOTS);
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        const unsigned long long base64 = (unsigned long long)c * chunk;
        if (base64 >= count) break;
        const uint32_t base = (uint32_t)base64;
        const uint32_t end = count - base > chunk ? base + chunk : count;
        uint32_t p = base + slot;
        uint2 e1 = make_uint2(0u, 0u), e2 = make_uint2(0u, 0u);
        if (p < end) e1 = pyr_stream_load(in.data + p);
        if (p + SLOTS < end) e2 = pyr_stream_load(in.data + p + SLOTS);
        auto next_entry = [&]() {                                          // entry of the step after the next
            uint2 e = make_uint2(0u, 0u);
            if (p + 2 * SLOTS < end) e = pyr_stream_load(in.data + p + 2 * SLOTS);
            return e;
        };
        Stage sa = fetch(e1, p < end), sb;
        for (uint32_t p0 = base; p0 < end;) {
            sb = fetch(e2, p + SLOTS < end);                               // boxes of the next step
            e2 = next_entry();
            process(sa);
            p0 += SLOTS; p += SLOTS;
            if (p0 >= end) break;
            sa = fetch(e2, p + SLOTS < end);
            e2 = next_entry();
            process(sb);
            p0 += SLOTS; p += SLOTS;
        }
    }
    if (nbuf) flush(nbuf);
}

// ---- 3a. refine over conservatively quantised boxes -------------------------------------------------------------------
// The refinement levels only PRUNE: any test that never rejects a pair of overlapping boxes keeps the result exact (the
// leaf tiles evaluate the reference's predicate itself). So the refine kernel need not see the float boxes: it reads
// 15-bit integer boxes in the frame of the target tree's root box, rounded OUTWARDS (lo down, up up, clamped — every
// step monotone, hence conservative whatever the frame), packed so that ONE box test is three 32-bit subtractions:
//   target box  B' = (lo.x, lo.y | lo.z, M - up.x | M - up.y, M - up.z)          two 15-bit fields per word, M = 32767
//   query box   A' = (up.x, up.y | up.z, M - lo.x | M - lo.y, M - lo.z)  | G     G = 0x80008000: a guard bit above each field
//   overlap  <=>  every field of A' >= the same field of B'  <=>  no guard bit borrows in (A' | G) - B'
// 12 bytes per target box instead of 24 and 16 instead of 32 per query box: the 8 children of a target group are 6
// LDS.128 per lane instead of 12, one LDG.128 + STS.128 per slot lane instead of two, and a test is 3 IADD + 2 LOP3
// instead of 6 FSETP + SEL. ncu had the float kernel at 93 % of the LSU data pipe (shared-memory wavefronts).
struct QBoxT { uint32_t w[3]; };
struct alignas(16) QBoxU { uint32_t w[3]; uint32_t pad; };
constexpr uint32_t kQGuard = 0x80008000u;
constexpr float kQMax = 32767.0f;

template <class T> IBVH_D uint32_t quant_down(T v, T origin, T scale) {
    T x = floor((v - origin) * scale);
    if (!(x > T(0))) x = T(0);                    // (also NaN)
    if (x > T(kQMax)) x = T(kQMax);
    return (uint32_t)x;
}
template <class T> IBVH_D uint32_t quant_up(T v, T origin, T scale) {
    T x = ceil((v - origin) * scale);
    if (!(x > T(0))) x = T(0);
    if (x > T(kQMax)) x = T(kQMax);
    return (uint32_t)x;
}
// SRC = BBox<T> (aligned node copy -> QBoxT) or UBox<T> (query pyramid -> QBoxU); root = the target tree's root box
template <class T, class SRC, class DST>
__global__ void __launch_bounds__(256) pyr_quantize_kernel(const SRC* __restrict__ src, int64_t n, const BBox<T>* __restrict__ root, DST* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BBox<T> r = load_struct(root);
    BBox<T> b;
    if constexpr (std::is_same<SRC, BBox<T>>::value) b = load_struct(src + i); else b = load16(src + i).b;
    uint32_t lo[3], up[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T scale = T(kQMax) / (r.up[k] - r.lo[k]);
        lo[k] = quant_down(b.lo[k], r.lo[k], scale);
        up[k] = quant_up(b.up[k], r.lo[k], scale);
    }
    const uint32_t M = 32767u;
    if constexpr (std::is_same<DST, QBoxT>::value) {
        dst[i].w[0] = lo[0] | (lo[1] << 16);
        dst[i].w[1] = lo[2] | ((M - up[0]) << 16);
        dst[i].w[2] = (M - up[1]) | ((M - up[2]) << 16);
    } else {
        dst[i].w[0] = (up[0] | (up[1] << 16)) | kQGuard;
        dst[i].w[1] = (up[2] | ((M - lo[0]) << 16)) | kQGuard;
        dst[i].w[2] = ((M - lo[1]) | ((M - lo[2]) << 16)) | kQGuard;
        dst[i].pad = 0u;
    }
}

template <int KIND>
__global__ void __launch_bounds__(kPyrWarps * 32) pyr_refine_q_kernel(const QBoxU* __restrict__ Uf, const QBoxT* __restrict__ Nf,
                                                                     uint32_t f_first, uint32_t f_nqg, uint32_t f_ntg,
                                                                     PairList in, PairList out, uint32_t* ticket) {
    constexpr int F = 1 << kPyrFan;          // 8
    constexpr int SLOTS = 32 / F;            // 4 pairs per warp step
    constexpr int PIECES = F * (int)sizeof(QBoxT) / 16;                   // 6 sixteen-byte pieces = the 8 child boxes of one target group
    constexpr int SLOT_BYTES = F * (int)sizeof(QBoxT) + 16;               // 112: the four slots start in different banks
    static_assert(F == 8 && sizeof(QBoxT) == 12 && PIECES <= F, "piece mapping");
    __shared__ __align__(16) unsigned char s_raw[kPyrWarps][SLOTS][SLOT_BYTES];
    __shared__ uint2 s_buf[kPyrWarps][32 * F + kPyrFlush];
    __shared__ uint32_t s_n[kPyrWarps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / F, i = lane % F;
    if (lane == 0) s_n[w] = 0;
    __syncwarp();
    unsigned long long count64 = *in.count;
    if (count64 > in.cap) count64 = in.cap;
    const uint32_t count = (uint32_t)count64;
    uint32_t nbuf = 0;
    auto flush = [&](uint32_t n) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out.count, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t b0 = nbuf - n;
        for (uint32_t k = lane; k < n; k += 32) if (base + k < out.cap) pyr_stream_store(out.data + base + k, s_buf[w][b0 + k]);
        nbuf = b0;
    };
    struct Stage { uint32_t Ac, Bc0; bool a_ok; uint4 u; uint4 tp; };
    auto fetch = [&](uint2 pr, bool have) -> Stage {
        Stage sg;
        sg.Ac = (pr.x << kPyrFan) + (uint32_t)i;
        sg.Bc0 = pr.y << kPyrFan;
        const uint32_t ua = sg.Ac - f_first;                               // wraps to a huge value if Ac < f_first
        sg.a_ok = have && ua < f_nqg;
        sg.u = __ldg(reinterpret_cast<const uint4*>(Uf + min(ua, f_nqg - 1u)));
        sg.tp = make_uint4(0u, 0u, 0u, 0u);
        if (i < PIECES) sg.tp = __ldg(reinterpret_cast<const uint4*>(Nf + sg.Bc0) + i);     // (the copy is padded to whole groups)
        return sg;
    };
    volatile uint32_t* s_nv = s_n;
    auto process = [&](const Stage& cur) {
        if (i < PIECES) reinterpret_cast<uint4*>(s_raw[w][slot])[i] = cur.tp;
        __syncwarp();
        uint32_t hits = 0;
        if (cur.a_ok) {
            const uint4* sp = reinterpret_cast<const uint4*>(s_raw[w][slot]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {                                  // 3 pieces = 4 boxes
                const uint4 p0 = sp[3 * q], p1 = sp[3 * q + 1], p2 = sp[3 * q + 2];
                const uint32_t bw[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t r = (cur.u.x - bw[3 * j]) & (cur.u.y - bw[3 * j + 1]) & (cur.u.z - bw[3 * j + 2]);
                    if ((r & kQGuard) == kQGuard) hits |= 1u << (4 * q + j);
                }
            }
            bool edge = cur.Bc0 + (uint32_t)F > f_ntg;
            if constexpr (KIND == kSingle) edge = edge || cur.Ac > cur.Bc0;
            if (edge) {
                const uint32_t nval = f_ntg - cur.Bc0;
                uint32_t allowed = nval >= (uint32_t)F ? ((1u << F) - 1u) : ((1u << nval) - 1u);
                if constexpr (KIND == kSingle) {
                    if (cur.Ac > cur.Bc0) {                                  // child j admissible iff Bc0 + j >= Ac
                        const uint32_t lo = cur.Ac - cur.Bc0;
                        allowed &= lo >= (uint32_t)F ? 0u : ~((1u << lo) - 1u);
                    }
                }
                hits &= allowed;
            }
        }
        if (hits) {
            uint32_t wpos = atomicAdd(&s_n[w], (uint32_t)__popc(hits));
            do {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                s_buf[w][wpos++] = make_uint2(cur.Ac, cur.Bc0 + (uint32_t)j);
            } while (hits);
        }
        __syncwarp();
        nbuf = s_nv[w];
        if (nbuf >= (uint32_t)kPyrFlush) { flush(nbuf & ~31u); __syncwarp(); if (lane == 0) s_nv[w] = nbuf; }
        __syncwarp();
    };
    const uint32_t chunk = SLOTS * pyr_chunk_steps(count, SLOTS);
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        const unsigned long long base64 = (unsigned long long)c * chunk;
        if (base64 >= count) break;
        const uint32_t base = (uint32_t)base64;
        const uint32_t end = count - base > chunk ? base + chunk : count;
        uint32_t p = base + slot;
        uint2 e1 = make_uint2(0u, 0u), e2 = make_uint2(0u, 0u);
        if (p < end) e1 = pyr_stream_load(in.data + p);
        if (p + SLOTS < end) e2 = pyr_stream_load(in.data + p + SLOTS);
        auto next_entry = [&]() {
            uint2 e = make_uint2(0u, 0u);
            if (p + 2 * SLOTS < end) e = pyr_stream_load(in.data + p + 2 * SLOTS);
            return e;
        };
        Stage sa = fetch(e1, p < end), sb;
        for (uint32_t p0 = base; p0 < end;) {
            sb = fetch(e2, p + SLOTS < end);
            e2 = next_entry();
            process(sa);
            p0 += SLOTS; p += SLOTS;
            if (p0 >= end) break;
            sa = fetch(e2, p + SLOTS < end);
            e2 = next_entry();
            process(sb);
            p0 += SLOTS; p += SLOTS;
        }
    }
    if (nbuf) flush(nbuf);
}

// ---- 3a'. the same with TWO (or four) query children per lane ---------------------------------------------------------------------
// ncu has pyr_refine_q_kernel issue-bound (80 %) at 200 warp instructions per 256-test step, of which the tests are 58: the
// rest is per-step overhead (entry / box fetch, staging, masks, append, loop). Here a lane owns the children 2i and 2i+1 of A,
// a pair takes 4 lanes and a warp step covers 8 pairs = 512 tests: the target boxes read from shared memory serve two
// queries and the per-step overhead serves twice the tests (the move that took the leaf-tile kernel from 1193 M to 955 M
// instructions in round 1). Same pair lists as a set; the order inside a flush differs (the leaf tiles do not depend on it).
template <int KIND, int QPL>
__global__ void __launch_bounds__(kPyrWarps * 32) pyr_refine_q2_kernel(const QBoxU* __restrict__ Uf, const QBoxT* __restrict__ Nf,
                                                                      uint32_t f_first, uint32_t f_nqg, uint32_t f_ntg,
                                                                      PairList in, PairList out, uint32_t* ticket) {
    constexpr int F = 1 << kPyrFan;          // 8
    constexpr int LPP = F / QPL;             // lanes per pair: 4 (QPL = 2) or 2 (QPL = 4)
    constexpr int SLOTS = 32 / LPP;          // pairs per warp step: 8 or 16
    constexpr int PIECES = F * (int)sizeof(QBoxT) / 16;                   // 6
    constexpr int PPL = (PIECES + LPP - 1) / LPP;                         // 16-byte pieces of the target group a lane stages: 2 or 3
    constexpr int SLOT_BYTES = F * (int)sizeof(QBoxT) + 16;               // 112
    static_assert(F == 8 && sizeof(QBoxT) == 12 && (QPL == 2 || QPL == 4), "piece mapping");
    __shared__ __align__(16) unsigned char s_raw[kPyrWarps][SLOTS][SLOT_BYTES];
    __shared__ uint2 s_buf[kPyrWarps][32 * QPL * F + kPyrFlush];
    __shared__ uint32_t s_n[kPyrWarps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / LPP, i = lane % LPP;
    if (lane == 0) s_n[w] = 0;
    __syncwarp();
    unsigned long long count64 = *in.count;
    if (count64 > in.cap) count64 = in.cap;
    const uint32_t count = (uint32_t)count64;
    uint32_t nbuf = 0;
    auto flush = [&](uint32_t n) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out.count, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t b0 = nbuf - n;
        for (uint32_t k = lane; k < n; k += 32) if (base + k < out.cap) pyr_stream_store(out.data + base + k, s_buf[w][b0 + k]);
        nbuf = b0;
    };
    struct Stage { uint32_t Ac, Bc0, ok; uint4 u[QPL]; uint4 tp[PPL]; };
    auto fetch = [&](uint2 pr, bool have) -> Stage {
        Stage sg;
        sg.Ac = (pr.x << kPyrFan) + (uint32_t)(QPL * i);                   // my first child of A (the others follow it)
        sg.Bc0 = pr.y << kPyrFan;
        const uint32_t ua = sg.Ac - f_first;                               // wraps if Ac < f_first; ua + c wraps back when a later child is inside
        sg.ok = 0u;
#pragma unroll
        for (int c = 0; c < QPL; ++c) {
            if (have && ua + (uint32_t)c < f_nqg) sg.ok |= 1u << c;
            sg.u[c] = __ldg(reinterpret_cast<const uint4*>(Uf + min(ua + (uint32_t)c, f_nqg - 1u)));
        }
#pragma unroll
        for (int k = 0; k < PPL; ++k) {                                    // (the copy is padded to whole groups)
            sg.tp[k] = make_uint4(0u, 0u, 0u, 0u);
            if (i + k * LPP < PIECES) sg.tp[k] = __ldg(reinterpret_cast<const uint4*>(Nf + sg.Bc0) + i + k * LPP);
        }
        return sg;
    };
    volatile uint32_t* s_nv = s_n;
    auto process = [&](const Stage& cur) {
#pragma unroll
        for (int k = 0; k < PPL; ++k) if (i + k * LPP < PIECES) reinterpret_cast<uint4*>(s_raw[w][slot])[i + k * LPP] = cur.tp[k];
        __syncwarp();
        uint32_t hits = 0;                                                 // bits [8c, 8c + 8): targets hit by child c
        if (cur.ok) {
            const uint4* sp = reinterpret_cast<const uint4*>(s_raw[w][slot]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {                                  // 3 pieces = 4 boxes
                const uint4 p0 = sp[3 * q], p1 = sp[3 * q + 1], p2 = sp[3 * q + 2];
                const uint32_t bw[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int c = 0; c < QPL; ++c) {
                        const uint32_t r = (cur.u[c].x - bw[3 * j]) & (cur.u[c].y - bw[3 * j + 1]) & (cur.u[c].z - bw[3 * j + 2]);
                        if ((r & kQGuard) == kQGuard) hits |= 1u << (F * c + 4 * q + j);
                    }
                }
            }
            uint32_t keep = 0;
#pragma unroll
            for (int c = 0; c < QPL; ++c) if (cur.ok & (1u << c)) keep |= 0xffu << (F * c);
            bool edge = cur.Bc0 + (uint32_t)F > f_ntg;
            if constexpr (KIND == kSingle) edge = edge || cur.Ac + (uint32_t)(QPL - 1) > cur.Bc0;
            if (edge) {
                const uint32_t nval = f_ntg - cur.Bc0;
                const uint32_t inside = nval >= (uint32_t)F ? ((1u << F) - 1u) : ((1u << nval) - 1u);
                uint32_t allowed = 0;
#pragma unroll
                for (int c = 0; c < QPL; ++c) {
                    uint32_t a = inside;
                    if constexpr (KIND == kSingle) {                         // child j of B admissible for child c of A iff Bc0 + j >= Ac + c
                        if (cur.Ac + (uint32_t)c > cur.Bc0) { const uint32_t lo = cur.Ac + (uint32_t)c - cur.Bc0; a &= lo >= (uint32_t)F ? 0u : ~((1u << lo) - 1u); }
                    }
                    allowed |= a << (F * c);
                }
                keep &= allowed;
            }
            hits &= keep;
        }
        if (hits) {
            uint32_t wpos = atomicAdd(&s_n[w], (uint32_t)__popc(hits));
            do {
                const int b = __ffs(hits) - 1;
                hits &= hits - 1;
                s_buf[w][wpos++] = make_uint2(cur.Ac + (uint32_t)(b >> kPyrFan), cur.Bc0 + (uint32_t)(b & (F - 1)));
            } while (hits);
        }
        __syncwarp();
        nbuf = s_nv[w];
        if (nbuf >= (uint32_t)kPyrFlush) { flush(nbuf & ~31u); __syncwarp(); if (lane == 0) s_nv[w] = nbuf; }
        __syncwarp();
    };
    const uint32_t chunk = SLOTS * pyr_chunk_steps(count, SLOTS);
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        const unsigned long long base64 = (unsigned long long)c * chunk;
        if (base64 >= count) break;
        const uint32_t base = (uint32_t)base64;
        const uint32_t end = count - base > chunk ? base + chunk : count;
        uint32_t p = base + slot;
        uint2 e1 = make_uint2(0u, 0u), e2 = make_uint2(0u, 0u);
        if (p < end) e1 = pyr_stream_load(in.data + p);
        if (p + SLOTS < end) e2 = pyr_stream_load(in.data + p + SLOTS);
        auto next_entry = [&]() {
            uint2 e = make_uint2(0u, 0u);
            if (p + 2 * SLOTS < end) e = pyr_stream_load(in.data + p + 2 * SLOTS);
            return e;
        };
        Stage sa = fetch(e1, p < end), sb;
        for (uint32_t p0 = base; p0 < end;) {
            sb = fetch(e2, p + SLOTS < end);
            e2 = next_entry();
            process(sa);
            p0 += SLOTS; p += SLOTS;
            if (p0 >= end) break;
            sa = fetch(e2, p + SLOTS < end);
            e2 = next_entry();
            process(sb);
            p0 += SLOTS; p += SLOTS;
        }
    }
    if (nbuf) flush(nbuf);
}

// ---- 3b. refine, TMA form ------------------------------------------------------------------------------------------
// Same work, same pair lists as pyr_refine_kernel. Per warp step (4 pairs) ONE lane arms the stage's mbarrier and four
// lanes issue two bulk copies each: the 8 child boxes of B (one aligned run of F * sizeof(N) bytes in the aligned node
// copy) and the 8 query-pyramid boxes of A's children (one run of F * sizeof(UBox) bytes; the pyramid levels carry F
// boxes of padding on both sides so that the run of a group cut by the shard range is still readable). The boxes are
// read back from shared memory: B's as LDS.128 broadcasts, the lane's own A child as two LDS.128. Two stages per warp:
// the copies of step t+1 fly while step t is tested. Against the LDG -> STS form this drops every global load and
// shared store of box data from the LSU instruction stream.
template <int KIND, class T>
__global__ void __launch_bounds__(kPyrWarps * 32, 8) pyr_refine_tma_kernel(const UBox<T>* __restrict__ Uf, const BBox<T>* __restrict__ Nf,
                                                                       uint32_t f_first, uint32_t f_nqg, uint32_t f_ntg,
                                                                       PairList in, PairList out, uint32_t* ticket) {
    using N = BBox<T>;
    constexpr int F = 1 << kPyrFan;          // 8
    constexpr int SLOTS = 32 / F;            // 4 pairs per warp step
    constexpr int NRUN = F * (int)sizeof(N);                 // bytes of B's child boxes
    constexpr int URUN = F * (int)sizeof(UBox<T>);           // bytes of A's child boxes
    constexpr int SLOT_BYTES = NRUN + URUN + 16;             // + 16: neighbouring slots start in different banks
    static_assert(NRUN % 16 == 0 && URUN % 16 == 0, "bulk copies move 16-byte multiples");
    __shared__ __align__(128) unsigned char s_raw[kPyrWarps][2][SLOTS][SLOT_BYTES];
    __shared__ __align__(8) unsigned long long s_bar[kPyrWarps][2];
    __shared__ uint2 s_buf[kPyrWarps][32 * F + kPyrFlush];
    __shared__ uint32_t s_n[kPyrWarps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / F, i = lane % F;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[w][0]);
    const uint32_t raw0 = (uint32_t)__cvta_generic_to_shared(&s_raw[w][0][0][0]);
    if (lane == 0) { s_n[w] = 0; mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_fence_init(); }
    __syncwarp();
    unsigned long long count64 = *in.count;
    if (count64 > in.cap) count64 = in.cap;
    const uint32_t count = (uint32_t)count64;
    uint32_t nbuf = 0;
    uint32_t parity = 0;                                      // bit s = phase parity the next wait on stage s expects
    auto flush = [&](uint32_t n) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out.count, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t b0 = nbuf - n;
        for (uint32_t k = lane; k < n; k += 32) if (base + k < out.cap) pyr_stream_store(out.data + base + k, s_buf[w][b0 + k]);
        nbuf = b0;
    };
    struct Stage { uint32_t Ac, Bc0; bool a_ok; };
    // issue the copies of one step into stage `st` (warp-uniform call; `have` per lane)
    auto fetch = [&](uint2 pr, bool have, int st) -> Stage {
        Stage sg;
        sg.Ac = (pr.x << kPyrFan) + (uint32_t)i;
        sg.Bc0 = pr.y << kPyrFan;
        const uint32_t ua = sg.Ac - f_first;                               // wraps to a huge value if Ac < f_first
        sg.a_ok = have && ua < f_nqg;
        const unsigned issuers = __ballot_sync(0xffffffffu, have && i == 0);
        if (issuers) {
            __syncwarp();                                                  // every lane is done reading this stage (two steps ago)
            if (lane == 0) { fence_proxy_async(); mbar_arrive_expect_tx(bar0 + 8 * st, (uint32_t)__popc(issuers) * (uint32_t)(NRUN + URUN)); }
            __syncwarp();
            if (have && i == 0) {
                const uint32_t dst = raw0 + (uint32_t)((st * SLOTS + slot) * SLOT_BYTES);
                tma_bulk_g2s(dst, Nf + sg.Bc0, (uint32_t)NRUN, bar0 + 8 * st);
                // A's children: boxes (A << fan) - f_first .. + F of this level (>= -F + 1 thanks to the padding)
                const long long u0 = (long long)(pr.x << kPyrFan) - (long long)f_first;
                tma_bulk_g2s(dst + (uint32_t)NRUN, Uf + u0, (uint32_t)URUN, bar0 + 8 * st);
            }
        }
        return sg;
    };
    volatile uint32_t* s_nv = s_n;
    auto process = [&](const Stage& cur, int st) {
        mbar_wait(bar0 + 8 * st, (parity >> st) & 1u);
        parity ^= 1u << st;
        const unsigned char* sbase = s_raw[w][st][slot];
        uint32_t hits = 0;
        if (cur.a_ok) {
            alignas(16) UBox<T> u;
            {
                const uint4* sp = reinterpret_cast<const uint4*>(sbase + NRUN + i * (int)sizeof(UBox<T>));
                uint4* dp = reinterpret_cast<uint4*>(&u);
#pragma unroll
                for (int k = 0; k < (int)(sizeof(UBox<T>) / 16); ++k) dp[k] = sp[k];
            }
            static_assert(F % 2 == 0, "boxes are read in pairs");
#pragma unroll
            for (int j = 0; j < F; j += 2) {
                struct alignas(16) Two { N a, b; };
                static_assert(sizeof(Two) % 16 == 0, "pair of boxes");
                Two two;
                const uint4* sp = reinterpret_cast<const uint4*>(sbase + j * sizeof(N));
                uint4* dp = reinterpret_cast<uint4*>(&two);
#pragma unroll
                for (int k = 0; k < (int)(sizeof(Two) / 16); ++k) dp[k] = sp[k];
                hits |= box_contact_bit(u.b, two.a, 1u << j) | box_contact_bit(u.b, two.b, 1u << (j + 1));
            }
            bool edge = cur.Bc0 + (uint32_t)F > f_ntg;
            if constexpr (KIND == kSingle) edge = edge || cur.Ac > cur.Bc0;
            if (edge) {
                const uint32_t nval = f_ntg - cur.Bc0;
                uint32_t allowed = nval >= (uint32_t)F ? ((1u << F) - 1u) : ((1u << nval) - 1u);
                if constexpr (KIND == kSingle) {
                    if (cur.Ac > cur.Bc0) {
                        const uint32_t lo = cur.Ac - cur.Bc0;
                        allowed &= lo >= (uint32_t)F ? 0u : ~((1u << lo) - 1u);
                    }
                }
                hits &= allowed;
            }
        }
        if (hits) {
            uint32_t wpos = atomicAdd(&s_n[w], (uint32_t)__popc(hits));
            do {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                s_buf[w][wpos++] = make_uint2(cur.Ac, cur.Bc0 + (uint32_t)j);
            } while (hits);
        }
        __syncwarp();
        nbuf = s_nv[w];
        if (nbuf >= (uint32_t)kPyrFlush) { flush(nbuf & ~31u); __syncwarp(); if (lane == 0) s_nv[w] = nbuf; }
        __syncwarp();
    };
    const uint32_t chunk = SLOTS * pyr_chunk_steps(count, SLOTS);
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        const unsigned long long base64 = (unsigned long long)c * chunk;
        if (base64 >= count) break;
        const uint32_t base = (uint32_t)base64;
        const uint32_t end = count - base > chunk ? base + chunk : count;
        uint32_t p = base + slot;
        uint2 e1 = make_uint2(0u, 0u), e2 = make_uint2(0u, 0u);
        if (p < end) e1 = pyr_stream_load(in.data + p);
        if (p + SLOTS < end) e2 = pyr_stream_load(in.data + p + SLOTS);
        auto next_entry = [&]() {
            uint2 e = make_uint2(0u, 0u);
            if (p + 2 * SLOTS < end) e = pyr_stream_load(in.data + p + 2 * SLOTS);
            return e;
        };
        Stage sa = fetch(e1, p < end, 0), sb;
        for (uint32_t p0 = base; p0 < end;) {
            sb = fetch(e2, p + SLOTS < end, 1);                            // (no copies, no barrier traffic when no slot has an entry)
            e2 = next_entry();
            process(sa, 0);
            p0 += SLOTS; p += SLOTS;
            if (p0 >= end) break;
            sa = fetch(e2, p + SLOTS < end, 0);
            e2 = next_entry();
            process(sb, 1);
            p0 += SLOTS; p += SLOTS;
        }
    }
    if (nbuf) flush(nbuf);
}

// ---- 4. leaf tiles ------------------------------------------------------------------------------------------------
// MODE kAtomic: append contacts (unordered). kCount: only add the number of contacts to *total.
// PMODE (ordered protocol): 0 = none; 1 = count per query (atomicAdd counts[qi]); 2 = write (qpos, tpos) into the
// query's segment via a per-query cursor (fixed up into reference order by pyr_fixup_kernel).
// FLUSH: buffered hits per output-slot reservation (one device-scope atomic on this rank's own counter, also in the fused
// multi-GPU mode: every rank appends inside its own region of the list).
template <int KIND, int MODE, int PMODE, class LQ, class LT, class I, int FLUSH = kPyrFlush>
__global__ void __launch_bounds__(kPyrWarps * 32, pyr_tile_minb<LQ>()) pyr_leaf_tile_kernel(const LQ* __restrict__ qleaves, int64_t q_begin, int64_t q_end,
                                                                      DBvh<LT, BBox<typename LT::value_type>> bvh, PairList in, int flip,
                                                                      int64_t capacity, unsigned long long* total,
                                                                      I* counts, unsigned int* cursors, IndexPair<I>* contacts, int fused, uint32_t* ticket,
                                                                      const Packed<typename LQ::vol_t>* __restrict__ pq,
                                                                      const Packed<typename LT::vol_t>* __restrict__ pt,
                                                                      const I* __restrict__ qidx_arr, const I* __restrict__ tidx_arr) {
    // qidx_arr / tidx_arr: the leaves' indices as compact arrays (a build's sidecar) or nullptr (read them from the leaf structs)
    auto q_index = [&](uint32_t pos) -> I { return qidx_arr ? __ldg(qidx_arr + pos) : (I)qleaves[pos].index; };
    auto t_index = [&](uint32_t pos) -> I { return tidx_arr ? __ldg(tidx_arr + pos) : (I)bvh.leaves[pos].index; };
    // pq / pt: the query / target leaf volumes as 16-byte aligned records (pyr_pack_volumes_kernel)
    // fused != 0 (multi-GPU, atomic mode): `contacts` is the NVSwitch multicast alias of this rank's region of every
    // rank's list (multimem.st) and `capacity` the region's size; `total` is a local counter as in the other modes
    using T = typename LT::value_type;
    using N = BBox<T>;
    using VT = typename LT::vol_t;
    using VQ = typename LQ::vol_t;
    // Register tile: every lane tests QPL = 2 query leaves against the G = 4 target leaves of its pair, so a pair
    // takes 2 lanes and a warp step covers 16 pairs. The kernel is bound by the LSU data pipe (shared-memory
    // wavefronts) and by issue slots, and both costs per pair halve against one query per lane: the target
    // volumes read from shared memory serve two queries, the per-step overhead serves 16 pairs.
    constexpr int G = 1 << kPyrLeafLog;      // 4
    constexpr int QPL = sizeof(Packed<VT>) <= 16 ? IBVH_PYR_QPL : 2;      // query leaves per lane (larger volumes: shared-memory budget)
    constexpr int LPP = G / QPL;             // lanes per pair
    constexpr int SLOTS = 32 / LPP;          // 16 pairs per warp step
    constexpr bool kNeedParent = !std::is_same<VT, N>::value || !std::is_same<VQ, N>::value;   // box leaves: implied by the leaf test
    using TVol = Packed<VT>;
    constexpr int TPIECES = G * (int)sizeof(TVol) / 16;                    // 16-byte pieces of one target group
    // target volumes of the current and the next step (double buffer). TMA form: one 64-byte bulk copy per target group and
    // one per query group, completing on the stage's mbarrier; else cp.async with the conflict-free piece mapping below.
    constexpr bool kTma = IBVH_PYR_TMA != 0 && sizeof(TVol) <= 16 && sizeof(Packed<VQ>) <= 16;      // (16-byte sphere records; larger volumes keep cp.async: their double-buffered slots + query groups would not fit 48 KB)
    constexpr int QBYTES = kTma ? G * (int)sizeof(Packed<VQ>) : 0;          // TMA form: the slot also holds the query group
    constexpr int SLOT_BYTES = G * (int)sizeof(TVol) + QBYTES + 16;        // + 16: neighbouring slots start in different banks
    __shared__ __align__(128) unsigned char s_vraw[kPyrWarps][2][SLOTS][SLOT_BYTES];
    __shared__ __align__(8) unsigned long long s_bar[kPyrWarps][2];
    __shared__ uint2 s_buf[kPyrWarps][32 * QPL * G + FLUSH];               // a step appends <= 32 * QPL * G entries
    __shared__ uint32_t s_n[kPyrWarps];                                    // entries buffered per warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot = lane / LPP, i = lane % LPP;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[w][0]);
    if (lane == 0) {
        s_n[w] = 0;
        if constexpr (kTma) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_fence_init(); }
    }
    __syncwarp();
    uint32_t parity = 0;                                                   // TMA form: bit s = phase the next wait on stage s expects
    const bool positions = (flip & 2) != 0;          // bit 1 of `flip`: report 1-based leaf positions instead of .index
    flip &= 1;
    const uint32_t n_target = (uint32_t)bvh.ti.n;
    const N* __restrict__ parents = bvh.nodes + bvh.ti.level_start[bvh.ti.levels - 1];
    const uint32_t qb32 = (uint32_t)q_begin, qe32 = (uint32_t)q_end;
    unsigned long long count64 = *in.count;
    if (count64 > in.cap) count64 = in.cap;
    const uint32_t count = (uint32_t)count64;
    uint32_t nbuf = 0;
    unsigned long long ncount = 0;

    // Dense rare path: the n newest buffered candidates (qpos, tpos) passed the leaf test. In rounds of 32, each lane
    // re-loads its query, applies the leaf-parent test of (*) and keeps the survivors in place; then ONE atomic
    // reserves the output range for all survivors (same-address atomics are the limiter otherwise).
    auto flush = [&](uint32_t n) {
        const uint32_t b0 = nbuf - n;
        uint32_t kept = 0;                                                   // survivors compacted to s_buf[b0 .. b0+kept)
        for (uint32_t r = 0; r < n; r += 32) {
            const bool mine = r + lane < n;
            uint2 e = make_uint2(0u, 0u);
            if (mine) e = s_buf[w][b0 + r + lane];
            bool ok = mine;
            if constexpr (kNeedParent) {
                if (mine) {
                    const Packed<VQ> qv = load16(pq + e.x);
                    ok = iscontact(NodeOps<N>::convert(qv.v), load_struct(parents + (e.y >> 1)));
                }
            }
            if constexpr (PMODE == 1) {
                if (ok) atomic_inc(&counts[e.x - qb32]);
            } else if constexpr (PMODE == 2) {
                if (ok) {
                    const uint32_t qi = e.x - qb32;
                    const int64_t seg = qi == 0 ? 0 : (int64_t)counts[qi - 1];
                    const unsigned int rr = atomicAdd(&cursors[qi], 1u);
                    // (target index, target position) for now — the leaf is still warm in L1 / L2 here, so the index
                    // is fetched now; pyr_fixup_kernel sorts the segment by position and writes the reported pair
                    contacts[seg + rr] = IndexPair<I>{positions ? (I)(e.y + 1u) : t_index(e.y), (I)e.y};
                }
            } else {
                if constexpr (PMODE == 3) { if (ok) atomic_inc(&counts[e.x - qb32]); }
                const unsigned om = __ballot_sync(0xffffffffu, ok);
                __syncwarp();
                if (ok) s_buf[w][b0 + kept + __popc(om & ((1u << lane) - 1u))] = e;    // kept <= r: never overtakes the reads
                kept += __popc(om);
                __syncwarp();
            }
        }
        nbuf = b0;
        if constexpr (PMODE == 3) {
            // ordered protocol in ONE tile pass: the survivors are counted per query (above) and stashed, unordered,
            // as (query, target position, target index); pyr_scatter_kernel moves them into the query segments
            // once the counts are scanned. `contacts` is the stash here, `capacity` its size.
            if (kept) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(total, (unsigned long long)kept);
                base = __shfl_sync(0xffffffffu, base, 0);
                uint4* stash = reinterpret_cast<uint4*>(contacts);
                for (uint32_t k = lane; k < kept; k += 32) {
                    const uint2 e = s_buf[w][b0 + k];
                    const unsigned long long li = positions ? (unsigned long long)(e.y + 1u) : (unsigned long long)(long long)t_index(e.y);
                    if ((int64_t)(base + k) < capacity) {
                        const uint4 sv = make_uint4(e.x - qb32, e.y, (uint32_t)li, (uint32_t)(li >> 32));
#if IBVH_STREAM_HINTS
                        __stcs(stash + base + k, sv);
#else
                        stash[base + k] = sv;
#endif
                    }
                }
            }
        }
        if constexpr (PMODE == 0) {
            if constexpr (MODE == kCount) {
                ncount += kept;
            } else if (kept) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(total, (unsigned long long)kept);      // (fused: a LOCAL counter too; the slots index this rank's region)
                base = __shfl_sync(0xffffffffu, base, 0);
                auto to_pair = [&](uint2 e) -> IndexPair<I> {
                    const I qidx = positions ? (I)(e.x + 1u) : q_index(e.x);
                    const I li = positions ? (I)(e.y + 1u) : t_index(e.y);
                    I ea, eb;
                    if constexpr (KIND == kSingle) { if (qidx > li) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
                    else { if (flip) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
                    return IndexPair<I>{ea, eb};
                };
                if (fused && sizeof(IndexPair<I>) == 8 && (int64_t)(base + kept) <= capacity) {
                    // multicast stores do not coalesce across lanes: two contacts per 16-byte multimem.st
                    const uint32_t head = (uint32_t)(base & 1ull);
                    const uint32_t npair = (kept - head) >> 1;
                    for (uint32_t k = lane; k < npair; k += 32) {
                        const IndexPair<I> p0 = to_pair(s_buf[w][b0 + head + 2 * k]), p1 = to_pair(s_buf[w][b0 + head + 2 * k + 1]);
                        uint4 v;
                        v.x = (uint32_t)p0.a; v.y = (uint32_t)p0.b; v.z = (uint32_t)p1.a; v.w = (uint32_t)p1.b;
                        multimem_st_v4(contacts + base + head + 2 * k, v);
                    }
                    if (lane == 0 && head) multimem_store_pair(contacts + base, to_pair(s_buf[w][b0]));
                    if (lane == 1 && ((kept - head) & 1u)) multimem_store_pair(contacts + base + kept - 1, to_pair(s_buf[w][b0 + kept - 1]));
                } else {
                    for (uint32_t k = lane; k < kept; k += 32) {
                        const IndexPair<I> pr = to_pair(s_buf[w][b0 + k]);
                        const unsigned long long wp = base + k;
                        if ((int64_t)wp < capacity) {
                            if (fused) multimem_store_pair(contacts + wp, pr);
                            else pyr_stream_store(contacts + wp, pr);
                        }
                    }
                }
            }
        }
    };

    // Software pipeline: the target volumes of step t+1 travel global -> shared with cp.async (no registers), the
    // query volumes of step t+1 and the list entry of step t+2 are loaded into registers while step t is computed.
    constexpr int SPI = 128 / TPIECES;                                     // slots per copy instruction half (quarter-warp pairs)
    constexpr int CP_ROUNDS = SLOTS * TPIECES / 32;                        // copy instructions per step
    static_assert(TPIECES == 4 || TPIECES == 8 || TPIECES == 12, "piece mapping below");
    // piece `pi` (0 .. SLOTS * TPIECES) -> (slot, piece in slot). 16-byte volumes: lane L of round k copies piece L % 4
    // of slot 8k + L / 8 + 4 * ((L / 4) % 2). Larger volumes: plain order (a slot is >= 128 contiguous bytes).
    auto copy_slot = [&](int k) -> int {
        if constexpr (TPIECES == 4) return 8 * k + (lane >> 3) + 4 * ((lane >> 2) & 1);
        else return (k * 32 + lane) / TPIECES;
    };
    auto copy_piece = [&](int k) -> int {
        if constexpr (TPIECES == 4) return lane & 3;
        else return (k * 32 + lane) % TPIECES;
    };
    const uint32_t vraw_base = (uint32_t)__cvta_generic_to_shared(s_vraw[w][0][0]);
    const unsigned long long keep_pol = pyr_keep_policy();
    (void)keep_pol;
    struct Stage { uint32_t qpos0, j0, have; Packed<VQ> q[QPL]; };
    auto fetch = [&](uint2 pr, bool have, int buf) -> Stage {
        Stage sg;
        sg.qpos0 = pr.x << kPyrLeafLog;
        sg.j0 = pr.y << kPyrLeafLog;
        sg.have = have ? 1u : 0u;
        if constexpr (kTma) {
            // one lane of each pair issues two bulk copies: the target group and the query group (64 bytes each for spheres)
            const unsigned issuers = __ballot_sync(0xffffffffu, have && i == 0);
            if (issuers) {
                __syncwarp();                                              // every lane is done reading this stage
                if (lane == 0) { fence_proxy_async(); mbar_arrive_expect_tx(bar0 + 8 * buf, (uint32_t)__popc(issuers) * (uint32_t)(G * sizeof(TVol) + QBYTES)); }
                __syncwarp();
                if (have && i == 0) {
                    const uint32_t dst = vraw_base + (uint32_t)((buf * SLOTS + slot) * SLOT_BYTES);
                    tma_bulk_g2s(dst, pt + sg.j0, (uint32_t)(G * sizeof(TVol)), bar0 + 8 * buf);
                    tma_bulk_g2s(dst + (uint32_t)(G * sizeof(TVol)), pq + sg.qpos0, (uint32_t)QBYTES, bar0 + 8 * buf);
                }
            }
            return sg;
        }
#pragma unroll
        for (int k = 0; k < QPL; ++k) sg.q[k] = load16_keep(pq + sg.qpos0 + (uint32_t)(QPL * i + k), keep_pol);
        // Target copy global -> shared, 16 bytes per lane and instruction. The copy mapping is NOT the compute mapping:
        // for 16-byte volumes one instruction moves 8 whole slots, lanes 4c .. 4c+3 of a quarter-warp writing the
        // 64 bytes of slot a and the next four lanes those of slot a + 4 — with the 80-byte slot stride these are
        // the slot pairs whose bank ranges do not overlap (the compute mapping gave 14 wavefronts per copy, this 4).
        // The T group of the copied slot comes from the lane that owns the pair, by shuffle.
#pragma unroll
        for (int k = 0; k < CP_ROUNDS; ++k) {
            const int cslot = copy_slot(k);
            const uint32_t tj0 = __shfl_sync(0xffffffffu, sg.j0, cslot * LPP);
            const uint4* src = reinterpret_cast<const uint4*>(pt + tj0) + copy_piece(k);
            const uint32_t dst = vraw_base + (uint32_t)(buf * SLOTS * SLOT_BYTES + cslot * SLOT_BYTES + 16 * copy_piece(k));
#if IBVH_L2_KEEP > 0
            asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(keep_pol) : "memory");
#else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        return sg;
    };
    volatile uint32_t* s_nv = s_n;
    // append: bits [G * k, G * k + G) of `hits` are the targets hit by query k of the lane; one shared-memory atomic
    // reserves the lane's slots
    auto append = [&](uint32_t hits, uint32_t qp0, uint32_t j0) {
        if (hits) {
            uint32_t wpos = atomicAdd(&s_n[w], (uint32_t)__popc(hits));
            do {
                const int b = __ffs(hits) - 1;
                hits &= hits - 1;
                s_buf[w][wpos++] = make_uint2(qp0 + (uint32_t)(b / G), j0 + (uint32_t)(b % G));
            } while (hits);
        }
        __syncwarp();
        nbuf = s_nv[w];
        if (nbuf >= (uint32_t)FLUSH) { flush(nbuf & ~31u); __syncwarp(); if (lane == 0) s_nv[w] = nbuf; }
        __syncwarp();
    };
    auto process = [&](Stage& cur, int buf) {
        if constexpr (kTma) {
            mbar_wait(bar0 + 8 * buf, (parity >> buf) & 1u);
            parity ^= 1u << buf;
            if (cur.have) {
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    const uint4* sp = reinterpret_cast<const uint4*>(s_vraw[w][buf][slot] + G * sizeof(TVol) + (QPL * i + k) * sizeof(Packed<VQ>));
                    uint4* dp = reinterpret_cast<uint4*>(&cur.q[k]);
#pragma unroll
                    for (int x = 0; x < (int)(sizeof(Packed<VQ>) / 16); ++x) dp[x] = sp[x];
                }
            }
        } else {
            asm volatile("cp.async.wait_group 1;" ::: "memory");               // this step's targets have landed (next step's may be in flight)
            __syncwarp();
        }
        uint32_t hits[QPL];
#pragma unroll
        for (int k = 0; k < QPL; ++k) hits[k] = 0;
        // (BSphere{Float32}: leaf_contact is the packed-FP32 form, traverse_tile.cuh — FADD2 / FMUL2 straight on the register
        // pairs of the 16-byte loads, 7 instructions per test instead of 11: 1.229 -> 1.121 ms at 10 M leaves)
#pragma unroll
        for (int j = 0; j < G; ++j) {
            alignas(16) TVol tv;
            const uint4* sp = reinterpret_cast<const uint4*>(s_vraw[w][buf][slot] + j * sizeof(TVol));
            uint4* dp = reinterpret_cast<uint4*>(&tv);
#pragma unroll
            for (int k = 0; k < (int)(sizeof(TVol) / 16); ++k) dp[k] = sp[k];
#pragma unroll
            for (int k = 0; k < QPL; ++k) if (leaf_contact(cur.q[k].v, tv.v)) hits[k] |= 1u << j;
        }
        const uint32_t qp0 = cur.qpos0 + (uint32_t)(QPL * i);
        // rare masks: tail lanes of a chunk, query groups cut by the shard range, and (single tree) the diagonal
        // tile, where only targets strictly right of the query count. Targets past the end are NaN padding.
        bool edge = !cur.have || cur.qpos0 < qb32 || cur.qpos0 + (uint32_t)G > qe32;
        if constexpr (KIND == kSingle) edge = edge || cur.qpos0 + (uint32_t)G > cur.j0;
        if (edge) {
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                const uint32_t qp = qp0 + (uint32_t)k;
                uint32_t allowed = (cur.have && qp >= qb32 && qp < qe32) ? ((1u << G) - 1u) : 0u;
                if constexpr (KIND == kSingle) {
                    if (qp >= cur.j0) {                                        // only leaves strictly right of the query
                        const uint32_t lo = qp - cur.j0 + 1u;
                        allowed &= lo >= (uint32_t)G ? 0u : ~((1u << lo) - 1u);
                    }
                }
                hits[k] &= allowed;
            }
        }
        uint32_t all = 0;
#pragma unroll
        for (int k = 0; k < QPL; ++k) all |= hits[k] << (G * k);
        append(all, qp0, cur.j0);
    };
    const uint32_t chunk = SLOTS * pyr_chunk_steps(count, SLOTS);        // see pyr_refine_kernel
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        const unsigned long long base64 = (unsigned long long)c * chunk;
        if (base64 >= count) break;
        const uint32_t base = (uint32_t)base64;
        const uint32_t end = count - base > chunk ? base + chunk : count;
        if constexpr (!kTma) asm volatile("cp.async.wait_all;" ::: "memory");   // the previous chunk's look-ahead copy is done
        __syncwarp();
        // the chunk's list entries into L1 (one 128-byte line = 16 entries per lane)
        if (base + 16u * lane < end) asm volatile("prefetch.global.L1 [%0];" ::"l"(in.data + base + 16u * lane));
        uint32_t p = base + slot;
        uint2 e1 = make_uint2(0u, 0u), e2 = make_uint2(0u, 0u);
        if (p < end) e1 = pyr_stream_load(in.data + p);
        if (p + SLOTS < end) e2 = pyr_stream_load(in.data + p + SLOTS);
        auto next_entry = [&]() {
            uint2 e = make_uint2(0u, 0u);
            if (p + 2 * SLOTS < end) e = pyr_stream_load(in.data + p + 2 * SLOTS);
            return e;
        };
        // optional (-DIBVH_PYR_AHEAD=k): volumes of the step k steps away towards L2. Measured: no gain at k = 4 or 8
        // (1.241 / 1.240 ms against 1.248 ms), so it is off; the L1 prefetch of the entries above is worth 0.03 ms.
        constexpr uint32_t kAhead = IBVH_PYR_AHEAD;
        auto l2_ahead = [&]() {
            if (kAhead && i == 0 && p + kAhead * SLOTS < end) {
                const uint2 e = pyr_stream_load(in.data + p + kAhead * SLOTS);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pq + (e.x << kPyrLeafLog)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pt + (e.y << kPyrLeafLog)));
            }
        };
        Stage sa = fetch(e1, p < end, 0), sb;
        for (uint32_t p0 = base; p0 < end;) {
            sb = fetch(e2, p + SLOTS < end, 1);
            e2 = next_entry();
            l2_ahead();
            process(sa, 0);
            p0 += SLOTS; p += SLOTS;
            if (p0 >= end) break;
            sa = fetch(e2, p + SLOTS < end, 0);
            e2 = next_entry();
            l2_ahead();
            process(sb, 1);
            p0 += SLOTS; p += SLOTS;
        }
    }
    if constexpr (!kTma) asm volatile("cp.async.wait_all;" ::: "memory");
    if (nbuf) flush(nbuf);
    if constexpr (MODE == kCount && PMODE == 0) {
        if (lane == 0 && ncount) atomicAdd(total, ncount);
    }
}

// Ordered protocol, single tile pass: stash entry (query, target position, target index) -> the query's segment.
template <class I>
__global__ void __launch_bounds__(256) pyr_scatter_kernel(const uint4* __restrict__ stash, const unsigned long long* __restrict__ total, int64_t cap,
                                                         const I* __restrict__ counts, unsigned int* cursors, IndexPair<I>* contacts) {
    unsigned long long n = *total;
    if ((long long)n > cap) n = (unsigned long long)cap;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#if IBVH_STREAM_HINTS
        const uint4 e = __ldcs(stash + i);
#else
        const uint4 e = stash[i];
#endif
        const int64_t seg = e.x == 0 ? 0 : (int64_t)counts[e.x - 1];
        const unsigned int rr = atomicAdd(&cursors[e.x], 1u);
        const long long li = (long long)((unsigned long long)e.z | ((unsigned long long)e.w << 32));
        contacts[seg + rr] = IndexPair<I>{(I)li, (I)e.y};
    }
}

// Ordered protocol, last step: each query's segment holds (target index, target position) in arrival order; sort it by
// position (ascending target position == the reference's DFS order) and convert to the reported index pair.
// Three tiers by segment length m: <= 8 in registers; <= kFixupInsertion by insertion sort in place (a few hundred moves
// at most); longer segments go on two lists — up to kFixupWarp entries one WARP per segment, beyond that one BLOCK per
// segment — and are sorted in place by a bitonic network whose every compare-exchange is ascending (partner i ^ (k - 1)
// in the first step of a merge, i ^ j afterwards), so that any length works without padding: an out-of-range partner is
// a virtual +inf and never swaps. Round 1 sorted long segments by insertion: a leaf touching 10^4 others cost 10^8 moves
// in one thread.
constexpr int kFixupInsertion = 32;
constexpr int kFixupWarp = 1024;

template <class I, class SYNC>
IBVH_D void fixup_bitonic(IndexPair<I>* seg, int64_t m, int tid, int nthreads, SYNC&& sync) {
    for (int64_t k = 2; (k >> 1) < m; k <<= 1) {
        for (int64_t i = tid; i < m; i += nthreads) {                     // first step of the merge: mirror partner
            const int64_t l = i ^ (k - 1);
            if (l > i && l < m) { const IndexPair<I> x = seg[i], y = seg[l]; if (x.b > y.b) { seg[i] = y; seg[l] = x; } }
        }
        sync();
        for (int64_t j = k >> 2; j > 0; j >>= 1) {
            for (int64_t i = tid; i < m; i += nthreads) {
                const int64_t l = i ^ j;
                if (l > i && l < m) { const IndexPair<I> x = seg[i], y = seg[l]; if (x.b > y.b) { seg[i] = y; seg[l] = x; } }
            }
            sync();
        }
    }
}

template <int KIND, class LQ, class LT, class I>
__global__ void __launch_bounds__(256) pyr_fixup_kernel(const LQ* __restrict__ qleaves, const I* __restrict__ qidx_arr, int64_t q_begin,
                                                       int64_t q_count, int flip, const I* __restrict__ counts, IndexPair<I>* contacts,
                                                       uint32_t* long_lists, uint32_t* long_counts, uint32_t long_cap) {
    // A block owns 256 consecutive queries, whose segments are ONE contiguous range of the list: the range travels
    // global -> shared -> global with coalesced accesses and every thread sorts its segment in shared memory (round 2 had
    // each thread read and write its own 8-byte entries in global memory: 0.51 ms for 2 x 318 MB at 10 M leaves). A range
    // that does not fit (a dense cluster) is sorted in place in global memory as before.
    constexpr int CAP = 24576 / (int)sizeof(IndexPair<I>);
    __shared__ IndexPair<I> s_seg[CAP];
    const int64_t q0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t qlast = (q0 + blockDim.x < q_count ? q0 + blockDim.x : q_count) - 1;
    const int64_t B = q0 == 0 ? 0 : (int64_t)counts[q0 - 1];
    const int64_t M = (int64_t)counts[qlast] - B;
    const bool staged = M <= CAP;                                        // block-uniform
    if (staged) {
        for (int64_t k = threadIdx.x; k < M; k += blockDim.x) s_seg[k] = contacts[B + k];
        __syncthreads();
    }
    const int64_t qi = q0 + threadIdx.x;
    const int64_t b = qi >= q_count ? 0 : (qi == 0 ? 0 : (int64_t)counts[qi - 1]);
    const int64_t e = qi >= q_count ? 0 : (int64_t)counts[qi];
    if (e > b) {
        const bool positions = (flip & 2) != 0;
        const int f = flip & 1;
        const I qidx = positions ? (I)(q_begin + qi + 1) : (qidx_arr ? __ldg(qidx_arr + q_begin + qi) : (I)qleaves[q_begin + qi].index);
        auto report = [&](I li) -> IndexPair<I> {
            I ea, eb;
            if constexpr (KIND == kSingle) { if (qidx > li) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
            else { if (f) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
            return IndexPair<I>{ea, eb};
        };
        IndexPair<I>* seg = staged ? s_seg + (b - B) : contacts + b;
        // entries are (target index, target position): sort by position — in registers for the usual handful of hits
        constexpr int kReg = 8;
        const int64_t m = e - b;
        bool queued = false;
        if (m <= kReg) {
            IndexPair<I> v[kReg];
#pragma unroll
            for (int x = 0; x < kReg; ++x) v[x] = x < m ? seg[x] : IndexPair<I>{I(0), I(0)};
#pragma unroll
            for (int x = 1; x < kReg; ++x) {
#pragma unroll
                for (int y = x; y > 0; --y) {
                    const bool sw = y < m && v[y - 1].b > v[y].b;       // slots >= m never move
                    const IndexPair<I> lo = sw ? v[y] : v[y - 1], hi = sw ? v[y - 1] : v[y];
                    v[y - 1] = lo; v[y] = hi;
                }
            }
#pragma unroll
            for (int x = 0; x < kReg; ++x) if (x < m) seg[x] = report(v[x].a);
        } else {
            if (m > kFixupInsertion) {
                // long segment: queue it for the cooperative kernels that follow (list 0: one warp each, list 1: one block each);
                // they work on global memory, where a staged block writes the segment back unchanged
                const int which = m > kFixupWarp ? 1 : 0;
                const uint32_t slot = atomicAdd(&long_counts[which], 1u);
                if (slot < long_cap) { long_lists[(size_t)which * long_cap + slot] = (uint32_t)qi; queued = true; }
                // (slot >= long_cap cannot happen: long_cap covers every possible long segment; the insertion sort below would do)
            }
            if (!queued) {
                for (int64_t x = 1; x < m; ++x) {
                    const IndexPair<I> v = seg[x];
                    int64_t y = x - 1;
                    while (y >= 0 && seg[y].b > v.b) { seg[y + 1] = seg[y]; --y; }
                    seg[y + 1] = v;
                }
                for (int64_t x = 0; x < m; ++x) seg[x] = report(seg[x].a);
            }
        }
    }
    if (staged) {
        __syncthreads();
        for (int64_t k = threadIdx.x; k < M; k += blockDim.x) contacts[B + k] = s_seg[k];
    }
}

// long segments: bitonic sort in place by target position, then the reported pairs. WARP_EACH: one warp per listed
// query (grid-stride over warps), else one block per listed query.
template <int KIND, bool WARP_EACH, class LQ, class I>
__global__ void __launch_bounds__(256) pyr_fixup_long_kernel(const LQ* __restrict__ qleaves, int64_t q_begin, int flip, const I* __restrict__ counts,
                                                            IndexPair<I>* contacts, const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count,
                                                            uint32_t long_cap) {
    const bool positions = (flip & 2) != 0;
    flip &= 1;
    uint32_t nlist = *list_count;
    if (nlist > long_cap) nlist = long_cap;
    const int lane = threadIdx.x & 31;
    const uint32_t unit = WARP_EACH ? (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : blockIdx.x;
    const uint32_t units = WARP_EACH ? gridDim.x * (blockDim.x >> 5) : gridDim.x;
    for (uint32_t s = unit; s < nlist; s += units) {
        const int64_t qi = (int64_t)list[s];
        const int64_t b = qi == 0 ? 0 : (int64_t)counts[qi - 1];
        const int64_t m = (int64_t)counts[qi] - b;
        IndexPair<I>* seg = contacts + b;
        const I qidx = positions ? (I)(q_begin + qi + 1) : (I)qleaves[q_begin + qi].index;
        const int tid = WARP_EACH ? lane : (int)threadIdx.x;
        const int nth = WARP_EACH ? 32 : (int)blockDim.x;
        if constexpr (WARP_EACH) fixup_bitonic<I>(seg, m, tid, nth, [] { __syncwarp(); });
        else fixup_bitonic<I>(seg, m, tid, nth, [] { __syncthreads(); });
        for (int64_t x = tid; x < m; x += nth) {
            const I li = seg[x].a;
            I ea, eb;
            if constexpr (KIND == kSingle) { if (qidx > li) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
            else { if (flip) { ea = li; eb = qidx; } else { ea = qidx; eb = li; } }
            seg[x] = IndexPair<I>{ea, eb};
        }
        if constexpr (WARP_EACH) __syncwarp(); else __syncthreads();
    }
}

}  // namespace ibvh
