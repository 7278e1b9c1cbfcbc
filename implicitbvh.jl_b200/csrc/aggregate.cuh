// aggregate.cuh — gather of the Morton-sorted leaves fused with the bottom-up bounding-volume merge.
// Replaces the tail of the BVH constructor: the struct movement of AK.sort! (src/build.jl:248-253)
// and aggregate_oibvh! / _aggregate_last_level_at! / _aggregate_level_at! (src/build.jl:366-523).
//
// B200 design: one CTA owns a tile of TILE consecutive *sorted* leaf positions. It gathers the tile's
// leaves through the sort permutation into shared memory, streams them back to the caller's leaf
// array with coalesced vector stores, and — because the implicit tree maps a tile of 2^k leaves onto
// a contiguous run of 2^(k-j) nodes on each of the k levels above — merges up to log2(TILE) levels
// entirely in shared memory, writing each level's run out once. The remaining top levels are walked
// by re-running the same tile scheme on the last written level (a handful of CTAs), so a 10 M-leaf
// tree needs 3 launches instead of the reference's 24 dependent per-level launches.
#pragma once
#include "common.cuh"

namespace ibvh {

// Source record of the gather: a wrapped leaf (in-place path: the copy streamed out by the encode
// kernel) or a raw volume (wrap path: index = original position + 1, build.jl:345-349).
template <class L, class SRC> struct GatherSrc;
template <class L> struct GatherSrc<L, L> {
    static IBVH_D Words<L> make(const L* src, uint32_t p, typename L::mor_t key, int = 8) {
        Words<L> wv = load_words(src + p);
        words_set_morton<L>(wv, key);
        return wv;
    }
};
template <class L> struct GatherSrc<L, typename L::vol_t> {
    static IBVH_D Words<L> make(const typename L::vol_t* src, uint32_t p, typename L::mor_t key, int vec = 4) {
        Words<L> wv = zero_words<L>();
        words_set_volume<L>(wv, load_volume(src + p, vec));
        words_set_index<L>(wv, (typename L::idx_t)(p + 1u));
        words_set_morton<L>(wv, key);
        return wv;
    }
};

// Sidecar outputs of the build (handle.cuh: Sidecar): every pointer may be null.
template <class N> struct SideLevels { N* lvl[34]; };
template <class V, class N> struct SideOut {
    Packed<V>* pt;                               // packed leaf volumes [n + 8] (the tail is NaN padding)
    N* lvl[34];                                  // lvl[l]: 64-byte aligned run receiving a copy of the nodes of tree level l
    UBox<typename N::value_type>* u[3];          // query-pyramid levels k = 2, 5, 8 (exact unions of the leaves' boxes), index = group
};

// Stand-alone gather: leaves[i] = source[perm[i]] with morton = keys_sorted[i]. Two leaves per thread, all index
// loads before the random loads. The merge then runs from the sorted leaves (gather_merge_kernel<GATHER=false>):
// one extra coalesced read of the leaves, but neither kernel waits on the other's latency inside a CTA.
template <class L, class SRC>
__global__ void __launch_bounds__(256) gather_kernel(const SRC* __restrict__ src, const uint32_t* __restrict__ perm,
                                                    const typename L::mor_t* __restrict__ keys_sorted, L* __restrict__ leaves, int64_t n, int vec,
                                                    Packed<typename L::vol_t>* __restrict__ packed, typename L::idx_t* __restrict__ idx_out) {
    constexpr int PER = 4;
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x) * PER + threadIdx.x;
    uint32_t pp[PER];
    typename L::mor_t kk[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int64_t i = base + (int64_t)u * blockDim.x;
        const bool ok = i < n;
        pp[u] = ok ? perm[i] : 0u;
        kk[u] = ok ? keys_sorted[i] : typename L::mor_t(0);
    }
    Words<L> wv[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int64_t i = base + (int64_t)u * blockDim.x;
        if (i < n) wv[u] = GatherSrc<L, SRC>::make(src, pp[u], kk[u], vec);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int64_t i = base + (int64_t)u * blockDim.x;
        if (i < n) {
            store_words(leaves + i, wv[u]);
            if (packed) {                        // sidecar: the volume alone as a 16-byte aligned record (zero padding)
                alignas(16) Packed<typename L::vol_t> r;
                memset(&r, 0, sizeof(r));
                r.v = words_volume<L>(wv[u]);
                store16(packed + i, r);
            }
            if (idx_out) idx_out[i] = words_index<L>(wv[u]);      // sidecar: the indices alone (4 / 8 bytes per leaf: L2-resident at 10 M)
        }
    }
    if (packed && blockIdx.x == 0 && threadIdx.x < 8) {      // NaN records past the end: no bounds masks on the target side
        alignas(16) Packed<typename L::vol_t> r;
        memset(&r, 0xFF, sizeof(r));
        store16(packed + n + threadIdx.x, r);
    }
}

struct LevelPlan {
    int32_t src_level;       // level whose nodes (or leaves, if == levels) are the tile input
    int32_t stop_level;      // lowest level (numerically) this launch produces: max(built_level, src_level - log2(TILE))
};

// Merge the levels above a tile whose input volumes sit in shared memory.
//   in_count: number of input slots of the tile (TILE); tile index t; input level `lvl_in`.
//   sbuf0/sbuf1: ping-pong node buffers of TILE/2 and TILE/4 entries.
template <class N, int TILE, int THREADS, class LOADPAIR>
IBVH_D void merge_tile_levels(N* nodes, const TreeInfo& ti, int lvl_in, int stop_level, int64_t tile, N* sbuf0, N* sbuf1, LOADPAIR first_level,
                              N* const* side_lvl = nullptr) {
    // first produced level: lvl_in - 1, from the tile input through `first_level(j, left_only)`
    int lvl = lvl_in - 1;
    int count = TILE / 2;
    N* cur = sbuf0;
    N* nxt = sbuf1;
    {
        const int64_t base = tile * count;                 // 0-based index within level `lvl`
        const int64_t nreal = ti.level_nreal[lvl];
        const int64_t nreal_child = ti.level_nreal[lvl + 1];
        for (int j = threadIdx.x; j < count; j += THREADS) {
            int64_t gi = base + j;
            if (gi < nreal) {
                bool right_virtual = (2 * gi + 1) >= nreal_child;
                N v = first_level(j, right_virtual);
                cur[j] = v;
                if (lvl >= stop_level) {
                    nodes[ti.level_start[lvl] + gi] = v;
                    if (side_lvl && side_lvl[lvl]) side_lvl[lvl][gi] = v;
                }
            }
        }
    }
    __syncthreads();
    while (lvl - 1 >= stop_level && count > 1) {
        lvl -= 1;
        count >>= 1;
        const int64_t base = tile * count;
        const int64_t nreal = ti.level_nreal[lvl];
        const int64_t nreal_child = ti.level_nreal[lvl + 1];
        for (int j = threadIdx.x; j < count; j += THREADS) {
            int64_t gi = base + j;
            if (gi < nreal) {
                // _aggregate_level_at!, build.jl:503-523: right child virtual -> copy the left child
                N v = ((2 * gi + 1) >= nreal_child) ? cur[2 * j] : merge(cur[2 * j], cur[2 * j + 1]);
                nxt[j] = v;
                nodes[ti.level_start[lvl] + gi] = v;
                if (side_lvl && side_lvl[lvl]) side_lvl[lvl][gi] = v;
            }
        }
        __syncthreads();
        N* t = cur; cur = nxt; nxt = t;
    }
}

// Gather + write-back + bottom levels. Dynamic shared memory: TILE leaves + TILE/2 + TILE/4 nodes.
// GATHER == false: leaves are already sorted in `leaves` (stand-alone ibvh_aggregate): no permutation,
// no write-back.
template <class L, class SRC, class N, int TILE, int THREADS, bool GATHER>
__global__ void __launch_bounds__(THREADS) gather_merge_kernel(const SRC* __restrict__ src, const uint32_t* __restrict__ perm,
                                                              const typename L::mor_t* __restrict__ keys_sorted,
                                                              L* leaves, N* nodes, TreeInfo ti, int stop_level, int vec,
                                                              SideOut<typename L::vol_t, N> so) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    L* sleaf = (L*)smem_raw;
    N* sbuf0 = (N*)(smem_raw + ((sizeof(L) * TILE + 15) & ~size_t(15)));
    N* sbuf1 = sbuf0 + TILE / 2;
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * TILE;
    const int tile_n = (int)min((int64_t)TILE, ti.n - base);

    if constexpr (GATHER) {
        // all permutation / key loads first, then all (random) source loads: TILE/THREADS independent
        // gathers in flight per thread instead of one dependent chain at a time
        constexpr int PER = TILE / THREADS;
        uint32_t pp[PER];
        typename L::mor_t kk[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            int j = threadIdx.x + u * THREADS;
            bool ok = j < tile_n;
            pp[u] = ok ? perm[base + j] : 0u;
            kk[u] = ok ? keys_sorted[base + j] : typename L::mor_t(0);
        }
        Words<L> wv[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            int j = threadIdx.x + u * THREADS;
            if (j < tile_n) wv[u] = GatherSrc<L, SRC>::make(src, pp[u], kk[u], vec);
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            int j = threadIdx.x + u * THREADS;
            if (j < tile_n) store_words(sleaf + j, wv[u]);
        }
        __syncthreads();
        // coalesced write-back of the tile (sizeof(L) is a multiple of 4; full tiles are 16-byte multiples)
        const size_t bytes = (size_t)tile_n * sizeof(L);
        unsigned char* dst = (unsigned char*)(leaves + base);
        if ((bytes & 15) == 0 && ((uintptr_t)dst & 15) == 0) {
            const uint4* s4 = (const uint4*)smem_raw;
            uint4* d4 = (uint4*)dst;
            for (size_t k = threadIdx.x; k < bytes / 16; k += THREADS) d4[k] = s4[k];
        } else {
            const uint32_t* s1 = (const uint32_t*)smem_raw;
            uint32_t* d1 = (uint32_t*)dst;
            for (size_t k = threadIdx.x; k < bytes / 4; k += THREADS) d1[k] = s1[k];
        }
    } else {
        for (int j = threadIdx.x; j < tile_n; j += THREADS) sleaf[j] = leaves[base + j];
        __syncthreads();
    }
    if (ti.levels < 2) return;
    // sidecar: the three finest levels of the query pyramid — EXACT unions of the leaves' own boxes (not tree nodes: a
    // leaf-parent box need not contain its leaves' boxes in floating point, merge.jl:62-68) for groups of 4, 32, 256 leaves
    if constexpr (std::is_same<N, BBox<typename N::value_type>>::value) {
        if (so.u[0]) {
            using T = typename N::value_type;
            static_assert(sizeof(UBox<T>) * (TILE / 4) <= sizeof(N) * (TILE / 2), "the pyramid boxes of a tile fit the first node buffer");
            UBox<T>* su = reinterpret_cast<UBox<T>*>(sbuf0);
            for (int g = threadIdx.x; g < TILE / 4; g += THREADS) {
                BBox<T> u = empty_box<T>();
#pragma unroll
                for (int m = 0; m < 4; ++m) if (4 * g + m < tile_n) u = merge(u, NodeOps<BBox<T>>::convert(sleaf[4 * g + m].volume));
                su[g].b = u;
                so.u[0][tile * (TILE / 4) + g].b = u;
            }
            __syncthreads();
            UBox<T>* su1 = su + TILE / 4;
            if (so.u[1]) {
                for (int g = threadIdx.x; g < TILE / 32; g += THREADS) {
                    BBox<T> u = empty_box<T>();
#pragma unroll
                    for (int m = 0; m < 8; ++m) u = merge(u, su[8 * g + m].b);
                    su1[g].b = u;
                    so.u[1][tile * (TILE / 32) + g].b = u;
                }
            }
            __syncthreads();
            if (so.u[2]) {
                for (int g = threadIdx.x; g < TILE / 256; g += THREADS) {
                    BBox<T> u = empty_box<T>();
#pragma unroll
                    for (int m = 0; m < 8; ++m) u = merge(u, su1[8 * g + m].b);
                    so.u[2][tile * (TILE / 256) + g].b = u;
                }
            }
            __syncthreads();
        }
    }
    // _aggregate_last_level_at!, build.jl:427-457
    merge_tile_levels<N, TILE, THREADS>(nodes, ti, ti.levels, stop_level, tile, sbuf0, sbuf1,
        [&](int j, bool right_virtual) -> N {
            return right_virtual ? NodeOps<N>::convert(sleaf[2 * j].volume)
                                 : NodeOps<N>::merge_leaves(sleaf[2 * j].volume, sleaf[2 * j + 1].volume);
        }, so.lvl);
}

// Upper levels: tile input = nodes of level `src_level` already in global memory.
template <class N, int TILE, int THREADS>
__global__ void __launch_bounds__(THREADS) merge_levels_kernel(N* nodes, TreeInfo ti, int src_level, int stop_level, SideLevels<N> side) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    N* sin = (N*)smem_raw;
    N* sbuf0 = sin + TILE;
    N* sbuf1 = sbuf0 + TILE / 2;
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * TILE;
    const int64_t nreal_in = ti.level_nreal[src_level];
    const int tile_n = (int)max((int64_t)0, min((int64_t)TILE, nreal_in - base));
    const N* in = nodes + ti.level_start[src_level] + base;
    for (int j = threadIdx.x; j < tile_n; j += THREADS) sin[j] = in[j];
    __syncthreads();
    merge_tile_levels<N, TILE, THREADS>(nodes, ti, src_level, stop_level, tile, sbuf0, sbuf1,
        [&](int j, bool right_virtual) -> N { return right_virtual ? sin[2 * j] : merge(sin[2 * j], sin[2 * j + 1]); }, side.lvl);
}

}  // namespace ibvh
