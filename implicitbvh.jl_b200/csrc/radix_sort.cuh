// radix_sort.cuh — stable LSD radix sort of (Morton key, uint32 permutation) pairs, "onesweep"
// style: the per-pass digit histograms are produced up front (by the Morton encode kernel), and
// every pass is ONE kernel that ranks a tile in shared memory, publishes its per-digit counts and
// resolves its global offsets with a decoupled look-back over the preceding tiles.
//
// Replaces AK.sort!(leaves, by = bv -> bv.morton) (src/build.jl:248-253), which moves the whole
// 24/32-byte structs through a comparison sort. Here only 4/8-byte keys + a 4-byte permutation
// move; the structs are gathered once afterwards (aggregate.cuh). Ties keep input order (stable),
// the tie rule adopted in SURVEY.md §8c.
#pragma once
#include "common.cuh"
#include "morton.cuh"

namespace ibvh {

constexpr int kSortThreads = 256;          // == kRadixBins: thread d owns digit d in the scans
constexpr int kSortWarps = kSortThreads / 32;
// keys per thread: 16 for 2/4-byte keys, 12 for 8-byte keys (keeps the tile under 48 KB of static smem)
template <class K> constexpr int sort_items() { return sizeof(K) == 8 ? 12 : 16; }
template <class K> constexpr int sort_tile() { return kSortThreads * sort_items<K>(); }

constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kFlagAgg = 1u << 30;    // tile-local count published
constexpr uint32_t kFlagIncl = 2u << 30;   // inclusive prefix published
constexpr uint32_t kValMask = ~kFlagMask;

// exclusive scan of each pass's 256-bin histogram, in place. grid = passes, block = 256.
static __global__ void __launch_bounds__(256) scan_hist_kernel(uint32_t* hist) {
    __shared__ uint32_t wsum[8];
    uint32_t* h = hist + blockIdx.x * kRadixBins;
    uint32_t v = h[threadIdx.x];
    uint32_t incl = v;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int j = 0; j < w; ++j) base += wsum[j];
    h[threadIdx.x] = base + incl - v;
}

template <class K> IBVH_D uint32_t digit_of(K key, int shift) { return (uint32_t)(key >> shift) & (kRadixBins - 1); }

// One radix pass. keys_in/vals_in -> keys_out/vals_out. vals_in == nullptr: values are the global
// item index (first pass: the permutation starts as iota and need not be read).
// hist_excl: this pass's exclusive-scanned global histogram. lookback: [tiles][256] zero-initialised.
#ifndef IBVH_SORT_MINB
#define IBVH_SORT_MINB 4
#endif
template <class K>
__global__ void __launch_bounds__(kSortThreads, IBVH_SORT_MINB) onesweep_kernel(const K* __restrict__ keys_in, K* __restrict__ keys_out,
                                                               const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                                                               int64_t n, const uint32_t* __restrict__ hist_excl,
                                                               volatile uint32_t* lookback, uint32_t* ticket, int shift) {
    constexpr int kSortItems = sort_items<K>();
    constexpr int kSortTile = sort_tile<K>();
    __shared__ K skeys[kSortTile];
    __shared__ uint32_t svals[kSortTile];
    __shared__ uint32_t whist[kSortWarps][kRadixBins];   // per-warp digit counts -> exclusive offsets over warps
    __shared__ uint32_t tile_start[kRadixBins];          // exclusive scan of the tile's digit counts
    __shared__ int64_t gofs[kRadixBins];                 // global position = gofs[d] + position in the tile-sorted order
    __shared__ uint32_t s_tile;
    __shared__ uint32_t wsum[kSortWarps];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < kSortWarps * kRadixBins; i += kSortThreads) (&whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * kSortTile;
    const int tile_n = (int)min((int64_t)kSortTile, n - tile_base);

    // ---- load (warp-striped) and rank within the warp ------------------------------------------
    K key[kSortItems];
    uint32_t val[kSortItems];
    uint32_t rank[kSortItems];
    const int wbase = w * 32 * kSortItems;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        int li = wbase + r * 32 + lane;
        bool valid = li < tile_n;
        key[r] = valid ? keys_in[tile_base + li] : K(0);
        val[r] = valid ? (vals_in ? vals_in[tile_base + li] : (uint32_t)(tile_base + li)) : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        int li = wbase + r * 32 + lane;
        bool valid = li < tile_n;
        unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            uint32_t d = digit_of(key[r], shift);
            // peers = lanes holding the same digit. Eight ballots instead of match.any: MATCH runs on the ADU pipe,
            // which this kernel saturated (72 % ADU, 18 % issue slots in the first ncu capture).
            unsigned peers = vmask;
#pragma unroll
            for (int b = 0; b < kRadixBits; ++b) {
                const bool bit = (d >> b) & 1u;
                const unsigned bal = __ballot_sync(vmask, bit);
                peers &= bit ? bal : ~bal;
            }
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) { old = whist[w][d]; whist[w][d] = old + __popc(peers); }
            old = __shfl_sync(peers, old, leader);
            rank[r] = old + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit: exclusive offsets over warps, tile count, publish, tile-level scan -----------
    uint32_t tcount = 0;
    {
        const int d = tid;
#pragma unroll
        for (int j = 0; j < kSortWarps; ++j) { uint32_t c = whist[j][d]; whist[j][d] = tcount; tcount += c; }
        lookback[(size_t)tile * kRadixBins + d] = (tile == 0 ? kFlagIncl : kFlagAgg) | tcount;
        uint32_t incl = tcount;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int j = 0; j < w; ++j) base += wsum[j];
        tile_start[d] = base + incl - tcount;
    }
    __syncthreads();

    // ---- reorder the tile in shared memory ----------------------------------------------------------
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        int li = wbase + r * 32 + lane;
        if (li < tile_n) {
            uint32_t d = digit_of(key[r], shift);
            uint32_t pos = tile_start[d] + whist[w][d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = val[r];
        }
    }

    // ---- decoupled look-back (thread d resolves digit d) ----------------------------------------------
    {
        const int d = tid;
        uint32_t excl = 0;
        if (tile > 0) {
            // Walk back over the predecessors' (flag | count) words, kBatch independent loads in flight per round:
            // with ~3 resident tiles per SM the chain to the nearest published inclusive prefix is hundreds of
            // tiles long, and one dependent L2 round trip per tile was what bounded the whole pass.
            constexpr int kBatch = 8;
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t v[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k) v[k] = (t - k >= 0) ? lookback[(size_t)(t - k) * kRadixBins + d] : uint32_t(2u << 30);
                int used = 0;
#pragma unroll
                for (int k = 0; k < kBatch; ++k) {
                    if (!done && used == k) {
                        const uint32_t f = v[k] & kFlagMask;
                        if (f != 0) {                       // published: consume it
                            excl += v[k] & kValMask;
                            used = k + 1;
                            if (f == kFlagIncl) done = true;
                        }
                    }
                }
                t -= used;                                  // an unpublished predecessor is simply re-read
            }
            lookback[(size_t)tile * kRadixBins + d] = kFlagIncl | (excl + tcount);
        }
        gofs[d] = (int64_t)hist_excl[d] + (int64_t)excl - (int64_t)tile_start[d];
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive positions with one digit go to consecutive addresses ----------
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int i = k * kSortThreads + tid;
        if (i < tile_n) {
            K kk = skeys[i];
            int64_t dst = gofs[digit_of(kk, shift)] + i;
            keys_out[dst] = kk;
            vals_out[dst] = svals[i];
        }
    }
}

// keys[i] = leaves[i].morton  (stand-alone ibvh_sort_leaves entry) + histograms
template <class L>
__global__ void __launch_bounds__(256) extract_keys_kernel(const L* __restrict__ leaves, int64_t n, typename L::mor_t* __restrict__ keys,
                                                          L* __restrict__ copy_out, uint32_t* __restrict__ hist) {
    using M = typename L::mor_t;
    constexpr int P = radix_passes<M>();
    __shared__ uint32_t sh[P][kRadixBins];
    for (int i = threadIdx.x; i < P * kRadixBins; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Words<L> wv = load_words(leaves + i);
        M m = words_morton<L>(wv);
        keys[i] = m;
        store_words(copy_out + i, wv);
#pragma unroll
        for (int p = 0; p < P; ++p) atomicAdd(&sh[p][(uint32_t)(m >> (p * kRadixBits)) & (kRadixBins - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * kRadixBins; i += blockDim.x) {
        uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

}  // namespace ibvh
