// radix_sort.cuh — stable LSD radix sort of (Morton key, uint32 permutation) pairs, "onesweep"
// style: the per-pass digit histograms are produced up front (by the Morton encode kernel), and
// every pass is ONE kernel that ranks a tile in shared memory, publishes its per-digit counts and
// resolves its global offsets with a decoupled look-back over the preceding tiles.
//
// Replaces AK.sort!(leaves, by = bv -> bv.morton) (src/build.jl:248-253), which moves the whole
// 24/32-byte structs through a comparison sort. Here only 4/8-byte keys + a 4-byte permutation
// move; the structs are gathered once afterwards (aggregate.cuh). Ties keep input order (stable),
// the tie rule adopted in SURVEY.md §8c.
//
// Round-2 kernel (the round-1 one ran at 21 % of HBM peak: 183 instructions per key, 80 registers,
// issue-bound on the ranking chain). What changed and why:
//   * every ballot runs under the FULL warp mask. Round 1 ranked under `if (valid)` with a partial
//     mask, which nvcc turned into WARPSYNC / ENDCOLLECTIVE / BSSY brackets around each of the 8
//     ballots (268 VOTE + 156 WARPSYNC in the SASS of a 16-key thread). The last, partial tile is
//     padded with a sentinel key (all key bits set): sentinels carry the largest digit and sit at
//     the end of the tile order, so a stable ranking places them past every real key and one
//     subtraction fixes the tile's count of that digit — no validity masks anywhere in the loop.
//   * the top pass ballots only the bits the key still has (30-bit codes: 8 + 8 + 8 + 6).
//   * all peers read the warp's digit counter (one broadcast LDS) and the highest peer lane writes
//     it back: no leader election + shuffle round trip.
//   * 16-bit per-warp digit counters (a warp holds <= 512 keys): half the shared memory, so more
//     resident tiles; keys per thread and threads per tile are template parameters (sort_bench.cu
//     sweeps them on the GPU) — 512 threads x 8 keys keeps ~40 registers per thread.
//   * global positions are computed modulo 2^32 (n <= 2^31 since levels <= 32), and the look-back
//     word is 64-bit when n >= 2^30, which lifts round 1's n < 2^30 limit.
#pragma once
#include "common.cuh"
#include "morton.cuh"

namespace ibvh {

// ---- tile shape ----------------------------------------------------------------------------------------
#ifndef IBVH_SORT_THREADS
#define IBVH_SORT_THREADS 512
#endif
#ifndef IBVH_SORT_ITEMS
#define IBVH_SORT_ITEMS 8
#endif
#ifndef IBVH_SORT_MINB
#define IBVH_SORT_MINB 3
#endif
constexpr int kSortThreads = IBVH_SORT_THREADS;
template <class K> constexpr int sort_items() { return IBVH_SORT_ITEMS; }
template <class K> constexpr int sort_tile() { return kSortThreads * sort_items<K>(); }

// look-back words: 2 flag bits on top of a count / prefix
template <class LB> struct LookbackWord;
template <> struct LookbackWord<uint32_t> {
    static constexpr uint32_t kAgg = 1u << 30, kIncl = 2u << 30, kFlags = 3u << 30;
};
template <> struct LookbackWord<unsigned long long> {
    static constexpr unsigned long long kAgg = 1ull << 62, kIncl = 2ull << 62, kFlags = 3ull << 62;
};

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

// One bit of the warp-wide digit match: keep the lanes whose bit B of the digit equals mine. Written in PTX because
// nvcc derived TWO predicates per bit from the C++ form (shift + and + setp for the ballot, and + setp + sel for the
// mask: 5.5 instructions per bit); this is and+setp (one LOP3), VOTE, a predicated NOT and an AND: 4 per bit.
template <int B> IBVH_D void match_bit(uint32_t d, uint32_t& peers) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t, b;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
        "@!p not.b32 b, b;\n\t"
        "and.b32 %0, %0, b;\n\t}"
        : "+r"(peers) : "r"(d), "n"(1u << B));
}
template <int BITS> IBVH_D uint32_t match_digit(uint32_t d) {
    uint32_t peers = 0xffffffffu;
    match_bit<0>(d, peers);
    if constexpr (BITS > 1) match_bit<1>(d, peers);
    if constexpr (BITS > 2) match_bit<2>(d, peers);
    if constexpr (BITS > 3) match_bit<3>(d, peers);
    if constexpr (BITS > 4) match_bit<4>(d, peers);
    if constexpr (BITS > 5) match_bit<5>(d, peers);
    if constexpr (BITS > 6) match_bit<6>(d, peers);
    if constexpr (BITS > 7) match_bit<7>(d, peers);
    return peers;
}

template <class K, int THREADS, int ITEMS> constexpr size_t onesweep_smem_bytes() {
    return (size_t)THREADS * ITEMS * (sizeof(K) + 4) + (size_t)(THREADS / 32) * kRadixBins * 2 + 2 * kRadixBins * 4;
}

// One radix pass. keys_in/vals_in -> keys_out/vals_out. vals_in == nullptr: values are the global
// item index (first pass: the permutation starts as iota and need not be read).
// hist_excl: this pass's exclusive-scanned global histogram. lookback: [tiles][256] zero-initialised.
// BITS: digit bits that can differ in this pass (8, or what is left of the key in the top pass).
// sentinel: a key with every key bit set (pads the last tile).
template <class K, class LB, int THREADS, int ITEMS, int BITS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) onesweep_kernel(const K* __restrict__ keys_in, K* __restrict__ keys_out,
                                                                const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                                                                int64_t n, const uint32_t* __restrict__ hist_excl,
                                                                volatile LB* lookback, uint32_t* ticket, int shift, K sentinel) {
    static_assert(THREADS >= kRadixBins && THREADS % 32 == 0, "one thread per digit in the scans");
    static_assert(32 * ITEMS <= 65535 && THREADS * ITEMS <= 65535, "16-bit per-warp counters / offsets");
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    constexpr LB kAgg = LookbackWord<LB>::kAgg, kIncl = LookbackWord<LB>::kIncl, kFlags = LookbackWord<LB>::kFlags;
    extern __shared__ __align__(16) unsigned char onesweep_smem[];
    K* skeys = reinterpret_cast<K*>(onesweep_smem);                                   // TILE keys, tile-sorted
    uint32_t* svals = reinterpret_cast<uint32_t*>(skeys + TILE);                      // TILE values
    uint16_t* whist = reinterpret_cast<uint16_t*>(svals + TILE);                      // [WARPS][256] counts -> exclusive offsets over warps
    uint32_t* tile_start = reinterpret_cast<uint32_t*>(whist + WARPS * kRadixBins);   // exclusive scan of the tile's digit counts
    uint32_t* gofs = tile_start + kRadixBins;                                         // global position - position in the tile order (mod 2^32)
    __shared__ uint32_t s_tile;
    __shared__ uint32_t wsum[8];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * kRadixBins / 2; i += THREADS) reinterpret_cast<uint32_t*>(whist)[i] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int tile_n = (int)min((int64_t)TILE, n - tile_base);

    // ---- load, warp-striped: item r of lane l is tile position w * 32 * ITEMS + r * 32 + l -----------
    K key[ITEMS];
    const int wbase = w * 32 * ITEMS + lane;
    const bool full = tile_n == TILE;
    if (full) {
        const K* kp = keys_in + tile_base + wbase;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) key[r] = kp[r * 32];
    } else {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) key[r] = (wbase + r * 32 < tile_n) ? keys_in[tile_base + wbase + r * 32] : sentinel;
    }

    // ---- rank within the warp: peers = lanes holding the same digit (BITS full-mask ballots) -----------
    uint32_t rank[ITEMS];
    {
        uint16_t* wh = whist + w * kRadixBins;
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const uint32_t d = digit_of(key[r], shift);
            const uint32_t peers = match_digit<BITS>(d);
            const uint32_t old = wh[d];                                   // every peer reads the same counter (broadcast)
            rank[r] = old + __popc(peers & lt);
            if ((peers >> lane) == 1u) wh[d] = (uint16_t)(old + __popc(peers));   // the highest peer lane writes it back
            __syncwarp();
        }
    }
    // the values are only needed for the reorder: loaded here (not with the keys) they do not occupy registers during
    // the ranking loop, and their latency is covered by the digit scans below
    uint32_t val[ITEMS];
    if (vals_in) {
        if (full) {
            const uint32_t* vp = vals_in + tile_base + wbase;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) val[r] = vp[r * 32];
        } else {
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) val[r] = (wbase + r * 32 < tile_n) ? vals_in[tile_base + wbase + r * 32] : 0u;
        }
    } else {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) val[r] = (uint32_t)(tile_base + wbase + r * 32);
    }
    __syncthreads();

    // ---- per digit (thread d): exclusive offsets over warps, tile count, publish, tile-level scan ---------
    uint32_t tcount = 0, incl = 0;
    if (tid < kRadixBins) {
        const int d = tid;
#pragma unroll
        for (int j = 0; j < WARPS; ++j) { const uint32_t c = whist[j * kRadixBins + d]; whist[j * kRadixBins + d] = (uint16_t)tcount; tcount += c; }
        // sentinels of a partial tile were ranked like keys: they hold the largest digit and the last tile positions
        if (tile_n < TILE && d == (int)digit_of(sentinel, shift)) tcount -= (uint32_t)(TILE - tile_n);
        lookback[(size_t)tile * kRadixBins + d] = (tile == 0 ? kIncl : kAgg) | (LB)tcount;
        incl = tcount;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
        if (lane == 31) wsum[w] = incl;
    }
    __syncthreads();
    if (tid < kRadixBins) {
        uint32_t base = 0;
        for (int j = 0; j < w; ++j) base += wsum[j];
        tile_start[tid] = base + incl - tcount;
    }
    __syncthreads();

    // ---- reorder the tile in shared memory -------------------------------------------------------------
    {
        const uint16_t* wh = whist + w * kRadixBins;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const uint32_t d = digit_of(key[r], shift);
            const uint32_t pos = tile_start[d] + wh[d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = val[r];
        }
    }

    // ---- decoupled look-back (thread d resolves digit d) -----------------------------------------------
    if (tid < kRadixBins) {
        const int d = tid;
        LB excl = 0;
        if (tile > 0) {
            // Walk back over the predecessors' (flag | count) words, kBatch independent loads in flight per round:
            // the chain to the nearest published inclusive prefix is as long as the number of tiles in flight, and
            // one dependent L2 round trip per tile was what bounded the whole pass.
            constexpr int kBatch = 8;
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                LB v[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k) v[k] = (t - k >= 0) ? lookback[(size_t)(t - k) * kRadixBins + d] : kIncl;
                int used = 0;
#pragma unroll
                for (int k = 0; k < kBatch; ++k) {
                    if (!done && used == k) {
                        const LB f = v[k] & kFlags;
                        if (f != 0) {                       // published: consume it
                            excl += v[k] & ~kFlags;
                            used = k + 1;
                            if (f == kIncl) done = true;
                        }
                    }
                }
                t -= used;                                  // an unpublished predecessor is simply re-read
            }
            lookback[(size_t)tile * kRadixBins + d] = kIncl | (excl + (LB)tcount);
        }
        gofs[d] = hist_excl[d] + (uint32_t)excl - tile_start[d];
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive positions with one digit go to consecutive addresses ----------
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int i = k * THREADS + tid;
        if (i < tile_n) {
            const K kk = skeys[i];
            const uint32_t dst = gofs[digit_of(kk, shift)] + (uint32_t)i;
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

// ---- host driver ---------------------------------------------------------------------------------------
// keysA holds the input keys; hist the raw per-pass histograms; lookback is zeroed and holds
// passes * tiles * 256 words of LB. On return the sorted keys / the permutation are in *keys_out / *vals_out.
// LAUNCHER(name) brackets one launch (profiling scope); returns the first CUDA launch error.
// a key with every key bit set (15 / 30 / 63 bits): pads the last tile
template <class K> constexpr K sort_sentinel() { return (K)((1ull << MortonTraits<K>::key_bits) - 1ull); }

template <class K, class LB, int THREADS, int ITEMS, int MINB, int BITS>
inline cudaError_t launch_onesweep_pass(const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, int64_t n, const uint32_t* hist_excl,
                                        LB* lookback, uint32_t* ticket, int shift, cudaStream_t st) {
    auto kern = onesweep_kernel<K, LB, THREADS, ITEMS, BITS, MINB>;
    constexpr size_t smem = onesweep_smem_bytes<K, THREADS, ITEMS>();
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int64_t tiles = (n + (int64_t)THREADS * ITEMS - 1) / ((int64_t)THREADS * ITEMS);
    kern<<<(unsigned)tiles, THREADS, smem, st>>>(kin, kout, vin, vout, n, hist_excl, lookback, ticket, shift, sort_sentinel<K>());
    return cudaGetLastError();
}

template <class K, class LB, int THREADS, int ITEMS, int MINB, class SCOPE>
inline cudaError_t sort_pairs_impl(K* keysA, K* keysB, uint32_t* valsA, uint32_t* valsB, int64_t n, uint32_t* hist, void* lookback_raw,
                                   uint32_t* tickets, cudaStream_t st, K** keys_out, uint32_t** vals_out, SCOPE&& scope) {
    constexpr int P = radix_passes<K>();
    constexpr int kTopBits = MortonTraits<K>::key_bits - kRadixBits * (P - 1);
    const int64_t tiles = (n + (int64_t)THREADS * ITEMS - 1) / ((int64_t)THREADS * ITEMS);
    LB* lookback = (LB*)lookback_raw;
    {
        auto s = scope("scan_hist_kernel");
        scan_hist_kernel<<<P, 256, 0, st>>>(hist);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    K* kin = keysA; K* kout = keysB;
    uint32_t* vin = nullptr; uint32_t* vout = valsB; uint32_t* vother = valsA;
    for (int p = 0; p < P; ++p) {
        {
            auto s = scope("onesweep_kernel");
            if (p == P - 1 && kTopBits < kRadixBits)
                e = launch_onesweep_pass<K, LB, THREADS, ITEMS, MINB, kTopBits>(kin, kout, vin, vout, n, hist + p * kRadixBins, lookback + (size_t)p * tiles * kRadixBins, tickets + p, p * kRadixBits, st);
            else
                e = launch_onesweep_pass<K, LB, THREADS, ITEMS, MINB, kRadixBits>(kin, kout, vin, vout, n, hist + p * kRadixBins, lookback + (size_t)p * tiles * kRadixBins, tickets + p, p * kRadixBits, st);
        }
        if (e != cudaSuccess) return e;
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* nv = vout; vout = vother; vother = nv; vin = nv;
    }
    *keys_out = kin;
    *vals_out = vin;
    return cudaSuccess;
}

}  // namespace ibvh
