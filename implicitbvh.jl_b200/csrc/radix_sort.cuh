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
// Swept on a B200 with tools/sort_bench.cu (10 M pairs, 30-bit keys; gpurun_out/r2a..r2e, summary in profiles/):
//   threads x keys/thread (CTAs/SM)   256x16 (4)  512x8 (3)  384x12 (3)  512x12 (2)  512x16 (2)  1024x8 (1)
//   ms for the 4 passes                  0.263      0.263      0.259       0.256      0.236       0.315
// look-back words in flight per round (512x16): 1 -> 0.267, 2 -> 0.240, 3 -> 0.238, 4 -> 0.236, 8 -> 0.249, 16 -> 0.294 ms
// (the nearest published inclusive prefix is usually 1-3 tiles back: wider batches only add instructions);
// requesting them BEFORE the reorder: slower (0.277: the predecessors have not published yet, the words are re-read).
// 8-byte keys: 512x12 is best (0.623 ms for 8 passes at 10 M, 5.7 ms at 100 M).
#ifndef IBVH_SORT_THREADS
#define IBVH_SORT_THREADS 512
#endif
#ifndef IBVH_SORT_MINB
#define IBVH_SORT_MINB 2
#endif
#ifndef IBVH_SORT_PREFETCH
#define IBVH_SORT_PREFETCH 0      // request the nearest predecessors' look-back words before the reorder (measured: slower)
#endif
#ifndef IBVH_SORT_BATCH
#define IBVH_SORT_BATCH 4         // look-back words in flight per round
#endif
#ifndef IBVH_SORT_PAIRSTAGE
#define IBVH_SORT_PAIRSTAGE 1     // 4-byte keys: (key, value) staged in shared memory as one 8-byte record
#endif
#ifndef IBVH_SORT_PACKRANK
#define IBVH_SORT_PACKRANK 1      // two 16-bit ranks per register, pinned when computed
#endif
constexpr int kSortThreads = IBVH_SORT_THREADS;
#ifdef IBVH_SORT_ITEMS
template <class K> constexpr int sort_items() { return IBVH_SORT_ITEMS; }
#else
template <class K> constexpr int sort_items() { return sizeof(K) == 8 ? 12 : 16; }
#endif
template <class K> constexpr int sort_tile() { return kSortThreads * sort_items<K>(); }
// resident tiles per SM the register budget is set for
template <class K> constexpr int sort_minb() { return IBVH_SORT_MINB; }

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
    constexpr bool kPairStage = IBVH_SORT_PAIRSTAGE && sizeof(K) == 4;      // stage (key, value) as one 8-byte record
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
    for (int i = tid; i < WARPS * kRadixBins / 8; i += THREADS) reinterpret_cast<uint4*>(whist)[i] = make_uint4(0u, 0u, 0u, 0u);
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
    // Two 16-bit ranks per register, each pinned down as soon as it is known (the empty asm): nvcc otherwise keeps the
    // peer masks and counters of all ITEMS keys alive and finishes the ranks during the reorder, spilling to local memory.
#if IBVH_SORT_PACKRANK
    uint32_t rank2[(ITEMS + 1) / 2];
#else
    uint32_t rank1[ITEMS];
#endif
    {
        uint16_t* wh = whist + w * kRadixBins;
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const uint32_t d = digit_of(key[r], shift);
            const uint32_t peers = match_digit<BITS>(d);
            const uint32_t old = wh[d];                                   // every peer reads the same counter (broadcast)
            const uint32_t rk = old + __popc(peers & lt);
            if ((peers >> lane) == 1u) wh[d] = (uint16_t)(old + __popc(peers));   // the highest peer lane writes it back
#if IBVH_SORT_PACKRANK
            if (r & 1) rank2[r / 2] |= rk << 16; else rank2[r / 2] = rk;
            asm volatile("" : "+r"(rank2[r / 2]));
#else
            rank1[r] = rk;
#endif
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
    // Look-back, part 1: the words of the kBatch nearest predecessors are requested HERE, right after this tile's own
    // count is published — they travel while the scans, two barriers and the reorder run (ncu, round 2: with the
    // loads issued only after the reorder the walk was 16 % of the stall samples and the barrier behind it 15 %).
    constexpr int kBatch = IBVH_SORT_BATCH;
    uint32_t tcount = 0, incl = 0;
#if IBVH_SORT_PREFETCH
    LB pre[kBatch];
#endif
    if (tid < kRadixBins) {
        const int d = tid;
#pragma unroll
        for (int j = 0; j < WARPS; ++j) { const uint32_t c = whist[j * kRadixBins + d]; whist[j * kRadixBins + d] = (uint16_t)tcount; tcount += c; }
        // sentinels of a partial tile were ranked like keys: they hold the largest digit and the last tile positions
        if (tile_n < TILE && d == (int)digit_of(sentinel, shift)) tcount -= (uint32_t)(TILE - tile_n);
        lookback[(size_t)tile * kRadixBins + d] = (tile == 0 ? kIncl : kAgg) | (LB)tcount;
#if IBVH_SORT_PREFETCH
#pragma unroll
        for (int k = 0; k < kBatch; ++k) pre[k] = ((int64_t)tile - 1 - k >= 0) ? lookback[((size_t)tile - 1 - k) * kRadixBins + d] : kIncl;
#endif
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
#if IBVH_SORT_PACKRANK
            const uint32_t pos = tile_start[d] + wh[d] + ((r & 1) ? (rank2[r / 2] >> 16) : (rank2[r / 2] & 0xffffu));
#else
            const uint32_t pos = tile_start[d] + wh[d] + rank1[r];
#endif
            if constexpr (kPairStage) {
                reinterpret_cast<uint2*>(onesweep_smem)[pos] = make_uint2((uint32_t)key[r], val[r]);      // one 8-byte store per pair
            } else {
                skeys[pos] = key[r];
                svals[pos] = val[r];
            }
        }
    }

    // ---- decoupled look-back, part 2 (thread d resolves digit d) ---------------------------------------
    if (tid < kRadixBins) {
        const int d = tid;
        LB excl = 0;
        if (tile > 0) {
            // Walk back over the predecessors' (flag | count) words, kBatch independent loads in flight per round: the
            // chain to the nearest published inclusive prefix is as long as the number of tiles in flight. An entry is
            // consumed while every nearer one was published; the walk ends at the first inclusive prefix.
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            auto consume = [&](const LB (&v)[kBatch]) {
                bool alive = true;
                int used = 0;
#pragma unroll
                for (int k = 0; k < kBatch; ++k) {
                    const LB f = v[k] & kFlags;
                    const bool take = alive && f != 0;
                    if (take) { excl += v[k] & ~kFlags; ++used; }
                    if (take && f == kIncl) done = true;
                    alive = take && f != kIncl;
                }
                t -= used;                                  // an unpublished predecessor is simply re-read
            };
#if IBVH_SORT_PREFETCH
            consume(pre);
#endif
            while (!done) {
                LB v[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k) v[k] = (t - k >= 0) ? lookback[(size_t)(t - k) * kRadixBins + d] : kIncl;
                consume(v);
            }
            lookback[(size_t)tile * kRadixBins + d] = kIncl | (excl + (LB)tcount);
        }
        gofs[d] = hist_excl[d] + (uint32_t)excl - tile_start[d];
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive positions with one digit go to consecutive addresses ----------
    auto emit = [&](int i) {
        K kk; uint32_t vv;
        if constexpr (kPairStage) { const uint2 kv = reinterpret_cast<const uint2*>(onesweep_smem)[i]; kk = (K)kv.x; vv = kv.y; }
        else { kk = skeys[i]; vv = svals[i]; }
        const uint32_t dst = gofs[digit_of(kk, shift)] + (uint32_t)i;
        keys_out[dst] = kk;
        vals_out[dst] = vv;
    };
    if (full) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) emit(k * THREADS + tid);
    } else {
        for (int i = tid; i < tile_n; i += THREADS) emit(i);
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
