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
// 8-byte keys (8 passes, 10 M): 256x16 (3) 0.582, 512x12 (2) 0.623, 256x16 (4) 0.662, 512x16 (2) 0.679, 1024x8 (1) 0.741 ms.
// Digit width (IBVH_SORT_RB32 / RB64 in morton.cuh; same harness, gpurun_out/r2s): 30-bit keys in THREE 10-bit passes
// 0.263 ms against 0.237 for four 8-bit ones; 63-bit keys in SEVEN 9-bit passes 0.610 against 0.582 for eight 8-bit ones.
// A pass gets dearer faster than passes get fewer: the per-warp digit counters (warps x bins: 16 K counters beside an
// 8 K-key tile at 10 bits) are zeroed, scanned over the warps and looked back once per tile. 8 bits stays.
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
// threads per tile, keys per thread, resident tiles per SM the register budget is set for
#ifdef IBVH_SORT_THREADS
template <class K> constexpr int sort_threads() { return IBVH_SORT_THREADS; }
#else
template <class K> constexpr int sort_threads() { return sizeof(K) == 8 ? 256 : 512; }
#endif
#ifdef IBVH_SORT_ITEMS
template <class K> constexpr int sort_items() { return IBVH_SORT_ITEMS; }
#else
template <class K> constexpr int sort_items() { return 16; }
#endif
#ifdef IBVH_SORT_MINB
template <class K> constexpr int sort_minb() { return IBVH_SORT_MINB; }
#else
template <class K> constexpr int sort_minb() { return sizeof(K) == 8 ? 3 : 2; }
#endif
template <class K> constexpr int sort_tile() { return sort_threads<K>() * sort_items<K>(); }

// look-back words: 2 flag bits on top of a count / prefix
template <class LB> struct LookbackWord;
template <> struct LookbackWord<uint32_t> {
    static constexpr uint32_t kAgg = 1u << 30, kIncl = 2u << 30, kFlags = 3u << 30;
};
template <> struct LookbackWord<unsigned long long> {
    static constexpr unsigned long long kAgg = 1ull << 62, kIncl = 2ull << 62, kFlags = 3ull << 62;
};

// exclusive scan of each pass's histogram (BINS <= 1024 bins), in place. grid = passes, block = 1024.
template <int BINS>
static __global__ void __launch_bounds__(1024) scan_hist_kernel(uint32_t* hist) {
    __shared__ uint32_t wsum[32];
    uint32_t* h = hist + blockIdx.x * BINS;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t v = t < BINS ? h[t] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int j = 0; j < w; ++j) base += wsum[j];
    if (t < BINS) h[t] = base + incl - v;
}

template <int BINS, class K> IBVH_D uint32_t digit_of(K key, int shift) { return (uint32_t)(key >> shift) & (uint32_t)(BINS - 1); }

// One bit of the warp-wide digit match: keep the lanes whose bit B of the digit equals mine. Written in PTX because
// nvcc derived TWO predicates per bit from the C++ form (shift + and + setp for the ballot, and + setp + sel for the
// mask: 5.5 instructions per bit); this is and+setp (one LOP3, or one R2P for a whole byte), VOTE, a predicated NOT
// and an AND: ~3.4 per bit.
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
    static_assert(BITS >= 1 && BITS <= 10, "digit width");
    uint32_t peers = 0xffffffffu;
    match_bit<0>(d, peers);
    if constexpr (BITS > 1) match_bit<1>(d, peers);
    if constexpr (BITS > 2) match_bit<2>(d, peers);
    if constexpr (BITS > 3) match_bit<3>(d, peers);
    if constexpr (BITS > 4) match_bit<4>(d, peers);
    if constexpr (BITS > 5) match_bit<5>(d, peers);
    if constexpr (BITS > 6) match_bit<6>(d, peers);
    if constexpr (BITS > 7) match_bit<7>(d, peers);
    if constexpr (BITS > 8) match_bit<8>(d, peers);
    if constexpr (BITS > 9) match_bit<9>(d, peers);
    return peers;
}

template <class K, int THREADS, int ITEMS, int BINS> constexpr size_t onesweep_smem_bytes() {
    return (size_t)THREADS * ITEMS * (sizeof(K) + 4) + (size_t)(THREADS / 32) * BINS * 2 + 2 * (size_t)BINS * 4;
}

// One radix pass. keys_in/vals_in -> keys_out/vals_out. vals_in == nullptr: values are the global
// item index (first pass: the permutation starts as iota and need not be read).
// hist_excl: this pass's exclusive-scanned global histogram. lookback: [tiles][BINS] zero-initialised.
// RB: radix bits of the sort (BINS = 2^RB digit values); BITS: digit bits that can differ in this pass (RB, or what
// is left of the key in the top pass). sentinel: a key with every key bit set (pads the last tile).
// Thread t owns the digits t, t + THREADS, ... in the scans and in the look-back.
template <class K, class LB, int THREADS, int ITEMS, int RB, int BITS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) onesweep_kernel(const K* __restrict__ keys_in, K* __restrict__ keys_out,
                                                                const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                                                                int64_t n, const uint32_t* __restrict__ hist_excl,
                                                                volatile LB* lookback, uint32_t* ticket, int shift, K sentinel) {
    constexpr int BINS = 1 << RB;
    constexpr int DPT = (BINS + THREADS - 1) / THREADS;        // digits per thread
    static_assert(THREADS % 32 == 0, "whole warps");
    static_assert(32 * ITEMS <= 65535 && THREADS * ITEMS <= 65535, "16-bit per-warp counters / offsets");
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    constexpr bool kPairStage = IBVH_SORT_PAIRSTAGE && sizeof(K) == 4;      // stage (key, value) as one 8-byte record
    constexpr LB kAgg = LookbackWord<LB>::kAgg, kIncl = LookbackWord<LB>::kIncl, kFlags = LookbackWord<LB>::kFlags;
    extern __shared__ __align__(16) unsigned char onesweep_smem[];
    K* skeys = reinterpret_cast<K*>(onesweep_smem);                                   // TILE keys, tile-sorted
    uint32_t* svals = reinterpret_cast<uint32_t*>(skeys + TILE);                      // TILE values
    uint16_t* whist = reinterpret_cast<uint16_t*>(svals + TILE);                      // [WARPS][BINS] counts -> exclusive offsets over warps
    uint32_t* tile_start = reinterpret_cast<uint32_t*>(whist + WARPS * BINS);         // exclusive scan of the tile's digit counts
    uint32_t* gofs = tile_start + BINS;                                               // global position - position in the tile order (mod 2^32)
    __shared__ uint32_t s_tile;
    __shared__ uint32_t wsum[DPT][WARPS];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * BINS / 8; i += THREADS) reinterpret_cast<uint4*>(whist)[i] = make_uint4(0u, 0u, 0u, 0u);
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
        uint16_t* wh = whist + w * BINS;
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const uint32_t d = digit_of<BINS>(key[r], shift);
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

    // ---- per digit (thread t: digits t + k * THREADS): exclusive offsets over warps, tile count, publish, tile-level scan ----
    constexpr int kBatch = IBVH_SORT_BATCH;
    uint32_t tcount[DPT], incl[DPT];
    const int dsent = (int)digit_of<BINS>(sentinel, shift);
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int d = tid + k * THREADS;
        tcount[k] = 0; incl[k] = 0;
        if (d < BINS) {
            uint32_t tc = 0;
#pragma unroll
            for (int j = 0; j < WARPS; ++j) { const uint32_t c = whist[j * BINS + d]; whist[j * BINS + d] = (uint16_t)tc; tc += c; }
            // sentinels of a partial tile were ranked like keys: they hold the largest digit and the last tile positions
            if (tile_n < TILE && d == dsent) tc -= (uint32_t)(TILE - tile_n);
            lookback[(size_t)tile * BINS + d] = (tile == 0 ? kIncl : kAgg) | (LB)tc;
            tcount[k] = tc;
        }
        uint32_t in = tcount[k];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, in, off); if (lane >= off) in += o; }
        incl[k] = in;
        if (lane == 31) wsum[k][w] = in;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int d = tid + k * THREADS;
        if (d < BINS) {
            uint32_t base = 0;
            for (int kk = 0; kk < k; ++kk)
                for (int j = 0; j < WARPS; ++j) base += wsum[kk][j];
            for (int j = 0; j < w; ++j) base += wsum[k][j];
            tile_start[d] = base + incl[k] - tcount[k];
        }
    }
    __syncthreads();

    // ---- reorder the tile in shared memory -------------------------------------------------------------
    {
        const uint16_t* wh = whist + w * BINS;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const uint32_t d = digit_of<BINS>(key[r], shift);
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

    // ---- decoupled look-back (thread t resolves its digits) --------------------------------------------
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int d = tid + k * THREADS;
        if (d < BINS) {
            LB excl = 0;
            if (tile > 0) {
                // Walk back over the predecessors' (flag | count) words, kBatch independent loads in flight per round: the
                // chain to the nearest published inclusive prefix is as long as the number of tiles in flight. An entry is
                // consumed while every nearer one was published; the walk ends at the first inclusive prefix.
                int64_t t = (int64_t)tile - 1;
                bool done = false;
                while (!done) {
                    LB v[kBatch];
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) v[q] = (t - q >= 0) ? lookback[(size_t)(t - q) * BINS + d] : kIncl;
                    bool alive = true;
                    int used = 0;
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) {
                        const LB f = v[q] & kFlags;
                        const bool take = alive && f != 0;
                        if (take) { excl += v[q] & ~kFlags; ++used; }
                        if (take && f == kIncl) done = true;
                        alive = take && f != kIncl;
                    }
                    t -= used;                                  // an unpublished predecessor is simply re-read
                }
                lookback[(size_t)tile * BINS + d] = kIncl | (excl + (LB)tcount[k]);
            }
            gofs[d] = hist_excl[d] + (uint32_t)excl - tile_start[d];
        }
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive positions with one digit go to consecutive addresses ----------
    auto emit = [&](int i) {
        K kk; uint32_t vv;
        if constexpr (kPairStage) { const uint2 kv = reinterpret_cast<const uint2*>(onesweep_smem)[i]; kk = (K)kv.x; vv = kv.y; }
        else { kk = skeys[i]; vv = svals[i]; }
        const uint32_t dst = gofs[digit_of<BINS>(kk, shift)] + (uint32_t)i;
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
    constexpr int RB = radix_bits<M>(), BINS = radix_bins<M>();
    __shared__ uint32_t sh[P][BINS];
    for (int i = threadIdx.x; i < P * BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Words<L> wv = load_words(leaves + i);
        M m = words_morton<L>(wv);
        keys[i] = m;
        store_words(copy_out + i, wv);
#pragma unroll
        for (int p = 0; p < P; ++p) atomicAdd(&sh[p][(uint32_t)(m >> (p * RB)) & (BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * BINS; i += blockDim.x) {
        uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ---- host driver ---------------------------------------------------------------------------------------
// keysA holds the input keys; hist the raw per-pass histograms ([passes][BINS]); lookback is zeroed and holds
// passes * tiles * BINS words of LB. On return the sorted keys / the permutation are in *keys_out / *vals_out.
// a key with every key bit set (15 / 30 / 63 bits): pads the last tile
template <class K> constexpr K sort_sentinel() { return (K)((1ull << MortonTraits<K>::key_bits) - 1ull); }

template <class K, class LB, int THREADS, int ITEMS, int MINB, int BITS>
inline cudaError_t launch_onesweep_pass(const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, int64_t n, const uint32_t* hist_excl,
                                        LB* lookback, uint32_t* ticket, int shift, cudaStream_t st) {
    constexpr int RB = radix_bits<K>();
    auto kern = onesweep_kernel<K, LB, THREADS, ITEMS, RB, BITS, MINB>;
    constexpr size_t smem = onesweep_smem_bytes<K, THREADS, ITEMS, (1 << RB)>();
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
    constexpr int RB = radix_bits<K>(), BINS = radix_bins<K>();
    constexpr int kTopBits = MortonTraits<K>::key_bits - RB * (P - 1);
    const int64_t tiles = (n + (int64_t)THREADS * ITEMS - 1) / ((int64_t)THREADS * ITEMS);
    LB* lookback = (LB*)lookback_raw;
    {
        auto s = scope("scan_hist_kernel");
        scan_hist_kernel<BINS><<<P, 1024, 0, st>>>(hist);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    K* kin = keysA; K* kout = keysB;
    uint32_t* vin = nullptr; uint32_t* vout = valsB; uint32_t* vother = valsA;
    for (int p = 0; p < P; ++p) {
        {
            auto s = scope("onesweep_kernel");
            if (p == P - 1 && kTopBits < RB)
                e = launch_onesweep_pass<K, LB, THREADS, ITEMS, MINB, kTopBits>(kin, kout, vin, vout, n, hist + p * BINS, lookback + (size_t)p * tiles * BINS, tickets + p, p * RB, st);
            else
                e = launch_onesweep_pass<K, LB, THREADS, ITEMS, MINB, RB>(kin, kout, vin, vout, n, hist + p * BINS, lookback + (size_t)p * tiles * BINS, tickets + p, p * RB, st);
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
