// peer.cuh — device helpers of the multi-GPU exchange over NVLink peer memory (include/ibvh.h: ibvh_peer_t).
//
// Header layout of every rank's symmetric buffer (u64 slots, zeroed once at creation):
//   [r]        (r < 16)  ibvh_allgather_pairs: count announced by rank r, (epoch & 0xFFFFFF) << 40 | count
//   [128 + r]            ibvh_allgather_pairs: "rank r has finished writing" = epoch
//   [192 + r]            fused traversal: "rank r has finished its shard" = fused_seq
//   [256 .. 258]         fused traversal, rank 0 only: output-slot counters, rotated by fused_seq % 3
// Signals are st.release.sys / ld.acquire.sys on peer-mapped global memory; payload goes out as multimem.st
// (one store lands in every rank's buffer through the NVSwitch multicast alias).
#pragma once
#include <cstdint>

#include "../../include/ibvh.h"

namespace ibvh {

constexpr int kPeerCountSlot = 0;
constexpr int kPeerDoneSlot = 128;
constexpr int kPeerFusedDoneSlot = 192;
constexpr int kPeerCounterSlot = 256;
constexpr uint64_t kPeerTimeoutNs = 10ull * 1000 * 1000 * 1000;

struct PeerArgs {
    uint64_t buf[IBVH_MAX_PEERS];
    uint64_t mc;
    int64_t header_bytes, capacity_bytes;
    uint64_t epoch, fused_seq;
    int rank, world;
};

// The rank's region of the gathered list in entries of `pair_bytes`: the caller's region_begin[] if it set them
// (region_begin[world] > 0), else an equal split of the list area.
inline void peer_regions(const ibvh_peer_t* p, int pair_bytes, long long* begin /* world + 1 */) {
    const long long cap = p->capacity_bytes / pair_bytes;
    if (p->region_begin[p->world] > 0) {
        for (int r = 0; r <= p->world; ++r) begin[r] = p->region_begin[r] < cap ? p->region_begin[r] : cap;
        begin[0] = 0;
        for (int r = 1; r <= p->world; ++r) if (begin[r] < begin[r - 1]) begin[r] = begin[r - 1];
    } else {
        const long long each = (cap / p->world) & ~1ll;              // even: 16-byte aligned region starts for 8-byte pairs
        for (int r = 0; r <= p->world; ++r) begin[r] = each * r;
    }
    for (int r = 0; r <= p->world; ++r) begin[r] &= ~1ll;
}

inline PeerArgs make_peer_args(const ibvh_peer_t* p) {
    PeerArgs a;
    for (int r = 0; r < IBVH_MAX_PEERS; ++r) a.buf[r] = r < p->world ? p->buffers[r] : 0;
    a.mc = p->multicast;
    a.header_bytes = p->header_bytes; a.capacity_bytes = p->capacity_bytes;
    a.epoch = p->epoch; a.fused_seq = p->fused_seq; a.rank = p->rank; a.world = p->world;
    return a;
}

inline bool peer_ok(const ibvh_peer_t* p) {
    if (!p || p->world < 1 || p->world > IBVH_MAX_PEERS || p->rank < 0 || p->rank >= p->world || p->epoch == 0 ||
        p->header_bytes < 512 * 8 || (p->header_bytes & 255) || p->capacity_bytes < 0)
        return false;
    for (int r = 0; r < p->world; ++r)
        if (!p->buffers[r]) return false;
    return true;
}

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_u64(void* p, uint64_t v) {
    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_v4(void* p, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one 8- or 16-byte record through the multicast alias
template <class P> __device__ __forceinline__ void multimem_store_pair(P* dst, const P& v) {
    static_assert(sizeof(P) == 8 || sizeof(P) == 16, "index pair");
    if constexpr (sizeof(P) == 8) multimem_st_u64(dst, *reinterpret_cast<const uint64_t*>(&v));
    else multimem_st_v4(dst, *reinterpret_cast<const uint4*>(&v));
}

// ---- fused traversal + all-gather, round-2 protocol ---------------------------------------------------------
// Every rank owns a contiguous REGION of the gathered list (ibvh_peer_t.region_begin, sized by the caller from the
// previous step's per-rank counts): output slots are reserved with LOCAL atomics on a counter in the rank's own
// memory and the contacts go out through the multicast alias at region_begin[rank] + slot. Round 1 reserved every
// slot with a system-scope atomic on ONE counter of rank 0 (612 k NVLink atomics per 100 M-ray step, ~5 ns each,
// all serialised on one address): that was the limiter of the 8-GPU runs.
//   begin  (one warp, first kernel of the call): "rank r has no reader of the previous list left" -> every peer
//   wait   (one warp, right before the kernel that writes into the peers' lists): all ranks have said so
//   finish (one warp, after it): publish this rank's count to every peer, wait for all counts (which also says
//          "all of rank r's stores are out"), hand counts + total to the host
//   compact (optional): the regions carry slack, so the list has gaps: the entries beyond `total` are moved into the
//          gaps below it (the list is unordered anyway) — identical moves on every rank, local memory only.
// Signal words live in two banks selected by fused_seq & 1: a rank can only write bank (seq + 2) & 1 == seq & 1 again
// after it has finished call seq + 1, which needs every peer's word of call seq + 1, which a peer writes only after
// it has finished reading the words of call seq.
constexpr int kPeerFusedCountSlot2 = 192;   // + 16 * (seq & 1) + r : (seq & 0xFFFFFF) << 40 | count of rank r
constexpr int kPeerFusedReadySlot = 224;    // + 16 * (seq & 1) + r : seq = "rank r is ready to receive call seq"

static __global__ void peer_fused_begin_kernel(PeerArgs a) {
    const int lane = threadIdx.x;
    __threadfence_system();
    if (lane < a.world) st_release_sys((uint64_t*)a.buf[lane] + kPeerFusedReadySlot + 16 * (int)(a.fused_seq & 1) + a.rank, a.fused_seq);
}
// h_out (pinned): [1] = 1 on timeout
static __global__ void peer_fused_wait_kernel(PeerArgs a, int64_t* h_out) {
    const int lane = threadIdx.x;
    const uint64_t* sig = (const uint64_t*)a.buf[a.rank] + kPeerFusedReadySlot + 16 * (int)(a.fused_seq & 1);
    int st = 0;
    if (lane < a.world) {
        const uint64_t t0 = globaltimer_ns();
        while (ld_acquire_sys(sig + lane) != a.fused_seq) {
            if (globaltimer_ns() - t0 > kPeerTimeoutNs) { st = 1; break; }
            __nanosleep(64);
        }
    }
    st = __any_sync(0xffffffffu, st);
    if (lane == 0 && st) { h_out[1] = 1; __threadfence_system(); }
}
// h_out (pinned): [0] gathered total, [1] status (0 ok, 1 timeout), [2 + r] count of rank r
static __global__ void peer_fused_finish_kernel(PeerArgs a, const unsigned long long* local_count, int64_t* h_out) {
    const int lane = threadIdx.x;
    const int bank = 16 * (int)(a.fused_seq & 1);
    const uint64_t tag = a.fused_seq & 0xFFFFFF;
    const uint64_t mine = (uint64_t)*local_count & ((uint64_t(1) << 40) - 1);
    __threadfence_system();                       // this rank's multicast stores are ordered before the count it publishes
    int st = 0;
    uint64_t c = 0;
    if (lane < a.world) {
        st_release_sys((uint64_t*)a.buf[lane] + kPeerFusedCountSlot2 + bank + a.rank, (tag << 40) | mine);
        const uint64_t* sig = (const uint64_t*)a.buf[a.rank] + kPeerFusedCountSlot2 + bank + lane;
        const uint64_t t0 = globaltimer_ns();
        uint64_t v;
        while (((v = ld_acquire_sys(sig)) >> 40) != tag) {
            if (globaltimer_ns() - t0 > kPeerTimeoutNs) { st = 1; break; }
            __nanosleep(64);
        }
        c = v & ((uint64_t(1) << 40) - 1);
        h_out[2 + lane] = (int64_t)c;
    }
    st = __any_sync(0xffffffffu, st);
    uint64_t tot = c;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    if (lane == 0) {
        h_out[0] = (int64_t)tot;
        h_out[1] = st ? 1 : (h_out[1] == 1 ? 1 : 0);
        __threadfence_system();
    }
}

// gap filling of a segmented list in local memory: up to kPeerMaxMoves (src, dst, len) runs of 8-byte words
constexpr int kPeerMaxMoves = 2 * IBVH_MAX_PEERS + 2;
struct PeerMoves { int n; long long src[kPeerMaxMoves], dst[kPeerMaxMoves], len[kPeerMaxMoves]; };
static __global__ void __launch_bounds__(256) peer_compact_kernel(uint64_t* list, PeerMoves mv) {
    for (int m = 0; m < mv.n; ++m) {
        const uint64_t* s = list + mv.src[m];
        uint64_t* d = list + mv.dst[m];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < mv.len[m]; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
    }
}

// Host side: the moves that close the gaps of a list whose rank r holds counts[r] entries from begin[r] on.
// Sources are the entries at positions >= total (taken from the top down is not needed: any bijection does),
// destinations the uncovered positions < total. words = 8-byte words per entry.
inline PeerMoves make_compact_moves(int world, const long long* begin, const long long* counts, int words) {
    PeerMoves mv; mv.n = 0;
    long long total = 0;
    for (int r = 0; r < world; ++r) total += counts[r];
    long long gs[IBVH_MAX_PEERS + 1], ge[IBVH_MAX_PEERS + 1]; int ng = 0;      // gaps below total
    long long ss[IBVH_MAX_PEERS + 1], se[IBVH_MAX_PEERS + 1]; int ns = 0;      // source runs at or above total
    long long pos = 0;
    for (int r = 0; r < world; ++r) {
        const long long b = begin[r], e = begin[r] + counts[r];
        if (b > pos && pos < total) { gs[ng] = pos; ge[ng] = b < total ? b : total; if (ge[ng] > gs[ng]) ++ng; }
        if (e > total) { ss[ns] = b > total ? b : total; se[ns] = e; if (se[ns] > ss[ns]) ++ns; }
        if (e > pos) pos = e;
    }
    if (pos < total) { gs[ng] = pos; ge[ng] = total; ++ng; }                        // (cannot happen: sum of counts == total)
    int gi = 0, si = 0;
    while (gi < ng && si < ns && mv.n < kPeerMaxMoves) {
        const long long len = (ge[gi] - gs[gi]) < (se[si] - ss[si]) ? (ge[gi] - gs[gi]) : (se[si] - ss[si]);
        mv.src[mv.n] = ss[si] * words; mv.dst[mv.n] = gs[gi] * words; mv.len[mv.n] = len * words; ++mv.n;
        gs[gi] += len; ss[si] += len;
        if (gs[gi] == ge[gi]) ++gi;
        if (ss[si] == se[si]) ++si;
    }
    return mv;
}

}  // namespace ibvh
