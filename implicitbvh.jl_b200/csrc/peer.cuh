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

// Tail of a fused traversal (one warp): rank 0 first zeroes the counter the NEXT fused call will use; every rank
// tells every peer "my shard is written" and waits for all of them; then the gathered total is read from rank 0.
// h_out (pinned): [0] total, [1] status (0 ok, 1 timeout).
// Counter rotation mod 3: the counter zeroed here was last read in call seq-2, and every rank's read of call
// seq-2 precedes (stream order) its "done" signal of call seq-1, which rank 0 has waited for before it got here;
// adds of call seq+1 start only after the adder has seen rank 0's "done" of this call, i.e. after the zeroing.
static __global__ void peer_fused_finish_kernel(PeerArgs a, int64_t* h_out) {
    const int lane = threadIdx.x;
    uint64_t* sig = (uint64_t*)a.buf[a.rank];
    if (a.rank == 0 && lane == 0) sig[kPeerCounterSlot + (a.fused_seq + 1) % 3] = 0;
    __syncwarp();
    __threadfence_system();
    int st = 0;
    if (lane < a.world) {
        st_release_sys((uint64_t*)a.buf[lane] + kPeerFusedDoneSlot + a.rank, a.fused_seq);
        const uint64_t t0 = globaltimer_ns();
        while (ld_acquire_sys(sig + kPeerFusedDoneSlot + lane) != a.fused_seq) {
            if (globaltimer_ns() - t0 > kPeerTimeoutNs) { st = 1; break; }
            __nanosleep(64);
        }
    }
    st = __any_sync(0xffffffffu, st);
    if (lane == 0) {
        const uint64_t total = ld_acquire_sys((const uint64_t*)a.buf[0] + kPeerCounterSlot + a.fused_seq % 3);
        h_out[0] = (int64_t)total;
        h_out[1] = st;
        __threadfence_system();
    }
}

}  // namespace ibvh
