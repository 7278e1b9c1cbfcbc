// peer.cu — all-gather of contact / hit shards over NVLink peer memory, one kernel per rank
// (include/ibvh.h: ibvh_allgather_pairs). The reference has no multi-GPU path.
//
// Header layout and signalling primitives: peer.cuh. Payload goes out as multimem.st (one store lands in every
// rank's buffer through the NVSwitch multicast alias) or as plain peer stores when there is no multicast alias.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/ibvh.h"
#include "handle.cuh"
#include "peer.cuh"

using namespace ibvh;

namespace {

constexpr int kDoneSlot = kPeerDoneSlot;
constexpr uint64_t kCountMask = (uint64_t(1) << 40) - 1;
constexpr uint64_t kTimeoutNs = kPeerTimeoutNs;
constexpr size_t kSmallPeerTicket = 1024;   // u32 inside handle::d_small

// h_out (pinned): [0] total pairs, [1] offset of this rank's shard, [2] status (0 ok, 1 timeout, 2 capacity)
__global__ void __launch_bounds__(512) peer_allgather_kernel(PeerArgs a, const uint64_t* __restrict__ shard, int64_t count, int words_per_pair,
                                                             uint32_t* ticket, int64_t* h_out) {
    __shared__ int64_t s_off, s_total;
    __shared__ int s_status, s_last;
    uint64_t* sig = (uint64_t*)a.buf[a.rank];
    const uint64_t tag = a.epoch & 0xFFFFFF;

    // (1) counts: block 0 announces, every block reads its own (local) header
    if (blockIdx.x == 0 && threadIdx.x < a.world)
        st_release_sys((uint64_t*)a.buf[threadIdx.x] + a.rank, (tag << 40) | (uint64_t(count) & kCountMask));
    if (threadIdx.x == 0) {
        int status = 0;
        int64_t off = 0, tot = 0;
        const uint64_t t0 = globaltimer_ns();
        for (int r = 0; r < a.world && !status; ++r) {
            uint64_t v;
            while (((v = ld_acquire_sys(sig + r)) >> 40) != tag) {
                if (globaltimer_ns() - t0 > kTimeoutNs) { status = 1; break; }
                __nanosleep(64);
            }
            const int64_t c = int64_t(v & kCountMask);
            if (r < a.rank) off += c;
            tot += c;
        }
        if (!status && tot * words_per_pair * 8 > a.capacity_bytes) status = 2;
        s_off = off; s_total = tot; s_status = status;
    }
    __syncthreads();
    const int status = s_status;

    // (2) payload
    if (!status) {
        const int64_t n_words = count * words_per_pair;
        const int64_t dst0 = s_off * words_per_pair;
        const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
        if (a.mc) {
            uint64_t* dst = (uint64_t*)(a.mc + a.header_bytes) + dst0;
            if ((dst0 & 1) == 0 && (reinterpret_cast<uintptr_t>(shard) & 15) == 0) {   // 16-byte aligned on BOTH sides (the shard is a caller pointer: only 8 bytes are promised): 128-bit multicast stores
                const int64_t n4 = n_words >> 1;
                const uint4* s4 = (const uint4*)shard;
                for (int64_t i = tid; i < n4; i += nth) multimem_st_v4((uint4*)dst + i, s4[i]);
                if (tid == 0 && (n_words & 1)) multimem_st_u64(dst + n_words - 1, shard[n_words - 1]);
            } else {
                for (int64_t i = tid; i < n_words; i += nth) multimem_st_u64(dst + i, shard[i]);
            }
        } else {
            for (int64_t i = tid; i < n_words; i += nth) {
                const uint64_t v = shard[i];
                for (int k = 0; k < a.world; ++k) {
                    const int p = (a.rank + k) % a.world;     // staggered: ranks hit different peers at the same time
                    ((uint64_t*)(a.buf[p] + a.header_bytes))[dst0 + i] = v;
                }
            }
        }
    }

    // (3) completion: last block of this rank signals every peer, then waits for all of them
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    // The done barrier also runs when the list area is too small (status 2: the same verdict on every rank): a fast
    // rank's next collective must not overwrite the count slots while a slow rank still spins on this epoch's tag.
    // Only a timeout (a peer that never arrived) skips it.
    int st = status;
    if (status != 1 && threadIdx.x < a.world) {
        st_release_sys((uint64_t*)a.buf[threadIdx.x] + kDoneSlot + a.rank, a.epoch);
        const uint64_t t0 = globaltimer_ns();
        while (ld_acquire_sys(sig + kDoneSlot + threadIdx.x) != a.epoch) {
            if (globaltimer_ns() - t0 > kTimeoutNs) { st = 1; break; }
            __nanosleep(64);
        }
    }
    st = __syncthreads_or(st == 1) ? 1 : status;
    if (threadIdx.x == 0) {
        *ticket = 0;
        h_out[0] = s_total; h_out[1] = s_off; h_out[2] = st;
        __threadfence_system();
    }
}

}  // namespace

extern "C" IBVH_API int ibvh_allgather_pairs(ibvh_handle_t* h, const ibvh_peer_t* peer, const void* d_shard, int64_t count,
                                             int32_t pair_bytes, int64_t* out_total, int64_t* out_offset, void* stream) {
    if (!h) return IBVH_ERR_ARGUMENT;
    if (!peer_ok(peer) || count < 0 || (pair_bytes != 8 && pair_bytes != 16) || (count > 0 && !d_shard) || (uint64_t(count) > kCountMask)) {
        h->set_error("ibvh_allgather_pairs: bad arguments");
        return IBVH_ERR_ARGUMENT;
    }
    ibvh::DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    PeerArgs a = make_peer_args(peer);
    if (h->cfg.peer_no_multicast) a.mc = 0;           // debug knob (IBVH_PEER_NO_MULTICAST, read at ibvh_create): plain peer stores
    int64_t* h_out = (int64_t*)(h->h_pinned + 3072);
    h_out[2] = -1;
    const int grid = h->sm_count * 2;                 // all CTAs co-resident: every block spins on the local header
    {
        ibvh::ProfScope ps(h, st, "peer_allgather_kernel");
        peer_allgather_kernel<<<grid, 512, 0, st>>>(a, (const uint64_t*)d_shard, count, pair_bytes / 8, (uint32_t*)(h->d_small + kSmallPeerTicket), h_out);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { h->set_cuda_error(e, "peer_allgather_kernel"); return IBVH_ERR_CUDA; }
    if (out_total) *out_total = h_out[0];
    if (out_offset) *out_offset = h_out[1];
    if (h_out[2] == 1) { h->set_error("ibvh_allgather_pairs: a peer did not arrive within 10 s"); return IBVH_ERR_PEER; }
    if (h_out[2] == 2) { h->set_error("ibvh_allgather_pairs: list area too small for the gathered pairs"); return IBVH_ERR_CAPACITY; }
    if (h_out[2] != 0) { h->set_error("ibvh_allgather_pairs: kernel did not report"); return IBVH_ERR_CUDA; }
    return IBVH_OK;
}

// Host-only: the gap-filling moves a fused traversal applies to its segmented list (peer.cuh: make_compact_moves), exposed
// so that the multi-GPU host logic can be tested without a GPU. Units are list ENTRIES. Returns the number of moves.
extern "C" IBVH_API int ibvh_peer_compact_plan(int32_t world, const int64_t* region_begin, const int64_t* counts,
                                               int64_t* src, int64_t* dst, int64_t* len, int32_t max_moves) {
    if (world < 1 || world > IBVH_MAX_PEERS || !region_begin || !counts || !src || !dst || !len) return -1;
    long long b[IBVH_MAX_PEERS + 1], c[IBVH_MAX_PEERS];
    for (int r = 0; r < world; ++r) { b[r] = region_begin[r]; c[r] = counts[r]; }
    b[world] = region_begin[world];
    const PeerMoves mv = make_compact_moves(world, b, c, 1);
    if (mv.n > max_moves) return -1;
    for (int k = 0; k < mv.n; ++k) { src[k] = mv.src[k]; dst[k] = mv.dst[k]; len[k] = mv.len[k]; }
    return mv.n;
}
