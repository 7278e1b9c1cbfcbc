// handle.cuh — the opaque per-device handle: error text, SM count, a grow-only scratch arena
// (replaces the temporaries AcceleratedKernels allocates inside AK.sort! / AK.accumulate! /
// AK.mapreduce) and a small pinned host block for scalar read-backs.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

struct ibvh_handle {
    int device = 0;
    int sm_count = 148;
    char err[512] = {0};

    // development / debug knobs: environment variables read ONCE, at ibvh_create (never on the call path)
    struct Config {
        bool fused_gather = false;            // IBVH_FUSED_GATHER: gather fused into the bottom-level merge kernel
        bool rays_reference_shaped = false;   // IBVH_RAYS_REFERENCE_SHAPED: one thread per ray, local stack (proxy of the reference kernel)
        bool rays_static = false;             // IBVH_RAYS_STATIC: one ray per thread, no persistent refill
        int rays_wide_div = 16;               // IBVH_RAYS_WIDE_DIV: the long-ray queue holds rays / this many entries (32 bytes each; 4096 .. 4 M)
        int rays_wide = 512;                  // IBVH_RAYS_WIDE: node steps after which a long ray is exported to rays_wide_kernel (one warp per ray); 0 = never
        bool debug = false;                   // IBVH_DEBUG: print schedule statistics to stderr
        bool peer_no_multicast = false;       // IBVH_PEER_NO_MULTICAST: plain peer stores instead of multimem.st
        bool force_wide_lookback = false;     // IBVH_SORT_WIDE_LOOKBACK: 64-bit look-back words at any size (tests the n >= 2^30 path)
        bool no_sidecar = false;              // IBVH_NO_SIDECAR: the build does not keep the traversal's packed records
        int pyr_grid = 40;                    // IBVH_PYR_GRID: CTAs launched per SM by the leaf-tile kernel (8 resident; they draw chunks from a ticket counter)
        int pyr_grid_refine = 8;              // IBVH_PYR_GRID_REFINE: ... by the refine kernels (7 resident). Swept at 10 M leaves with the final kernels
                                              // (gpurun_out/r2aq): CTAs per SM 8 / 12 / 16 / 20 / 28 / 40 -> refine 0.659 / 0.664 / 0.670 / 0.675 / 0.694 /
                                              // 0.769 ms, leaf tiles 1.133 / 1.128 / 1.128 / 1.126 / 1.122 / 1.118 ms (both were 20 before)
        bool pyr_quant = true;                // IBVH_PYR_QUANT=0: refine over the float boxes instead of the conservatively quantised ones
        int pyr_q2 = 2;                       // IBVH_PYR_Q2: query children per lane of the quantised refine kernel: 2 (default), 4, or 0 = the one-child form
        bool pyr_tma = false;                 // IBVH_PYR_TMA=1: refine kernel with TMA bulk copies + mbarrier instead of LDG -> STS (measured slower: see traverse_pyramid.cuh)
        void parse() {
            auto on = [](const char* k) { const char* v = getenv(k); return v != nullptr && v[0] != '\0' && !(v[0] == '0' && v[1] == '\0'); };
            fused_gather = on("IBVH_FUSED_GATHER");
            rays_reference_shaped = on("IBVH_RAYS_REFERENCE_SHAPED");
            rays_static = on("IBVH_RAYS_STATIC");
            if (const char* v = getenv("IBVH_RAYS_WIDE")) { int g = atoi(v); if (g >= 0) rays_wide = g; }
            if (const char* v = getenv("IBVH_RAYS_WIDE_DIV")) { int g = atoi(v); if (g > 0) rays_wide_div = g; }
            debug = on("IBVH_DEBUG");
            peer_no_multicast = on("IBVH_PEER_NO_MULTICAST");
            force_wide_lookback = on("IBVH_SORT_WIDE_LOOKBACK");
            no_sidecar = on("IBVH_NO_SIDECAR");
            pyr_tma = on("IBVH_PYR_TMA");
            if (const char* v = getenv("IBVH_PYR_QUANT")) pyr_quant = !(v[0] == '0' && v[1] == '\0');
            if (const char* v = getenv("IBVH_PYR_Q2")) { int g = atoi(v); pyr_q2 = (g == 4) ? 4 : (g == 0 ? 0 : 2); }
            if (const char* v = getenv("IBVH_PYR_GRID")) { int g = atoi(v); if (g > 0) pyr_grid = g; }
            if (const char* v = getenv("IBVH_PYR_GRID_REFINE")) { int g = atoi(v); if (g > 0) pyr_grid_refine = g; }
        }
    } cfg;

    // grow-only arena
    char* ws = nullptr;
    size_t ws_bytes = 0;
    size_t ws_off = 0;

    // second grow-only buffer: the (query group -> target group) lists of the tiled traversal
    char* aux = nullptr;
    size_t aux_bytes = 0;
    int reserve_aux(size_t bytes) {
        if (bytes <= aux_bytes) return IBVH_OK;
        if (aux) { cudaFree(aux); aux = nullptr; aux_bytes = 0; }
        size_t want = bytes + (bytes >> 2) + (1u << 20);
        cudaError_t e = cudaMalloc((void**)&aux, want);
        if (e != cudaSuccess) { set_cuda_error(e, "cudaMalloc(aux)"); aux = nullptr; cudaGetLastError(); return IBVH_ERR_ALLOC; }
        aux_bytes = want;
        return IBVH_OK;
    }

    // BFS traversal (traverse_bfs.cuh): the two ping-pong lists of BVTT entries. Grow-only, allocated apart from each
    // other because the destination of a level may have to grow while the source is still being read.
    char* bfs_buf[2] = {nullptr, nullptr};
    size_t bfs_bytes[2] = {0, 0};
    int reserve_bfs(int which, size_t bytes) {
        if (bytes <= bfs_bytes[which]) return IBVH_OK;
        if (bfs_buf[which]) { cudaFree(bfs_buf[which]); bfs_buf[which] = nullptr; bfs_bytes[which] = 0; }
        size_t want = bytes + (bytes >> 3) + (1u << 16);
        cudaError_t e = cudaMalloc((void**)&bfs_buf[which], want);
        if (e != cudaSuccess) { set_cuda_error(e, "cudaMalloc(BFS list)"); bfs_buf[which] = nullptr; cudaGetLastError(); return IBVH_ERR_ALLOC; }
        bfs_bytes[which] = want;
        return IBVH_OK;
    }
    void free_bfs() {
        for (int k = 0; k < 2; ++k) { if (bfs_buf[k]) cudaFree(bfs_buf[k]); bfs_buf[k] = nullptr; bfs_bytes[k] = 0; }
        bfs_pending.valid = false;
    }
    // leaf-level list of a BFS call that returned IBVH_ERR_CAPACITY (the repeat call only redoes the leaf level)
    struct BfsPending {
        bool valid = false, fused = false;       // fused: the list is the LAST NODE level's (its leaf level runs fused, bfs_last_kernel)
        int kind = 0, cur = 0;
        const void *leaves1 = nullptr, *leaves2 = nullptr;
        long long n1 = 0, n2 = 0, start1 = 0, start2 = 0, checks = 0;
        unsigned long long count = 0;
    } bfs_pending;

    // Sidecar of a build (round 2): what the pyramid traversal used to re-derive from the tree on EVERY call — the leaf
    // volumes as 16-byte records, the refinement's node levels in 64-byte aligned runs, the three finest levels of the
    // query pyramid — is written once by the build's own gather / merge kernels (they hold the data in registers /
    // shared memory anyway) into library-owned memory, tagged with a build id the caller hands back in ibvh_bvh_t.
    // Two slots (a pair traversal needs the sidecars of two trees); a traversal whose id matches neither packs on the fly.
    struct Sidecar {
        unsigned long long id = 0;             // 0 = empty
        long long n = 0;
        int leaf_kind = 0, float_bytes = 0, built_level = 0, levels = 0;
        char* buf = nullptr;                   // grow-only
        size_t bytes = 0;
        size_t pt_off = 0, nt_off = 0, u_off = 0, idx_off = 0;      // byte offsets of the packed volumes, aligned node levels, query pyramid, leaf indices
        int index_bytes = 0;
        int u_levels = 0;                      // pyramid levels present (<= 3)
    } sidecars[2];
    unsigned long long next_build_id = 1, last_build_id = 0;
    int side_next = 0;
    Sidecar* find_sidecar(unsigned long long id, long long n) {
        if (id == 0) return nullptr;
        for (int k = 0; k < 2; ++k) if (sidecars[k].id == id && sidecars[k].n == n) return &sidecars[k];
        return nullptr;
    }

    // pyramid schedule: learned list-capacity factor (pairs per query group), see traverse_pyramid
    double pyr_factor = 0.0;
    long long stash_hint = 0;     // contact total of the last ordered pyramid traversal (sizes the hit stash of the next one)

    // persistent small device block: [0..47] scene bounds (6 x u64 ordered keys),
    // [64..) counters: total contacts (u64), tile tickets, stats.
    char* d_small = nullptr;
    static constexpr size_t kSmallBytes = 4096;

    // pinned host block for read-backs
    char* h_pinned = nullptr;
    static constexpr size_t kPinnedBytes = 4096;

    int64_t last_stats[4] = {0, 0, 0, 0};
    int64_t last_peer_counts[16] = {0};      // per-rank counts of the last fused multi-GPU traversal (ibvh_peer_last_counts)

    // one outstanding deferred traversal (IBVH_TRAVERSE_DEFER): what ibvh_traverse_finish needs to judge the
    // read-back that the enqueued work leaves in the pinned block
    struct Deferred {
        bool active = false;
        int nl = 0;
        unsigned long long cap[16] = {0};
        long long nqg[16] = {0};
        long long top_pairs = 0;
        long long capacity = 0;
        int fan2 = 64, leaf2 = 16;
    } deferred;
    cudaEvent_t ev_defer = nullptr;

    // side stream + events: the per-query pass over flagged groups overlaps the tile kernel
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // optional per-kernel timing (ibvh_profile_*): CUDA events recorded on the launching stream
    static constexpr int kMaxProf = 512;
    bool prof_on = false;
    int prof_n = 0;
    const char* prof_name[kMaxProf];
    cudaEvent_t prof_ev[kMaxProf][2];
    int prof_created = 0;

    void set_cuda_error(cudaError_t e, const char* what) {
        snprintf(err, sizeof(err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    }
    void set_error(const char* msg) { snprintf(err, sizeof(err), "%s", msg); }

    // Arena: reset() at the start of an entry point, alloc() bump-allocates 256-byte aligned
    // blocks. Growing frees and reallocates (contents are scratch, never live across calls).
    void reset() { ws_off = 0; }
    int reserve(size_t bytes) {
        if (bytes <= ws_bytes) return IBVH_OK;
        if (ws) { cudaError_t e = cudaFree(ws); ws = nullptr; ws_bytes = 0; if (e != cudaSuccess) { set_cuda_error(e, "cudaFree(ws)"); return IBVH_ERR_CUDA; } }
        size_t want = bytes + (bytes >> 3) + (1u << 20);
        cudaError_t e = cudaMalloc((void**)&ws, want);
        if (e != cudaSuccess) { set_cuda_error(e, "cudaMalloc(workspace)"); ws = nullptr; cudaGetLastError(); return IBVH_ERR_ALLOC; }
        ws_bytes = want;
        return IBVH_OK;
    }
    template <class T> T* alloc(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (ws_off + bytes > ws_bytes) return nullptr;
        T* p = (T*)(ws + ws_off);
        ws_off += bytes;
        return p;
    }
    static size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

namespace ibvh {

// offsets inside d_small
constexpr size_t kSmallBounds = 0;        // 6 x u64
constexpr size_t kSmallTotal = 64;        // u64 contact total
constexpr size_t kSmallStats = 128;       // 3 x u64
constexpr size_t kSmallTickets = 256;     // 16 x u32 tile tickets (one per radix pass)
constexpr size_t kSmallBoundsF = 512;     // 6 floats/doubles: padded bounds actually used
constexpr size_t kSmallFixup = 960;       // 2 x u32: lengths of the long-segment lists of the ordered fix-up
constexpr size_t kSmallBfs = 1024;        // u64: entries appended by the current BFS level

// RAII scope that brackets one kernel launch with events when profiling is enabled.
struct ProfScope {
    ibvh_handle* h; cudaStream_t st; int slot = -1;
    ProfScope(ibvh_handle* h_, cudaStream_t st_, const char* name) : h(h_), st(st_) {
        if (!h->prof_on || h->prof_n >= ibvh_handle::kMaxProf) return;
        slot = h->prof_n++;
        if (slot >= h->prof_created) {
            cudaEventCreate(&h->prof_ev[slot][0]);
            cudaEventCreate(&h->prof_ev[slot][1]);
            h->prof_created = slot + 1;
        }
        h->prof_name[slot] = name;
        cudaEventRecord(h->prof_ev[slot][0], st);
    }
    ~ProfScope() { if (slot >= 0) cudaEventRecord(h->prof_ev[slot][1], st); }
};

struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; } if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false; target = dev; }
    ~DeviceGuard() { if (prev >= 0 && prev != target) cudaSetDevice(prev); }
    int target = 0;
};

}  // namespace ibvh
