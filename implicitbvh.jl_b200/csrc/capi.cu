// capi.cu — extern "C" entry points of libibvh_b200.so (declared in include/ibvh.h): argument
// checks mirroring the reference's @argcheck's, type dispatch onto the kernel templates, workspace
// carving and the launch sequences. No CPU fallback: without a device every call fails loudly.
// The file is compiled once per IBVH_PART_* macro (HOST, BUILD, SINGLE, PAIR, RAYS, BFS) so that the template
// instantiations of the six groups of entry points build in parallel (see Makefile).
#if !defined(IBVH_PART_HOST) && !defined(IBVH_PART_BUILD) && !defined(IBVH_PART_SINGLE) && !defined(IBVH_PART_PAIR) && !defined(IBVH_PART_RAYS) && !defined(IBVH_PART_BFS)
#define IBVH_PART_ALL 1
#endif
#include <climits>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <new>
#include <type_traits>

#include "aggregate.cuh"
#include "common.cuh"
#include "handle.cuh"
#include "morton.cuh"
#include "radix_sort.cuh"
#include "reference_shaped.cuh"
#include "traverse.cuh"
#include "traverse_bfs.cuh"
#include "traverse_tile.cuh"
#include "traverse_pyramid.cuh"

using namespace ibvh;

// nvcc derives the "unique" name of an anonymous namespace from the source file name, which is the same for all
// parts: use a per-part named namespace instead
#ifndef IBVH_NS
#define IBVH_NS ns_all
#endif
namespace IBVH_NS {

template <class X> struct Tag { using type = X; };

// ---- type dispatch --------------------------------------------------------------------------------
template <class V, class F> int dispatch_im(const ibvh_types_t& t, F&& f) {
    if (t.index_bytes == 4) {
        if (t.morton_bytes == 2) return f(Tag<Leaf<V, int32_t, uint16_t>>{});
        if (t.morton_bytes == 4) return f(Tag<Leaf<V, int32_t, uint32_t>>{});
        if (t.morton_bytes == 8) return f(Tag<Leaf<V, int32_t, uint64_t>>{});
    } else if (t.index_bytes == 8) {
        if (t.morton_bytes == 2) return f(Tag<Leaf<V, int64_t, uint16_t>>{});
        if (t.morton_bytes == 4) return f(Tag<Leaf<V, int64_t, uint32_t>>{});
        if (t.morton_bytes == 8) return f(Tag<Leaf<V, int64_t, uint64_t>>{});
    }
    return IBVH_ERR_UNSUPPORTED;
}
template <class F> int dispatch_leaf(const ibvh_types_t& t, F&& f) {
    if (t.float_bytes == 4) {
        if (t.leaf_kind == IBVH_BSPHERE) return dispatch_im<BSphere<float>>(t, f);
        if (t.leaf_kind == IBVH_BBOX) return dispatch_im<BBox<float>>(t, f);
    }
#ifdef IBVH_ENABLE_F64
    if (t.float_bytes == 8) {
        if (t.leaf_kind == IBVH_BSPHERE) return dispatch_im<BSphere<double>>(t, f);
        if (t.leaf_kind == IBVH_BBOX) return dispatch_im<BBox<double>>(t, f);
    }
#endif
    return IBVH_ERR_UNSUPPORTED;
}
// node type for a leaf type: BBox nodes for everything; BSphere nodes only over BSphere leaves
// (the reference has no BSphere(::BBox) conversion, merge.jl).
template <class L, class T, class F> int dispatch_node_kind(const ibvh_types_t& t, F&& f) {
    if (t.node_kind == IBVH_BBOX) return f(Tag<BBox<T>>{});
    if (t.node_kind == IBVH_BSPHERE) {
        if constexpr (L::vol_t::kind == IBVH_BSPHERE) return f(Tag<BSphere<T>>{});
        else return IBVH_ERR_ARGUMENT;
    }
    return IBVH_ERR_UNSUPPORTED;
}
// The node float type is the leaf float type unless node_float_bytes says otherwise; the one mixed combination built
// is the reference's default call: Float64 leaves under Float32 nodes (README.md:38-46, build.jl:198-205).
template <class L, class F> int dispatch_node(const ibvh_types_t& t, F&& f) {
    using T = typename L::value_type;
    const int nfb = t.node_float_bytes == 0 ? (int)sizeof(T) : t.node_float_bytes;
    if (nfb == (int)sizeof(T)) return dispatch_node_kind<L, T>(t, f);
#ifdef IBVH_ENABLE_F64
    if constexpr (sizeof(T) == 8) {
        if (nfb == 4) return dispatch_node_kind<L, float>(t, f);
    }
#endif
    return IBVH_ERR_UNSUPPORTED;
}

bool types_ok(const ibvh_types_t* t) {
    if (!t) return false;
    if (t->leaf_kind != IBVH_BSPHERE && t->leaf_kind != IBVH_BBOX) return false;
    if (t->node_kind != IBVH_BSPHERE && t->node_kind != IBVH_BBOX) return false;
    if (t->float_bytes != 4 && t->float_bytes != 8) return false;
    if (t->node_float_bytes != 0 && t->node_float_bytes != 4 && t->node_float_bytes != 8) return false;
    if (t->index_bytes != 4 && t->index_bytes != 8) return false;
    if (t->morton_bytes != 2 && t->morton_bytes != 4 && t->morton_bytes != 8) return false;
    return true;
}

int grid_for(int64_t n, int threads, int per_thread, int cap) {
    int64_t g = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

#define IBVH_LAUNCH_CHECK(h, what)                                                   \
    do {                                                                             \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) { (h)->set_cuda_error(_e, what); return IBVH_ERR_CUDA; } \
    } while (0)

// ---- radix sort driver ---------------------------------------------------------------------------------
// keysA holds the input keys; hist holds the raw per-pass histograms. On return the sorted keys / the
// permutation are in *keys_out / *vals_out (one of the ping-pong buffers). The look-back words are 32-bit
// (2 flag bits + a 30-bit prefix) below 2^30 leaves and 64-bit from there to the 2^31 that levels <= 32 allows.
inline bool sort_wide_lookback(int64_t n) { return n >= (int64_t(1) << 30); }

template <class M>
int sort_pairs(ibvh_handle* h, M* keysA, M* keysB, uint32_t* valsA, uint32_t* valsB, int64_t n, uint32_t* hist,
               void* lookback, uint32_t* tickets, cudaStream_t st, M** keys_out, uint32_t** vals_out) {
    auto scope = [&](const char* name) { return ProfScope(h, st, name); };
    cudaError_t e;
    if (sort_wide_lookback(n) || h->cfg.force_wide_lookback)
        e = sort_pairs_impl<M, unsigned long long, sort_threads<M>(), sort_items<M>(), sort_minb<M>()>(keysA, keysB, valsA, valsB, n, hist, lookback, tickets, st, keys_out, vals_out, scope);
    else
        e = sort_pairs_impl<M, uint32_t, sort_threads<M>(), sort_items<M>(), sort_minb<M>()>(keysA, keysB, valsA, valsB, n, hist, lookback, tickets, st, keys_out, vals_out, scope);
    if (e != cudaSuccess) { h->set_cuda_error(e, "onesweep_kernel"); return IBVH_ERR_CUDA; }
    return IBVH_OK;
}

constexpr int kMergeTile = 1024;
constexpr int kMergeThreads = 256;
constexpr int kMergeLevels = 10;   // log2(kMergeTile)

// ---- pyramid refinement driver (BBox nodes) -----------------------------------------------------------------
struct PyrPlan { int n; PyrLevel lv[kPyrMaxLevels]; int64_t u_total; int64_t t_total; };

// levels from the finest (k = 2) to the top; false if the schedule does not apply (tree too short, the needed
// tree levels are not built, or the coarsest usable level still has too many groups for an all-pairs start)
inline bool make_pyr_plan(const TreeInfo& ti, int64_t built_level, int64_t q_begin, int64_t q_count, PyrPlan* plan) {
    const int L = ti.levels;
    const int64_t q_end = q_begin + q_count;
    plan->n = 0; plan->u_total = 0; plan->t_total = 0;
    for (int k = kPyrLeafLog; plan->n < kPyrMaxLevels; k += kPyrFan) {
        const int tl = L - k;
        if (tl < 1 || tl < built_level) break;
        PyrLevel& v = plan->lv[plan->n];
        v.k = k; v.tree_level = tl; v.ntg = ti.level_nreal[tl]; v.tnode0 = ti.level_start[tl];
        // (the query-pyramid levels carry 2^fan boxes of padding on both sides: the TMA refine kernel copies the aligned run of
        // 2^fan children of a group even when the shard range cuts it)
        v.qg_first = q_begin >> k; v.nqg = ((q_end - 1) >> k) - v.qg_first + 1; v.u_off = plan->u_total + (int64_t(1) << kPyrFan);
        plan->u_total += v.nqg + (int64_t(2) << kPyrFan);
        v.t_off = plan->t_total;
        plan->t_total += (v.ntg + 15) & ~int64_t(7);               // whole groups of 8 + slack, 64-byte aligned starts
        plan->n += 1;
        if (v.nqg <= 2048 && v.ntg <= 2048) return true;          // good top level
    }
    if (plan->n == 0) return false;
    const PyrLevel& top = plan->lv[plan->n - 1];
    return top.nqg * top.ntg <= (int64_t(1) << 26);                 // an all-pairs start of <= 64 M box tests is still cheap
}


// ---- sidecar of a build (handle.cuh: Sidecar): layout + device pointers --------------------------------------------
// [packed volumes: n + 8 records][aligned node levels: full-range plan's t_total boxes][query pyramid: levels k = 2, 5, 8]
template <class L, class N> struct SidecarLayout {
    using V = typename L::vol_t;
    using T = typename N::value_type;
    PyrPlan plan;                 // the full-range plan (q = [0, n))
    size_t pt_off, nt_off, u_off, idx_off, bytes;
    int u_levels;
    int64_t u_level_off[3], u_level_n[3];
    bool ok;
};
template <class L, class N> SidecarLayout<L, N> sidecar_layout(const TreeInfo& ti, int64_t built_level) {
    SidecarLayout<L, N> s{};
    s.ok = false;
    if constexpr (!std::is_same<N, BBox<typename L::value_type>>::value) return s;     // pyramid schedule: BBox nodes of the leaf float type
    if (ti.n >= (int64_t(1) << 29)) return s;
    if (!make_pyr_plan(ti, built_level, 0, ti.n, &s.plan)) return s;
    size_t off = 0;
    s.pt_off = off; off += ibvh_handle::padded((size_t)(ti.n + 8) * sizeof(Packed<typename L::vol_t>));
    s.nt_off = off; off += ibvh_handle::padded((size_t)s.plan.t_total * sizeof(N));
    s.u_off = off;
    s.u_levels = s.plan.n < 3 ? s.plan.n : 3;
    int64_t uo = 0;
    for (int l = 0; l < s.u_levels; ++l) {
        // whole merge tiles worth of groups (the kernel writes every group of a tile) + the 2^fan padding either side
        const int64_t groups = ((ti.n + kMergeTile - 1) / kMergeTile) * (kMergeTile >> s.plan.lv[l].k);
        s.u_level_off[l] = uo + (int64_t(1) << kPyrFan);
        s.u_level_n[l] = groups;
        uo += groups + (int64_t(2) << kPyrFan);
    }
    off += ibvh_handle::padded((size_t)uo * sizeof(UBox<typename N::value_type>));
    s.idx_off = off; off += ibvh_handle::padded((size_t)ti.n * sizeof(typename L::idx_t));
    s.bytes = off;
    s.ok = true;
    return s;
}

template <class L, class N> size_t gather_smem_bytes() {
    return ((sizeof(L) * kMergeTile + 15) & ~size_t(15)) + sizeof(N) * (kMergeTile / 2 + kMergeTile / 4);
}
template <class N> size_t merge_smem_bytes() { return sizeof(N) * (kMergeTile + kMergeTile / 2 + kMergeTile / 4); }

// upper levels after the tile kernel produced levels-1 .. levels-kMergeLevels
template <class N>
int merge_upper_levels(ibvh_handle* h, N* nodes, const TreeInfo& ti, int stop_level, cudaStream_t st, const SideLevels<N>* side = nullptr) {
    int src = ti.levels - kMergeLevels;
    {   // function attributes are per device: set on every call (cheap)
        cudaError_t e = cudaFuncSetAttribute(merge_levels_kernel<N, kMergeTile, kMergeThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem_bytes<N>());
        if (e != cudaSuccess) { h->set_cuda_error(e, "cudaFuncSetAttribute(merge_levels)"); return IBVH_ERR_CUDA; }
    }
    while (src > stop_level) {
        int stop = src - kMergeLevels > stop_level ? src - kMergeLevels : stop_level;
        int64_t tiles = (ti.level_nreal[src] + kMergeTile - 1) / kMergeTile;
        { ProfScope _ps(h, st, "merge_levels_kernel");
        merge_levels_kernel<N, kMergeTile, kMergeThreads><<<(unsigned)tiles, kMergeThreads, merge_smem_bytes<N>(), st>>>(nodes, ti, src, stop, side ? *side : SideLevels<N>{});
        }
        IBVH_LAUNCH_CHECK(h, "merge_levels_kernel");
        src -= kMergeLevels;
    }
    return IBVH_OK;
}

template <class L, class SRC, class N, bool GATHER>
int launch_gather_merge(ibvh_handle* h, const SRC* src, const uint32_t* perm, const typename L::mor_t* keys_sorted, L* leaves, N* nodes,
                        const TreeInfo& ti, int stop_level, cudaStream_t st, const SideOut<typename L::vol_t, N>* so = nullptr) {
    auto kern = gather_merge_kernel<L, SRC, N, kMergeTile, kMergeThreads, GATHER>;
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gather_smem_bytes<L, N>());
        if (e != cudaSuccess) { h->set_cuda_error(e, "cudaFuncSetAttribute(gather_merge)"); return IBVH_ERR_CUDA; }
    }
    int64_t tiles = (ti.n + kMergeTile - 1) / kMergeTile;
    { ProfScope _ps(h, st, "gather_merge_kernel");
    const uintptr_t va = (uintptr_t)src;
    const int vec = std::is_same<L, SRC>::value ? 8 : (va % 16 == 0 ? 16 : (va % 8 == 0 ? 8 : 4));
    kern<<<(unsigned)tiles, kMergeThreads, gather_smem_bytes<L, N>(), st>>>(src, perm, keys_sorted, leaves, nodes, ti, stop_level, vec, so ? *so : SideOut<typename L::vol_t, N>{});
    }
    IBVH_LAUNCH_CHECK(h, "gather_merge_kernel");
    return IBVH_OK;
}

struct SortScratch {
    void* keysA; void* keysB; uint32_t* valsA; uint32_t* valsB; uint32_t* hist; void* lookback; uint32_t* tickets; void* copy;
};

template <class L> size_t build_workspace_bytes(int64_t n, bool need_copy) {
    using M = typename L::mor_t;
    constexpr int P = radix_passes<M>();
    const int64_t tiles = (n + sort_tile<M>() - 1) / sort_tile<M>();
    size_t b = 0;
    b += 2 * ibvh_handle::padded((size_t)n * sizeof(M));
    b += 2 * ibvh_handle::padded((size_t)n * 4);
    b += ibvh_handle::padded((size_t)P * radix_bins<M>() * 4);
    b += ibvh_handle::padded((size_t)P * tiles * radix_bins<M>() * 8);       // (8-byte look-back words from 2^30 leaves; sized for them always)
    if (need_copy) b += ibvh_handle::padded((size_t)n * sizeof(L));
    return b + 4096;
}

template <class L> int carve_sort_scratch(ibvh_handle* h, int64_t n, bool need_copy, SortScratch* s, cudaStream_t st) {
    using M = typename L::mor_t;
    constexpr int P = radix_passes<M>();
    if (n > (int64_t(1) << 31)) { h->set_error("n > 2^31 leaves: the implicit tree would need more than 32 levels"); return IBVH_ERR_UNSUPPORTED; }
    const size_t lb_bytes = (sort_wide_lookback(n) || h->cfg.force_wide_lookback) ? 8 : 4;
    const int64_t tiles = (n + sort_tile<M>() - 1) / sort_tile<M>();
    int rc = h->reserve(build_workspace_bytes<L>(n, need_copy));
    if (rc != IBVH_OK) return rc;
    h->reset();
    s->keysA = h->alloc<M>(n); s->keysB = h->alloc<M>(n);
    s->valsA = h->alloc<uint32_t>(n); s->valsB = h->alloc<uint32_t>(n);
    s->hist = h->alloc<uint32_t>((size_t)P * radix_bins<M>());
    s->lookback = h->alloc<unsigned char>((size_t)P * tiles * radix_bins<M>() * lb_bytes);
    s->copy = need_copy ? (void*)h->alloc<L>(n) : nullptr;
    s->tickets = (uint32_t*)(h->d_small + kSmallTickets);
    if (!s->keysA || !s->keysB || !s->valsA || !s->valsB || !s->hist || !s->lookback || (need_copy && !s->copy)) { h->set_error("workspace carve failed"); return IBVH_ERR_ALLOC; }
    IBVH_CUDA_TRY(h, cudaMemsetAsync(s->lookback, 0, (size_t)P * tiles * radix_bins<M>() * lb_bytes, st));
    return IBVH_OK;
}

// user-supplied bounds -> device (6 T's) through the pinned block
template <class T> int upload_user_bounds(ibvh_handle* h, const double* mins, const double* maxs, cudaStream_t st, T** d_out) {
    if (!mins || !maxs) return IBVH_ERR_ARGUMENT;
    // the pinned block may still be in flight from a previous call on this stream: wait for it
    IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
    T* hp = (T*)(h->h_pinned + 1024);
    for (int k = 0; k < 3; ++k) { hp[k] = (T)mins[k]; hp[3 + k] = (T)maxs[k]; }
    T* d = (T*)(h->d_small + kSmallBoundsF + 128);
    IBVH_CUDA_TRY(h, cudaMemcpyAsync(d, hp, 6 * sizeof(T), cudaMemcpyHostToDevice, st));
    *d_out = d;
    return IBVH_OK;
}

template <class T> int read_used_bounds(ibvh_handle* h, double* out_mins, double* out_maxs, cudaStream_t st) {
    if (!out_mins && !out_maxs) return IBVH_OK;
    T* hp = (T*)(h->h_pinned + 2048);
    IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, h->d_small + kSmallBoundsF, 6 * sizeof(T), cudaMemcpyDeviceToHost, st));
    IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) { if (out_mins) out_mins[k] = (double)hp[k]; if (out_maxs) out_maxs[k] = (double)hp[3 + k]; }
    return IBVH_OK;
}

// bounds + encode. SRC == L (in place, optional copy) or SRC == volume (wrap path)
template <class L, class SRC, bool COPY>
int launch_encode(ibvh_handle* h, const SRC* src, int64_t n, int compute_extrema, const double* mins, const double* maxs,
                  typename L::mor_t* keys, L* copy, uint32_t* hist, uint32_t* tickets, cudaStream_t st) {
    using T = typename L::value_type;
    using M = typename L::mor_t;
    using Ord = typename OrdOf<T>::type;
    constexpr int P = radix_passes<M>();
    Ord* bounds = (Ord*)(h->d_small + kSmallBounds);
    T* used = (T*)(h->d_small + kSmallBoundsF);
    T* user = nullptr;
    if (!compute_extrema) { int rc = upload_user_bounds<T>(h, mins, maxs, st, &user); if (rc != IBVH_OK) return rc; }
    { ProfScope _ps(h, st, "init_build_kernel");
    init_build_kernel<T><<<8, 256, 0, st>>>(bounds, hist, P * radix_bins<M>(), tickets, 16);
    }
    IBVH_LAUNCH_CHECK(h, "init_build_kernel");
    const int grid = grid_for(n, 256, 4, h->sm_count * 8);
    const uintptr_t va = (uintptr_t)src;
    const int vec = va % 16 == 0 ? 16 : (va % 8 == 0 ? 8 : 4);          // widest load the array's alignment allows
    if (compute_extrema) {
        { ProfScope _ps(h, st, "bounds_kernel");
        bounds_kernel<SRC, T><<<grid, 256, 0, st>>>(src, n, bounds, vec);
        }
        IBVH_LAUNCH_CHECK(h, "bounds_kernel");
    }
    { ProfScope _ps(h, st, "encode_kernel");
    encode_kernel<SRC, L, COPY><<<grid, 256, 0, st>>>(src, n, compute_extrema ? bounds : nullptr, user, used, keys, copy, hist, vec);
    }
    IBVH_LAUNCH_CHECK(h, "encode_kernel");
    return IBVH_OK;
}

// ---- BVH(...) pipeline -----------------------------------------------------------------------------------
template <class L, class N>
int build_impl(ibvh_handle* h, const void* d_volumes, void* d_leaves, int64_t n, void* d_nodes, int64_t built_level,
               int compute_extrema, const double* mins, const double* maxs, cudaStream_t st) {
    using M = typename L::mor_t;
    using V = typename L::vol_t;
    ibvh_tree_t tree;
    int rc = make_tree(n, &tree);
    if (rc != IBVH_OK) return rc;
    if (built_level < 1 || built_level > tree.levels) return IBVH_ERR_ARGUMENT;      // build.jl:314
    const bool wrap = d_volumes != nullptr;
    SortScratch s;
    rc = carve_sort_scratch<L>(h, n, !wrap, &s, st);
    if (rc != IBVH_OK) return rc;
    if (wrap) rc = launch_encode<L, V, false>(h, (const V*)d_volumes, n, compute_extrema, mins, maxs, (M*)s.keysA, nullptr, s.hist, s.tickets, st);
    else rc = launch_encode<L, L, true>(h, (const L*)d_leaves, n, compute_extrema, mins, maxs, (M*)s.keysA, (L*)s.copy, s.hist, s.tickets, st);
    if (rc != IBVH_OK) return rc;
    M* keys_sorted; uint32_t* perm;
    rc = sort_pairs<M>(h, (M*)s.keysA, (M*)s.keysB, s.valsA, s.valsB, n, s.hist, s.lookback, s.tickets, st, &keys_sorted, &perm);
    if (rc != IBVH_OK) return rc;
    TreeInfo ti = make_tree_info(tree);
    // the level above the leaves is always produced (build.jl:369), further levels down to built_level
    int stop_level = (int)(built_level < tree.levels - 1 ? built_level : tree.levels - 1);
    if (stop_level < 1) stop_level = 1;
    // sidecar of this build: packed volumes, aligned node levels, finest query-pyramid levels (handle.cuh)
    h->last_build_id = 0;
    SideOut<V, N> so{};
    SideLevels<N> sl{};
    typename L::idx_t* side_idx = nullptr;
    ibvh_handle::Sidecar* sc = nullptr;
    if (!h->cfg.no_sidecar && !h->cfg.fused_gather && tree.real_nodes >= 2) {
        SidecarLayout<L, N> lay = sidecar_layout<L, N>(ti, built_level);
        if (lay.ok) {
            sc = &h->sidecars[h->side_next];
            h->side_next ^= 1;
            sc->id = 0;
            if (lay.bytes > sc->bytes) {
                if (sc->buf) { cudaFree(sc->buf); sc->buf = nullptr; sc->bytes = 0; }
                const size_t want = lay.bytes + (lay.bytes >> 4) + (1u << 16);
                if (cudaMalloc((void**)&sc->buf, want) != cudaSuccess) { cudaGetLastError(); sc = nullptr; }      // no sidecar: the traversal packs on the fly
                else sc->bytes = want;
            }
        }
        if (sc) {
            sc->n = n; sc->leaf_kind = V::kind; sc->float_bytes = (int)sizeof(typename L::value_type); sc->built_level = (int)built_level; sc->levels = (int)tree.levels;
            sc->pt_off = lay.pt_off; sc->nt_off = lay.nt_off; sc->u_off = lay.u_off; sc->u_levels = lay.u_levels;
            sc->idx_off = lay.idx_off; sc->index_bytes = (int)sizeof(typename L::idx_t);
            side_idx = (typename L::idx_t*)(sc->buf + lay.idx_off);
            so.pt = (Packed<V>*)(sc->buf + lay.pt_off);
            for (int l = 0; l + 1 < lay.plan.n; ++l) {               // (the top level is read straight from the caller's nodes)
                const int tl = lay.plan.lv[l].tree_level;
                so.lvl[tl] = (N*)(sc->buf + lay.nt_off) + lay.plan.lv[l].t_off;
                sl.lvl[tl] = so.lvl[tl];
            }
            if constexpr (std::is_same<N, BBox<typename L::value_type>>::value) {
                for (int l = 0; l < lay.u_levels; ++l) so.u[l] = (UBox<typename N::value_type>*)(sc->buf + lay.u_off) + lay.u_level_off[l];
            }
        }
    }
    if (h->cfg.fused_gather) {
        if (wrap) rc = launch_gather_merge<L, V, N, true>(h, (const V*)d_volumes, perm, keys_sorted, (L*)d_leaves, (N*)d_nodes, ti, stop_level, st);
        else rc = launch_gather_merge<L, L, N, true>(h, (const L*)s.copy, perm, keys_sorted, (L*)d_leaves, (N*)d_nodes, ti, stop_level, st);
    } else {
        // gather (random reads, full occupancy) then tile merge from the sorted leaves (coalesced)
        const unsigned gb = (unsigned)((n + 1023) / 1024);
        { ProfScope _ps(h, st, "gather_kernel");
        const uintptr_t va = (uintptr_t)d_volumes;
        const int vec = va % 16 == 0 ? 16 : (va % 8 == 0 ? 8 : 4);
        if (wrap) gather_kernel<L, V><<<gb, 256, 0, st>>>((const V*)d_volumes, perm, keys_sorted, (L*)d_leaves, n, vec, so.pt, side_idx);
        else gather_kernel<L, L><<<gb, 256, 0, st>>>((const L*)s.copy, perm, keys_sorted, (L*)d_leaves, n, 8, so.pt, side_idx);
        }
        IBVH_LAUNCH_CHECK(h, "gather_kernel");
        rc = launch_gather_merge<L, L, N, false>(h, (const L*)nullptr, nullptr, nullptr, (L*)d_leaves, (N*)d_nodes, ti, stop_level, st, &so);
    }
    if (rc != IBVH_OK) return rc;
    if (tree.real_nodes >= 2) rc = merge_upper_levels<N>(h, (N*)d_nodes, ti, stop_level, st, &sl);
    if (rc == IBVH_OK && sc) { sc->id = h->next_build_id++; h->last_build_id = sc->id; }
    return rc;
}

// ---- traversal driver ---------------------------------------------------------------------------------------
template <class I>
int scan_counts(ibvh_handle* h, I* counts, int64_t n, long long* block_sums, unsigned long long* d_total, cudaStream_t st) {
    const int64_t blocks = (n + kScanTile - 1) / kScanTile;
    { ProfScope _ps(h, st, "scan_reduce_kernel");
    scan_reduce_kernel<I><<<(unsigned)blocks, kScanThreads, 0, st>>>(counts, n, block_sums);
    }
    IBVH_LAUNCH_CHECK(h, "scan_reduce_kernel");
    { ProfScope _ps(h, st, "scan_block_sums_kernel");
    scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, blocks, d_total);
    }
    IBVH_LAUNCH_CHECK(h, "scan_block_sums_kernel");
    { ProfScope _ps(h, st, "scan_apply_kernel");
    scan_apply_kernel<I><<<(unsigned)blocks, kScanThreads, 0, st>>>(counts, n, block_sums);
    }
    IBVH_LAUNCH_CHECK(h, "scan_apply_kernel");
    return IBVH_OK;
}

int read_total(ibvh_handle* h, unsigned long long* d_total, int64_t* out, cudaStream_t st) {
    unsigned long long* hp = (unsigned long long*)h->h_pinned;
    IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, d_total, 8, cudaMemcpyDeviceToHost, st));
    IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
    *out = (int64_t)*hp;
    return IBVH_OK;
}

// stats variants are compiled only for the default type set (BSphere{Float32} / Int32 / UInt32 / BBox)
template <class LT, class N> constexpr bool stats_combo() {
    return std::is_same<LT, Leaf<BSphere<float>, int32_t, uint32_t>>::value && std::is_same<N, BBox<float>>::value;
}

template <int KIND, int MODE, bool PACKET, class LQ, class LT, class N, class I>
int launch_traverse(ibvh_handle* h, const LQ* qleaves, const typename LT::value_type* points, const typename LT::value_type* dirs,
                    const DBvh<LT, N>& bvh, const TraverseArgs& a, I* counts, IndexPair<I>* contacts, cudaStream_t st) {
    if (a.q_count <= 0) return IBVH_OK;
    if constexpr (stats_combo<LT, N>() && MODE == kCount) {
        if (a.stats) {
            if constexpr (PACKET) {
                const int64_t warps = (a.q_count + 31) / 32;
                const int64_t blocks = (warps + kPacketWarps - 1) / kPacketWarps;
                lvt_packet_kernel<KIND, MODE, LQ, LT, N, I, true><<<(unsigned)blocks, kPacketWarps * 32, 0, st>>>(qleaves, bvh, a, counts, contacts);
            } else {
                const int64_t blocks = (a.q_count + 127) / 128;
                lvt_thread_kernel<KIND, MODE, LQ, LT, N, I, true><<<(unsigned)blocks, 128, 0, st>>>(qleaves, points, dirs, bvh, a, counts, contacts);
            }
            IBVH_LAUNCH_CHECK(h, "lvt stats kernel");
            return IBVH_OK;
        }
    }
    if constexpr (PACKET) {
        const int64_t warps = (a.q_count + 31) / 32;
        const int64_t blocks = (warps + kPacketWarps - 1) / kPacketWarps;
        { ProfScope _ps(h, st, "lvt_packet_kernel");
        lvt_packet_kernel<KIND, MODE, LQ, LT, N, I><<<(unsigned)blocks, kPacketWarps * 32, 0, st>>>(qleaves, bvh, a, counts, contacts);
        }
        IBVH_LAUNCH_CHECK(h, "lvt_packet_kernel");
    } else {
        const int64_t blocks = (a.q_count + 127) / 128;
        if constexpr (KIND == kRays) {
            // rays: stackless two-children schedule (IBVH_RAYS_REFERENCE_SHAPED=1 keeps the reference-shaped proxy)
            if (!h->cfg.rays_reference_shaped) {
                if (h->cfg.rays_static) {
                    { ProfScope _ps(h, st, "rays_kernel");
                    rays_kernel<MODE, LT, N, I><<<(unsigned)blocks, 128, 0, st>>>(points, dirs, bvh, a, counts, contacts);
                    }
                } else {
                    // [ticket | long-ray queue length | ticket of rays_wide_kernel]
                    unsigned long long* ticket = (unsigned long long*)(h->d_small + kSmallTotal + 24);
                    IBVH_CUDA_TRY(h, cudaMemsetAsync(ticket, 0, 24, st));
                    const unsigned pblocks = (unsigned)std::min<int64_t>(blocks, (int64_t)h->sm_count * 12);
                    // long rays are exported to a queue and finished one warp per ray (traverse.cuh)
                    RayWideQueue wq{nullptr, ticket + 1, 0u, 0u};
                    if (a.wq_data) { wq.data = (RayWideEntry*)a.wq_data; wq.cap = a.wq_cap; wq.after = a.wq_after; }
                    constexpr int kHB = MODE == kAtomic ? 512 : 128;                 // hit buffer of the fused multi-GPU variant
                    { ProfScope _ps(h, st, "rays_persistent_kernel");
                    if (a.peer && MODE == kAtomic) rays_persistent_kernel<MODE, LT, N, I, kHB><<<pblocks, 128, 0, st>>>(points, dirs, bvh, a, counts, contacts, ticket, wq);
                    else rays_persistent_kernel<MODE, LT, N, I><<<pblocks, 128, 0, st>>>(points, dirs, bvh, a, counts, contacts, ticket, wq);
                    }
                    if (wq.data) {
                        IBVH_LAUNCH_CHECK(h, "rays_persistent_kernel");
                        const unsigned wblocks = (unsigned)std::min<int64_t>((int64_t)wq.cap / 4 + 1, (int64_t)h->sm_count * 8);
                        ProfScope _ps(h, st, "rays_wide_kernel");
                        if (a.peer && MODE == kAtomic) rays_wide_kernel<MODE, LT, N, I, kHB><<<wblocks, 128, 0, st>>>(points, dirs, bvh, a, counts, contacts, wq, ticket + 2);
                        else rays_wide_kernel<MODE, LT, N, I><<<wblocks, 128, 0, st>>>(points, dirs, bvh, a, counts, contacts, wq, ticket + 2);
                    }
                }
                IBVH_LAUNCH_CHECK(h, "rays_kernel");
                return IBVH_OK;
            }
        }
        { ProfScope _ps(h, st, "lvt_thread_kernel");
        lvt_thread_kernel<KIND, MODE, LQ, LT, N, I><<<(unsigned)blocks, 128, 0, st>>>(qleaves, points, dirs, bvh, a, counts, contacts);
        }
        IBVH_LAUNCH_CHECK(h, "lvt_thread_kernel");
    }
    return IBVH_OK;
}

// rays: capacity of the long-ray queue (0 = long rays stay with their lanes) — see rays_wide_kernel
inline int64_t ray_queue_cap(const ibvh_handle* h, const TraverseArgs& a, int levels) {
    if (h->cfg.rays_wide <= 0 || a.start_level >= levels || a.q_count <= 0) return 0;
    return std::min<int64_t>(int64_t(1) << 22, std::max<int64_t>(4096, a.q_count / h->cfg.rays_wide_div));
}
inline size_t ray_queue_bytes(int64_t cap) { return cap > 0 ? ibvh_handle::padded((size_t)cap * sizeof(RayWideEntry)) : 0; }
// carve the queue from the (already reserved and reset) workspace
inline void ray_queue_carve(ibvh_handle* h, TraverseArgs* a, int64_t cap) {
    a->wq_data = nullptr; a->wq_cap = 0; a->wq_after = 0;
    if (cap <= 0) return;
    a->wq_data = h->alloc<RayWideEntry>((size_t)cap);
    if (a->wq_data) { a->wq_cap = (uint32_t)cap; a->wq_after = (uint32_t)h->cfg.rays_wide; }
}

template <int KIND, bool PACKET, class LQ, class LT, class N, class I>
int traverse_impl(ibvh_handle* h, const LQ* qleaves, const typename LT::value_type* points, const typename LT::value_type* dirs,
                  const DBvh<LT, N>& bvh, TraverseArgs a, uint32_t flags, void* d_counts, void* d_contacts, int64_t capacity,
                  int64_t* num_contacts, cudaStream_t st) {
    unsigned long long* d_total = (unsigned long long*)(h->d_small + kSmallTotal);
    a.total = d_total;
    a.stats = nullptr;
    const bool want_stats = (flags & IBVH_TRAVERSE_STATS) != 0;
    unsigned long long* d_stats = (unsigned long long*)(h->d_small + kSmallStats);
    if (want_stats) { IBVH_CUDA_TRY(h, cudaMemsetAsync(d_stats, 0, 32, st)); a.stats = d_stats; }
    a.capacity = d_contacts ? capacity : 0;
    *num_contacts = 0;
    if (a.q_count <= 0 && !(KIND == kRays && a.peer)) return IBVH_OK;      // (a fused call is collective: an empty shard still joins)
    const bool unordered = (flags & IBVH_TRAVERSE_UNORDERED) != 0 && d_contacts != nullptr;
    int rc;
    const int64_t wq_cap = KIND == kRays ? ray_queue_cap(h, a, bvh.ti.levels) : 0;
    if constexpr (KIND == kRays) {
        if (a.peer) {
            // fused ray traversal + all-gather of the hits (see traverse_pyramid for the contact version)
            if (!(flags & IBVH_TRAVERSE_UNORDERED) || !peer_ok(a.peer) || !a.peer->multicast || a.peer->fused_seq == 0 ||
                h->cfg.rays_reference_shaped || h->cfg.rays_static) {
                h->set_error("fused multi-GPU ray traversal needs IBVH_TRAVERSE_UNORDERED, the persistent schedule and a multicast alias");
                return IBVH_ERR_UNSUPPORTED;
            }
            PeerArgs pa = make_peer_args(a.peer);
            long long region[IBVH_MAX_PEERS + 1];
            peer_regions(a.peer, (int)sizeof(IndexPair<I>), region);
            a.capacity = region[pa.rank + 1] - region[pa.rank];
            int64_t* h_peer = (int64_t*)(h->h_pinned + 3072);
            h_peer[1] = 0;
            IBVH_CUDA_TRY(h, cudaMemsetAsync(d_total, 0, 8, st));       // local slot counter
            rc = h->reserve(ray_queue_bytes(wq_cap) + 4096);
            if (rc != IBVH_OK) return rc;
            h->reset();
            ray_queue_carve(h, &a, wq_cap);
            { ProfScope _ps(h, st, "peer_fused_begin_kernel");
            peer_fused_begin_kernel<<<1, 32, 0, st>>>(pa);
            }
            IBVH_LAUNCH_CHECK(h, "peer_fused_begin_kernel");
            { ProfScope _ps(h, st, "peer_fused_wait_kernel");
            peer_fused_wait_kernel<<<1, 32, 0, st>>>(pa, h_peer);
            }
            IBVH_LAUNCH_CHECK(h, "peer_fused_wait_kernel");
            rc = a.q_count > 0 ? launch_traverse<KIND, kAtomic, PACKET, LQ, LT, N, I>(h, qleaves, points, dirs, bvh, a, (I*)nullptr, (IndexPair<I>*)(pa.mc + pa.header_bytes) + region[pa.rank], st) : IBVH_OK;
            if (rc != IBVH_OK) return rc;
            { ProfScope _ps(h, st, "peer_fused_finish_kernel");
            peer_fused_finish_kernel<<<1, 32, 0, st>>>(pa, d_total, h_peer);
            }
            IBVH_LAUNCH_CHECK(h, "peer_fused_finish_kernel");
            IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
            if (h_peer[1] != 0) { h->set_error("fused ray traversal: a peer did not arrive within 10 s"); return IBVH_ERR_PEER; }
            *num_contacts = h_peer[0];
            long long counts_r[IBVH_MAX_PEERS];
            bool over = false;
            for (int r = 0; r < pa.world; ++r) { counts_r[r] = h_peer[2 + r]; h->last_peer_counts[r] = counts_r[r]; if (counts_r[r] > region[r + 1] - region[r]) over = true; }
            if (over) return IBVH_ERR_CAPACITY;
            PeerMoves mv = make_compact_moves(pa.world, region, counts_r, (int)sizeof(IndexPair<I>) / 8);
            if (mv.n > 0) {
                { ProfScope _ps(h, st, "peer_compact_kernel");
                peer_compact_kernel<<<h->sm_count * 2, 256, 0, st>>>((uint64_t*)(pa.buf[pa.rank] + pa.header_bytes), mv);
                }
                IBVH_LAUNCH_CHECK(h, "peer_compact_kernel");
            }
            return IBVH_OK;
        }
    }
    if (unordered) {
        IBVH_CUDA_TRY(h, cudaMemsetAsync(d_total, 0, 8, st));
        if (wq_cap > 0) {
            rc = h->reserve(ray_queue_bytes(wq_cap) + 4096);
            if (rc != IBVH_OK) return rc;
            h->reset();
            ray_queue_carve(h, &a, wq_cap);
        }
        rc = launch_traverse<KIND, kAtomic, PACKET, LQ, LT, N, I>(h, qleaves, points, dirs, bvh, a, (I*)nullptr, (IndexPair<I>*)d_contacts, st);
        if (rc != IBVH_OK) return rc;
        rc = read_total(h, d_total, num_contacts, st);
        if (rc != IBVH_OK) return rc;
        return *num_contacts > capacity ? IBVH_ERR_CAPACITY : IBVH_OK;
    }
    // ordered: count -> inclusive scan -> (read total) -> write
    const int64_t blocks = (a.q_count + kScanTile - 1) / kScanTile;
    size_t need = ibvh_handle::padded((size_t)blocks * 8) + (d_counts ? 0 : ibvh_handle::padded((size_t)a.q_count * sizeof(I))) + ray_queue_bytes(wq_cap) + 4096;
    rc = h->reserve(need);
    if (rc != IBVH_OK) return rc;
    h->reset();
    long long* block_sums = h->alloc<long long>(blocks);
    I* counts = d_counts ? (I*)d_counts : h->alloc<I>(a.q_count);
    if (!block_sums || !counts) { h->set_error("workspace carve failed"); return IBVH_ERR_ALLOC; }
    ray_queue_carve(h, &a, wq_cap);                       // (used by the count pass and by the write pass)
    if ((flags & IBVH_TRAVERSE_COUNTS_VALID) && d_counts && d_contacts) {
        // second pass only (traverse_single.jl:75): the total is the last entry of the scan
        I* hp = (I*)(h->h_pinned + 64);
        IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, counts + (a.q_count - 1), sizeof(I), cudaMemcpyDeviceToHost, st));
        IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
        *num_contacts = (int64_t)*hp;
        if (*num_contacts == 0) return IBVH_OK;
        if (*num_contacts > capacity) return IBVH_ERR_CAPACITY;
        return launch_traverse<KIND, kWrite, PACKET, LQ, LT, N, I>(h, qleaves, points, dirs, bvh, a, counts, (IndexPair<I>*)d_contacts, st);
    }
    rc = launch_traverse<KIND, kCount, PACKET, LQ, LT, N, I>(h, qleaves, points, dirs, bvh, a, counts, (IndexPair<I>*)nullptr, st);
    if (rc != IBVH_OK) return rc;
    rc = scan_counts<I>(h, counts, a.q_count, block_sums, d_total, st);
    if (rc != IBVH_OK) return rc;
    rc = read_total(h, d_total, num_contacts, st);
    if (rc != IBVH_OK) return rc;
    if (want_stats) {
        unsigned long long* hp = (unsigned long long*)(h->h_pinned + 128);
        IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, d_stats, 32, cudaMemcpyDeviceToHost, st));
        IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
        for (int k = 0; k < 4; ++k) h->last_stats[k] = (int64_t)hp[k];
        a.stats = nullptr;
    }
    if (!d_contacts || *num_contacts == 0) return IBVH_OK;
    if (*num_contacts > capacity) return IBVH_ERR_CAPACITY;
    return launch_traverse<KIND, kWrite, PACKET, LQ, LT, N, I>(h, qleaves, points, dirs, bvh, a, counts, (IndexPair<I>*)d_contacts, st);
}

// ---- tiled traversal driver (BBox nodes): group walk -> scan -> lists -> tiles ------------------------------
constexpr int kTileG = 8;          // leaves per group
constexpr int kTileLogG = 3;
constexpr uint32_t kTileSegCap = 128;    // target groups recorded per query group before it is flagged
constexpr uint32_t kTileStepCap = 2048;  // walk steps per query group before it is flagged

template <class LT> bool tiled_applicable(const TreeInfo& ti, int start_level) {
    return ti.levels - kTileLogG >= 1 && start_level <= ti.levels - kTileLogG && ti.levels >= kTileLogG + 2;
}

template <int KIND, class LQ, class LT, class I>
int traverse_tiled(ibvh_handle* h, const LQ* qleaves, int64_t n_query_total, const DBvh<LT, BBox<typename LT::value_type>>& bvh,
                   const TraverseArgs& ta, uint32_t flags, void* d_counts, void* d_contacts, int64_t capacity, int64_t* num_contacts,
                   cudaStream_t st) {
    constexpr int G = kTileG;
    using N = BBox<typename LT::value_type>;
    unsigned long long* d_total = (unsigned long long*)(h->d_small + kSmallTotal);
    uint32_t* d_fcount = (uint32_t*)(h->d_small + kSmallTotal + 16);
    *num_contacts = 0;
    if (ta.q_count <= 0) return IBVH_OK;
    GroupArgs a{};
    a.q_begin = ta.q_begin; a.q_count = ta.q_count; a.n_groups = (ta.q_count + G - 1) / G;
    a.start_level = ta.start_level; a.group_level = bvh.ti.levels - kTileLogG; a.flip = ta.flip; a.positions = ta.positions;
    a.seg_cap = kTileSegCap; a.step_cap = kTileStepCap;
    a.capacity = d_contacts ? capacity : 0; a.total = d_total;
    a.dbg = nullptr;
    const bool dbg = h->cfg.debug;
    if (dbg) { a.dbg = (unsigned long long*)(h->d_small + 1024); IBVH_CUDA_TRY(h, cudaMemsetAsync(a.dbg, 0, 640, st)); }
    const bool unordered = (flags & IBVH_TRAVERSE_UNORDERED) != 0 && d_contacts != nullptr;
    const int64_t qblocks = (a.q_count + kScanTile - 1) / kScanTile;
    size_t need = 2 * ibvh_handle::padded((size_t)a.n_groups * 4) + ibvh_handle::padded((size_t)qblocks * 8) +
                  (d_counts ? 0 : ibvh_handle::padded((size_t)a.q_count * sizeof(I))) + 4096;
    int rc = h->reserve(need);
    if (rc != IBVH_OK) return rc;
    h->reset();
    uint32_t* gcounts = h->alloc<uint32_t>(a.n_groups);
    uint32_t* flist = h->alloc<uint32_t>(a.n_groups);
    long long* qsums = h->alloc<long long>(qblocks);
    I* counts = d_counts ? (I*)d_counts : h->alloc<I>(a.q_count);
    if (!gcounts || !flist || !qsums || !counts) { h->set_error("workspace carve failed"); return IBVH_ERR_ALLOC; }
    rc = h->reserve_aux((size_t)a.n_groups * a.seg_cap * 4);
    if (rc != IBVH_OK) return rc;
    uint32_t* glist = (uint32_t*)h->aux;

    // phase 1 (one pass, bounded): per-group target lists + the list of flagged groups
    IBVH_CUDA_TRY(h, cudaMemsetAsync(d_fcount, 0, 4, st));
    const unsigned wblocks = (unsigned)((a.n_groups + kWalkThreads - 1) / kWalkThreads);
    { ProfScope _ps(h, st, "group_walk_kernel");
    group_walk_kernel<KIND, G, LQ, LT><<<wblocks, kWalkThreads, 0, st>>>(qleaves, n_query_total, bvh, a, gcounts, glist, flist, d_fcount);
    }
    IBVH_LAUNCH_CHECK(h, "group_walk_kernel");
    if (dbg) {
        unsigned long long hbuf[80];
        cudaMemcpy(hbuf, a.dbg, 640, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ibvh debug] groups=%lld walk steps=%llu (%.1f/group) flagged=%llu\n  steps log2 hist:", (long long)a.n_groups,
                hbuf[0], (double)hbuf[0] / a.n_groups, hbuf[1]);
        for (int k = 0; k < 20; ++k) fprintf(stderr, " %llu", hbuf[2 + k]);
        fprintf(stderr, "\n  pairs log2 hist:");
        for (int k = 0; k < 12; ++k) fprintf(stderr, " %llu", hbuf[40 + k]);
        fprintf(stderr, "\n");
        a.dbg = nullptr;
    }
    // the per-query kernel picks up the members of flagged groups (grid-stride over flist)
    TraverseArgs fa = ta;
    fa.total = d_total; fa.stats = nullptr; fa.capacity = a.capacity;
    fa.qmap = flist; fa.qmap_count = d_fcount; fa.qmap_group = G;
    const unsigned fblocks = (unsigned)std::min<int64_t>((a.q_count + 127) / 128, (int64_t)h->sm_count * 8);
    constexpr int SLOTS = 32 / G;
    const int64_t twarps = (a.n_groups + SLOTS - 1) / SLOTS;
    const unsigned tblocks = (unsigned)((twarps + kTileWarps - 1) / kTileWarps);
    auto run = [&](auto mode_tag, I* cnts, IndexPair<I>* out) -> int {
        constexpr int MODE = decltype(mode_tag)::value;
        IBVH_CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));           // everything the side stream needs is enqueued
        // members of flagged groups: per-query kernel on the high-priority side stream, submitted first so that
        // its few long-running blocks start immediately and overlap the tile kernel
        IBVH_CUDA_TRY(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        { ProfScope _ps(h, h->side, "lvt_thread_kernel");
        lvt_thread_kernel<KIND, MODE, LQ, LT, N, I><<<fblocks, 128, 0, h->side>>>(qleaves, nullptr, nullptr, bvh, fa, cnts, out);
        }
        IBVH_LAUNCH_CHECK(h, "lvt_thread_kernel(flagged groups)");
        if constexpr (MODE == kWrite) {
            { ProfScope _ps(h, st, "tile_kernel");
            tile_kernel<KIND, MODE, G, LQ, LT, I><<<tblocks, kTileWarps * 32, 0, st>>>(qleaves, n_query_total, bvh, a, gcounts, glist, cnts, out);
            }
        } else {
            { ProfScope _ps(h, st, "tile_flat_kernel");
            tile_flat_kernel<KIND, MODE, G, LQ, LT, I><<<tblocks, kTileWarps * 32, 0, st>>>(qleaves, n_query_total, bvh, a, gcounts, glist, cnts, out);
            }
        }
        IBVH_LAUNCH_CHECK(h, "tile_kernel");
        IBVH_CUDA_TRY(h, cudaEventRecord(h->ev_join, h->side));
        IBVH_CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join, 0));
        return IBVH_OK;
    };
    if (unordered) {
        IBVH_CUDA_TRY(h, cudaMemsetAsync(d_total, 0, 8, st));
        rc = run(std::integral_constant<int, kAtomic>{}, (I*)nullptr, (IndexPair<I>*)d_contacts);
        if (rc != IBVH_OK) return rc;
        rc = read_total(h, d_total, num_contacts, st);
        if (rc != IBVH_OK) return rc;
        return *num_contacts > capacity ? IBVH_ERR_CAPACITY : IBVH_OK;
    }
    if ((flags & IBVH_TRAVERSE_COUNTS_VALID) && d_counts && d_contacts) {
        I* hp = (I*)(h->h_pinned + 64);
        IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, counts + (a.q_count - 1), sizeof(I), cudaMemcpyDeviceToHost, st));
        IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
        *num_contacts = (int64_t)*hp;
    } else {
        rc = run(std::integral_constant<int, kCount>{}, counts, (IndexPair<I>*)nullptr);
        if (rc != IBVH_OK) return rc;
        rc = scan_counts<I>(h, counts, a.q_count, qsums, d_total, st);
        if (rc != IBVH_OK) return rc;
        rc = read_total(h, d_total, num_contacts, st);
        if (rc != IBVH_OK) return rc;
    }
    if (!d_contacts || *num_contacts == 0) return IBVH_OK;
    if (*num_contacts > capacity) return IBVH_ERR_CAPACITY;
    return run(std::integral_constant<int, kWrite>{}, counts, (IndexPair<I>*)d_contacts);
}

template <int KIND, class LQ, class LT, class I>
int traverse_pyramid(ibvh_handle* h, const LQ* qleaves, int64_t n_query_total, const DBvh<LT, BBox<typename LT::value_type>>& bvh,
                     const PyrPlan& plan, const TraverseArgs& ta, uint32_t flags, void* d_counts, void* d_contacts, int64_t capacity,
                     int64_t* num_contacts, cudaStream_t st) {
    using T = typename LT::value_type;
    using N = BBox<T>;
    *num_contacts = 0;
    // fused multi-GPU mode (peer.cuh, round-2 protocol): slots come from a LOCAL counter and index this rank's region of
    // the gathered list; the contacts go out through the multicast alias; counts are exchanged by the finish kernel
    const bool fused = ta.peer != nullptr;
    PeerArgs pa{};
    unsigned long long* out_total = (unsigned long long*)(h->d_small + kSmallTotal);
    IndexPair<I>* out_ptr = (IndexPair<I>*)d_contacts;
    int64_t* h_peer = (int64_t*)(h->h_pinned + 3072);
    long long region[IBVH_MAX_PEERS + 1] = {0};
    if (fused) {
        pa = make_peer_args(ta.peer);
        peer_regions(ta.peer, (int)sizeof(IndexPair<I>), region);
        out_ptr = (IndexPair<I>*)(pa.mc + pa.header_bytes) + region[pa.rank];
        capacity = region[pa.rank + 1] - region[pa.rank];
        h_peer[1] = 0;
        { ProfScope _ps(h, st, "peer_fused_begin_kernel");
        peer_fused_begin_kernel<<<1, 32, 0, st>>>(pa);
        }
        IBVH_LAUNCH_CHECK(h, "peer_fused_begin_kernel");
    }
    auto fused_finish = [&]() -> int {
        { ProfScope _ps(h, st, "peer_fused_finish_kernel");
        peer_fused_finish_kernel<<<1, 32, 0, st>>>(pa, out_total, h_peer);
        }
        IBVH_LAUNCH_CHECK(h, "peer_fused_finish_kernel");
        IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
        if (h_peer[1] != 0) { h->set_error("fused traversal: a peer did not arrive within 10 s"); return IBVH_ERR_PEER; }
        *num_contacts = h_peer[0];
        long long counts_r[IBVH_MAX_PEERS];
        bool over = false;
        for (int r = 0; r < pa.world; ++r) { counts_r[r] = h_peer[2 + r]; h->last_peer_counts[r] = counts_r[r]; if (counts_r[r] > region[r + 1] - region[r]) over = true; }
        if (over) return IBVH_ERR_CAPACITY;
        PeerMoves mv = make_compact_moves(pa.world, region, counts_r, (int)sizeof(IndexPair<I>) / 8);
        if (mv.n > 0) {
            { ProfScope _ps(h, st, "peer_compact_kernel");
            peer_compact_kernel<<<h->sm_count * 2, 256, 0, st>>>((uint64_t*)(pa.buf[pa.rank] + pa.header_bytes), mv);
            }
            IBVH_LAUNCH_CHECK(h, "peer_compact_kernel");
        }
        return IBVH_OK;
    };
    auto fused_wait = [&]() -> int {
        { ProfScope _ps(h, st, "peer_fused_wait_kernel");
        peer_fused_wait_kernel<<<1, 32, 0, st>>>(pa, h_peer);
        }
        IBVH_LAUNCH_CHECK(h, "peer_fused_wait_kernel");
        return IBVH_OK;
    };
    if (fused && ta.q_count <= 0) {
        IBVH_CUDA_TRY(h, cudaMemsetAsync(out_total, 0, 8, st));
        int rcw = fused_wait();
        if (rcw != IBVH_OK) return rcw;
    }
    if (ta.q_count <= 0) return fused ? fused_finish() : IBVH_OK;
    const int64_t q_begin = ta.q_begin, q_end = ta.q_begin + ta.q_count;
    unsigned long long* d_total = (unsigned long long*)(h->d_small + kSmallTotal);
    unsigned long long* d_cnt = (unsigned long long*)(h->d_small + 1536);      // one list counter per level
    uint32_t* d_tick = (uint32_t*)(h->d_small + 1536 + sizeof(unsigned long long) * kPyrMaxLevels);   // chunk tickets: [l] refine of level l, [16 + k] k-th tile launch
    int tile_launch = 0;
    const bool unordered = fused || ((flags & IBVH_TRAVERSE_UNORDERED) != 0 && d_contacts != nullptr);
    const bool count_only = !fused && d_contacts == nullptr;
    // ordered protocol with a contacts buffer and no valid counts yet: ONE tile pass that counts per query and stashes
    // the hits, then scan + scatter into the query segments (instead of a count pass and a write pass of the tile kernel)
    const bool stash_mode = !unordered && !count_only && capacity > 0 && !((flags & IBVH_TRAVERSE_COUNTS_VALID) && d_counts);
    const int nl = plan.n;
    const int grid = h->sm_count * h->cfg.pyr_grid;               // leaf-tile kernel
    const int grid_refine = h->sm_count * h->cfg.pyr_grid_refine;   // refine kernels
    const int pflip = (ta.flip ? 1 : 0) | (ta.positions ? 2 : 0);      // bit 1: report leaf positions (IBVH_TRAVERSE_POSITIONS)
    // The 16-byte hit stash of the one-pass ordered protocol is sized from what the traversals on this handle actually
    // produced (last total + 25 %, or 6 hits per query before the first one), never from the caller's `capacity`: a
    // generously pre-sized cache1 must not pull twice its size of scratch along. A stash that turns out too small costs
    // a write pass of the tile kernel (the per-query counts are valid either way), not an error.
    int64_t stash_cap = 0;
    if (stash_mode) {
        const int64_t est = h->stash_hint > 0 ? h->stash_hint + h->stash_hint / 4 + 4096 : 6 * ta.q_count + 4096;
        stash_cap = capacity < est ? capacity : est;
    }

    // ordered protocol scratch: counts (if the caller gave none), cursors, scan sums
    const int64_t qblocks = (ta.q_count + kScanTile - 1) / kScanTile;
    // (a segment longer than kFixupInsertion needs that many contacts: at most capacity / kFixupInsertion of them exist)
    const uint32_t long_cap = (uint32_t)std::min<int64_t>(ta.q_count, (capacity > 0 ? capacity : 0) / kFixupInsertion + 1);
    size_t need = ibvh_handle::padded((size_t)qblocks * 8) + ibvh_handle::padded((size_t)ta.q_count * 4) +
                  (d_counts ? 0 : ibvh_handle::padded((size_t)ta.q_count * sizeof(I))) + ibvh_handle::padded((size_t)long_cap * 8) + 4096;
    int rc = h->reserve(need);
    if (rc != IBVH_OK) return rc;
    h->reset();
    long long* qsums = h->alloc<long long>(qblocks);
    unsigned int* cursors = h->alloc<unsigned int>(ta.q_count);
    I* counts = d_counts ? (I*)d_counts : h->alloc<I>(ta.q_count);
    uint32_t* long_lists = h->alloc<uint32_t>((size_t)long_cap * 2);
    if (!qsums || !cursors || !counts || !long_lists) { h->set_error("workspace carve failed"); return IBVH_ERR_ALLOC; }

    double factor = h->pyr_factor > 0 ? h->pyr_factor : (KIND == kSingle ? 40.0 : 80.0);
    unsigned long long need_cap[kPyrMaxLevels] = {0};
    for (int attempt = 0; attempt < 4; ++attempt) {
        // carve: pyramid boxes + one pair list per level (+ what no sidecar provides: packed volumes, aligned node levels)
        unsigned long long cap[kPyrMaxLevels];
        using VQ = typename LQ::vol_t;
        using VT = typename LT::vol_t;
        const bool same_leaves = (const void*)qleaves == (const void*)bvh.leaves && std::is_same<LQ, LT>::value;
        // sidecars written by ibvh_build (handle.cuh): the target tree's and — pair traversal — the query tree's
        ibvh_handle::Sidecar* sct = h->find_sidecar(ta.t_build_id, bvh.ti.n);
        ibvh_handle::Sidecar* scq = same_leaves ? sct : h->find_sidecar(ta.q_build_id, n_query_total);
        SidecarLayout<LT, N> lay_t{};
        SidecarLayout<LQ, N> lay_q{};
        if (sct) {
            lay_t = sidecar_layout<LT, N>(bvh.ti, sct->built_level);
            if (!lay_t.ok || lay_t.bytes > sct->bytes || sct->float_bytes != (int)sizeof(T) || sct->leaf_kind != VT::kind || lay_t.plan.n < nl) sct = nullptr;
        }
        if (scq) {
            if (same_leaves) {
                if (!sct) scq = nullptr;
            } else {
                ibvh_tree_t tq;
                make_tree(n_query_total, &tq);
                lay_q = sidecar_layout<LQ, N>(make_tree_info(tq), scq->built_level);
                if (!lay_q.ok || lay_q.bytes > scq->bytes || scq->float_bytes != (int)sizeof(T) || scq->leaf_kind != VQ::kind) scq = nullptr;
            }
        }
        const int q_ulevels = scq ? std::min(nl, same_leaves ? lay_t.u_levels : lay_q.u_levels) : 0;
        // refine over conservatively quantised boxes (traverse_pyramid.cuh, 3a): needs the target's root box as the frame
        const bool quant = h->cfg.pyr_quant && !h->cfg.pyr_tma && ta.t_built_level == 1 && nl >= 2;
        size_t bytes = ibvh_handle::padded((size_t)plan.u_total * sizeof(UBox<T>)) +
                       (quant ? ibvh_handle::padded((size_t)plan.u_total * sizeof(QBoxU)) + ibvh_handle::padded((size_t)plan.t_total * sizeof(QBoxT)) : 0) +
                       (sct ? 0 : ibvh_handle::padded((size_t)plan.t_total * sizeof(N)) + ibvh_handle::padded((size_t)(bvh.ti.n + 8) * sizeof(Packed<VT>))) +
                       ((same_leaves || scq) ? 0 : ibvh_handle::padded((size_t)(n_query_total + 8) * sizeof(Packed<VQ>))) +
                       (stash_mode ? ibvh_handle::padded((size_t)stash_cap * sizeof(uint4)) : 0);
        for (int l = 0; l < nl; ++l) {
            const PyrLevel& v = plan.lv[l];
            unsigned long long c = (unsigned long long)(factor * (double)(v.nqg > v.ntg && KIND != kSingle ? v.nqg : v.nqg)) + 4096ull;
            if (l == nl - 1) { unsigned long long all = (unsigned long long)(v.nqg * v.ntg); if (all < c) c = all; }
            if (need_cap[l] > c) c = need_cap[l] + need_cap[l] / 8 + 4096ull;
            cap[l] = c;
            bytes += ibvh_handle::padded((size_t)c * sizeof(uint2));
        }
        rc = h->reserve_aux(bytes);
        if (rc != IBVH_OK) return rc;
        char* ap = h->aux;
        UBox<T>* U = (UBox<T>*)ap; ap += ibvh_handle::padded((size_t)plan.u_total * sizeof(UBox<T>));
        QBoxU* Uq = nullptr; QBoxT* NTq = nullptr;
        if (quant) {
            Uq = (QBoxU*)ap; ap += ibvh_handle::padded((size_t)plan.u_total * sizeof(QBoxU));
            NTq = (QBoxT*)ap; ap += ibvh_handle::padded((size_t)plan.t_total * sizeof(QBoxT));
        }
        N* NT = nullptr;                                             // aligned copy of the target node levels
        Packed<VT>* PT = nullptr;
        if (sct) {
            NT = (N*)(sct->buf + lay_t.nt_off);
            PT = (Packed<VT>*)(sct->buf + lay_t.pt_off);
        } else {
            NT = (N*)ap; ap += ibvh_handle::padded((size_t)plan.t_total * sizeof(N));
            PT = (Packed<VT>*)ap; ap += ibvh_handle::padded((size_t)(bvh.ti.n + 8) * sizeof(Packed<VT>));
        }
        // compact leaf-index arrays of the sidecars (else the kernels fetch .index from the 24 / 32-byte leaf structs: two
        // random 32-byte sectors per contact, 2.5 GB of the tile kernel's 3.2 GB of DRAM traffic at 10 M leaves)
        const I* TIDX = (sct && sct->index_bytes == (int)sizeof(I)) ? (const I*)(sct->buf + lay_t.idx_off) : nullptr;
        const I* QIDX = same_leaves ? TIDX : ((scq && scq->index_bytes == (int)sizeof(I)) ? (const I*)(scq->buf + lay_q.idx_off) : nullptr);
        Packed<VQ>* PQ = (Packed<VQ>*)PT;
        if (!same_leaves) {
            if (scq) PQ = (Packed<VQ>*)(scq->buf + lay_q.pt_off);
            else { PQ = (Packed<VQ>*)ap; ap += ibvh_handle::padded((size_t)(n_query_total + 8) * sizeof(Packed<VQ>)); }
        }
        uint4* stash = nullptr;
        if (stash_mode) { stash = (uint4*)ap; ap += ibvh_handle::padded((size_t)stash_cap * sizeof(uint4)); }
        PairList lists[kPyrMaxLevels];
        for (int l = 0; l < nl; ++l) { lists[l].data = (uint2*)ap; lists[l].count = d_cnt + l; lists[l].cap = cap[l]; ap += ibvh_handle::padded((size_t)cap[l] * sizeof(uint2)); }
        IBVH_CUDA_TRY(h, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * kPyrMaxLevels + 128, st));   // list counters + chunk tickets
        IBVH_CUDA_TRY(h, cudaMemsetAsync(d_total, 0, 8, st));
        // per-level box pointers: query pyramid (index = group - qg_first of the shard's plan) and aligned target nodes.
        // A sidecar's pyramid covers the WHOLE tree; a group cut by the shard range then carries the box of all its leaves,
        // a superset of the in-range ones: conservative (never drops a pair), and the tile kernel masks the queries itself.
        const UBox<T>* Ulev[kPyrMaxLevels];
        const N* NTlev[kPyrMaxLevels];
        for (int l = 0; l < nl; ++l) {
            Ulev[l] = U + plan.lv[l].u_off;
            if (l < q_ulevels) {
                const char* ub = scq->buf + (same_leaves ? lay_t.u_off : lay_q.u_off);
                Ulev[l] = (const UBox<T>*)ub + (same_leaves ? lay_t.u_level_off[l] : lay_q.u_level_off[l]) + plan.lv[l].qg_first;
            }
            NTlev[l] = sct ? NT + lay_t.plan.lv[l].t_off : NT + plan.lv[l].t_off;
        }

        // 0. 16-byte aligned records: leaf volumes, and the node levels the refinement reads as targets (unless a sidecar has them)
        const int64_t qend_c = q_end < n_query_total ? q_end : n_query_total;
        if (same_leaves) {
            if (!sct) {
                // targets == queries: one pass packs the volumes AND builds the finest level of the query pyramid
                const int64_t groups = (bvh.ti.n + 8 + 3) / 4;
                { ProfScope _ps(h, st, "pyr_pack_groups_kernel");
                pyr_pack_groups_kernel<LT, T><<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(bvh.leaves, bvh.ti.n, bvh.ti.n + 8, PT, q_begin, qend_c, plan.lv[0], U);
                }
                IBVH_LAUNCH_CHECK(h, "pyr_pack_groups_kernel");
            }
        } else {
            if (!sct) {
                { ProfScope _ps(h, st, "pyr_pack_volumes_kernel");
                pyr_pack_volumes_kernel<LT><<<(unsigned)((bvh.ti.n + 8 + 255) / 256), 256, 0, st>>>(bvh.leaves, bvh.ti.n, bvh.ti.n + 8, PT);
                }
                IBVH_LAUNCH_CHECK(h, "pyr_pack_volumes_kernel");
            }
            if (!scq) {
                const int64_t groups = (n_query_total + 8 + 3) / 4;
                { ProfScope _ps(h, st, "pyr_pack_groups_kernel");
                pyr_pack_groups_kernel<LQ, T><<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(qleaves, n_query_total, n_query_total + 8, PQ, q_begin, qend_c, plan.lv[0], U);
                }
                IBVH_LAUNCH_CHECK(h, "pyr_pack_groups_kernel");
            }
        }
        if (!sct)
            for (int l = 0; l + 1 < nl; ++l)
                IBVH_CUDA_TRY(h, cudaMemcpyAsync(NT + plan.lv[l].t_off, bvh.nodes + plan.lv[l].tnode0, (size_t)plan.lv[l].ntg * sizeof(N), cudaMemcpyDeviceToDevice, st));
        // 1. query pyramid above the levels at hand (level 0 comes from the packing pass, levels 0 .. 2 from a sidecar)
        for (int l = std::max(1, q_ulevels); l < nl; ++l) {
            { ProfScope _ps(h, st, "pyr_up_kernel");
            pyr_up_kernel<T><<<(unsigned)((plan.lv[l].nqg + 255) / 256), 256, 0, st>>>(plan.lv[l - 1], plan.lv[l], Ulev[l - 1], U + plan.lv[l].u_off);
            }
            IBVH_LAUNCH_CHECK(h, "pyr_up_kernel");
        }
        // 2. top all-pairs
        {
            const PyrLevel& top = plan.lv[nl - 1];
            int64_t tot = top.nqg * top.ntg;
            int tg = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 16);
            { ProfScope _ps(h, st, "pyr_top_kernel");
            pyr_top_kernel<KIND, T><<<tg, 256, 0, st>>>(top, Ulev[nl - 1], bvh.nodes, lists[nl - 1]);
            }
            IBVH_LAUNCH_CHECK(h, "pyr_top_kernel");
        }
        // 3. refine down to the 4-leaf groups
        if (quant) {
            // the fine levels of every refinement step as 15-bit integer boxes in the frame of the target's root box
            for (int l = 0; l + 1 < nl; ++l) {
                const PyrLevel& v = plan.lv[l];
                { ProfScope _ps(h, st, "pyr_quantize_kernel");
                pyr_quantize_kernel<T, UBox<T>, QBoxU><<<(unsigned)((v.nqg + 255) / 256), 256, 0, st>>>(Ulev[l], v.nqg, bvh.nodes, Uq + v.u_off);
                }
                const int64_t nt_pad = (v.ntg + 7) & ~int64_t(7);
                { ProfScope _ps(h, st, "pyr_quantize_kernel");          // (one scope per launch: bench.py counts launches by scopes)
                pyr_quantize_kernel<T, N, QBoxT><<<(unsigned)((nt_pad + 255) / 256), 256, 0, st>>>(NTlev[l], nt_pad, bvh.nodes, NTq + v.t_off);
                }
                IBVH_LAUNCH_CHECK(h, "pyr_quantize_kernel");
            }
        }
        for (int l = nl - 1; l >= 1; --l) {
            { ProfScope _ps(h, st, "pyr_refine_kernel");
            if (quant && h->cfg.pyr_q2 == 4)
                pyr_refine_q2_kernel<KIND, 4><<<grid_refine, kPyrWarps * 32, 0, st>>>(Uq + plan.lv[l - 1].u_off, NTq + plan.lv[l - 1].t_off, (uint32_t)plan.lv[l - 1].qg_first, (uint32_t)plan.lv[l - 1].nqg, (uint32_t)plan.lv[l - 1].ntg, lists[l], lists[l - 1], d_tick + l);
            else if (quant && h->cfg.pyr_q2)
                pyr_refine_q2_kernel<KIND, 2><<<grid_refine, kPyrWarps * 32, 0, st>>>(Uq + plan.lv[l - 1].u_off, NTq + plan.lv[l - 1].t_off, (uint32_t)plan.lv[l - 1].qg_first, (uint32_t)plan.lv[l - 1].nqg, (uint32_t)plan.lv[l - 1].ntg, lists[l], lists[l - 1], d_tick + l);
            else if (quant)
                pyr_refine_q_kernel<KIND><<<grid_refine, kPyrWarps * 32, 0, st>>>(Uq + plan.lv[l - 1].u_off, NTq + plan.lv[l - 1].t_off, (uint32_t)plan.lv[l - 1].qg_first, (uint32_t)plan.lv[l - 1].nqg, (uint32_t)plan.lv[l - 1].ntg, lists[l], lists[l - 1], d_tick + l);
            else if (h->cfg.pyr_tma)
                pyr_refine_tma_kernel<KIND, T><<<grid_refine, kPyrWarps * 32, 0, st>>>(Ulev[l - 1], NTlev[l - 1], (uint32_t)plan.lv[l - 1].qg_first, (uint32_t)plan.lv[l - 1].nqg, (uint32_t)plan.lv[l - 1].ntg, lists[l], lists[l - 1], d_tick + l);
            else
                pyr_refine_kernel<KIND, T><<<grid_refine, kPyrWarps * 32, 0, st>>>(Ulev[l - 1], NTlev[l - 1], (uint32_t)plan.lv[l - 1].qg_first, (uint32_t)plan.lv[l - 1].nqg, (uint32_t)plan.lv[l - 1].ntg, lists[l], lists[l - 1], d_tick + l);
            }
            IBVH_LAUNCH_CHECK(h, "pyr_refine_kernel");
        }
        // 4. leaf tiles
        const int64_t qe = q_end < n_query_total ? q_end : n_query_total;
        if (fused) {
            // the tile kernel publishes into every rank's list: it must run exactly once, so the pair-list overflow
            // check (and the retry with larger lists) comes before it
            unsigned long long* hp = (unsigned long long*)h->h_pinned;
            IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp + 1, d_cnt, sizeof(unsigned long long) * kPyrMaxLevels, cudaMemcpyDeviceToHost, st));
            IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
            bool overflow = false;
            double worst = 0.0;
            for (int l = 0; l < nl; ++l) {
                unsigned long long c = hp[1 + l];
                need_cap[l] = c;
                if (c > cap[l]) overflow = true;
                double r = (double)c / (double)plan.lv[l].nqg;
                if (l < nl - 1 && r > worst) worst = r;
            }
            if (overflow) { factor = worst * 1.15 + 2.0; continue; }
            h->pyr_factor = worst * 1.25 + 4.0;
            {   // test counts of this traversal, from the pair-list sizes (ibvh_last_traversal_stats)
                const int64_t F2 = int64_t(1) << (2 * kPyrFan), G2 = int64_t(1) << (2 * kPyrLeafLog);
                int64_t box = plan.lv[nl - 1].nqg * plan.lv[nl - 1].ntg;
                for (int l = nl - 1; l >= 1; --l) box += (int64_t)std::min(hp[1 + l], cap[l]) * F2;
                h->last_stats[0] = box;
                h->last_stats[1] = (int64_t)std::min(hp[1], cap[0]) * G2;
                h->last_stats[2] = (int64_t)hp[1];
                h->last_stats[3] = nl;
            }
            {   // every rank has let go of the previous list before the first store of this call leaves
                int rcw = fused_wait();
                if (rcw != IBVH_OK) return rcw;
            }
            { ProfScope _ps(h, st, "pyr_leaf_tile_kernel");
            pyr_leaf_tile_kernel<KIND, kAtomic, 0, LQ, LT, I><<<grid, kPyrWarps * 32, 0, st>>>(qleaves, q_begin, qe, bvh, lists[0], pflip, capacity, out_total, (I*)nullptr, nullptr, out_ptr, 1, d_tick + 16 + (tile_launch++), PQ, PT, QIDX, TIDX);
            }
            IBVH_LAUNCH_CHECK(h, "pyr_leaf_tile_kernel");
            return fused_finish();
        }
        bool counts_valid = (flags & IBVH_TRAVERSE_COUNTS_VALID) && d_counts && d_contacts;
        if (unordered || count_only) {
            if (unordered) {
                { ProfScope _ps(h, st, "pyr_leaf_tile_kernel");
                pyr_leaf_tile_kernel<KIND, kAtomic, 0, LQ, LT, I><<<grid, kPyrWarps * 32, 0, st>>>(qleaves, q_begin, qe, bvh, lists[0], pflip, capacity, out_total, (I*)nullptr, nullptr, out_ptr, fused ? 1 : 0, d_tick + 16 + (tile_launch++), PQ, PT, QIDX, TIDX);
                }
            } else if (d_counts) {
                // count-only call of the ordered protocol: per-query counts + scan (cache2), total from the scan
                IBVH_CUDA_TRY(h, cudaMemsetAsync(counts, 0, (size_t)ta.q_count * sizeof(I), st));
                { ProfScope _ps(h, st, "pyr_leaf_tile_kernel");
                pyr_leaf_tile_kernel<KIND, kCount, 1, LQ, LT, I><<<grid, kPyrWarps * 32, 0, st>>>(qleaves, q_begin, qe, bvh, lists[0], pflip, 0, d_total, counts, nullptr, (IndexPair<I>*)nullptr, 0, d_tick + 16 + (tile_launch++), PQ, PT, QIDX, TIDX);
                }
                IBVH_LAUNCH_CHECK(h, "pyr_leaf_tile_kernel");
                rc = scan_counts<I>(h, counts, ta.q_count, qsums, d_total, st);
                if (rc != IBVH_OK) return rc;
            } else {
                { ProfScope _ps(h, st, "pyr_leaf_tile_kernel");
                pyr_leaf_tile_kernel<KIND, kCount, 0, LQ, LT, I><<<grid, kPyrWarps * 32, 0, st>>>(qleaves, q_begin, qe, bvh, lists[0], pflip, 0, d_total, (I*)nullptr, nullptr, (IndexPair<I>*)nullptr, 0, d_tick + 16 + (tile_launch++), PQ, PT, QIDX, TIDX);
                }
            }
            IBVH_LAUNCH_CHECK(h, "pyr_leaf_tile_kernel");
        } else if (!counts_valid) {
            IBVH_CUDA_TRY(h, cudaMemsetAsync(counts, 0, (size_t)ta.q_count * sizeof(I), st));
            { ProfScope _ps(h, st, "pyr_leaf_tile_kernel");
            if (stash_mode)
                pyr_leaf_tile_kernel<KIND, kAtomic, 3, LQ, LT, I><<<grid, kPyrWarps * 32, 0, st>>>(qleaves, q_begin, qe, bvh, lists[0], pflip, stash_cap, d_total, counts, nullptr, (IndexPair<I>*)stash, 0, d_tick + 16 + (tile_launch++), PQ, PT, QIDX, TIDX);
            else
                pyr_leaf_tile_kernel<KIND, kCount, 1, LQ, LT, I><<<grid, kPyrWarps * 32, 0, st>>>(qleaves, q_begin, qe, bvh, lists[0], pflip, 0, d_total, counts, nullptr, (IndexPair<I>*)nullptr, 0, d_tick + 16 + (tile_launch++), PQ, PT, QIDX, TIDX);
            }
            IBVH_LAUNCH_CHECK(h, "pyr_leaf_tile_kernel");
            rc = scan_counts<I>(h, counts, ta.q_count, qsums, d_total, st);
            if (rc != IBVH_OK) return rc;
        }
        // one read-back: contact total + the list counters (overflow check)
        unsigned long long* hp = (unsigned long long*)h->h_pinned;
        if (counts_valid && !unordered && !count_only) {
            I* hpi = (I*)(h->h_pinned + 512);
            IBVH_CUDA_TRY(h, cudaMemcpyAsync(hpi, counts + (ta.q_count - 1), sizeof(I), cudaMemcpyDeviceToHost, st));
            IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp + 1, d_cnt, sizeof(unsigned long long) * kPyrMaxLevels, cudaMemcpyDeviceToHost, st));
            IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
            hp[0] = (unsigned long long)*hpi;
        } else {
            IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, d_total, 8, cudaMemcpyDeviceToHost, st));
            IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp + 1, d_cnt, sizeof(unsigned long long) * kPyrMaxLevels, cudaMemcpyDeviceToHost, st));
            if (unordered && (flags & IBVH_TRAVERSE_DEFER)) {
                // deferred: everything is enqueued; ibvh_traverse_finish judges the read-back later
                IBVH_CUDA_TRY(h, cudaEventRecord(h->ev_defer, st));
                ibvh_handle::Deferred& df = h->deferred;
                df.active = true; df.nl = nl; df.capacity = capacity;
                df.top_pairs = plan.lv[nl - 1].nqg * plan.lv[nl - 1].ntg;
                df.fan2 = 1 << (2 * kPyrFan); df.leaf2 = 1 << (2 * kPyrLeafLog);
                for (int l = 0; l < nl; ++l) { df.cap[l] = cap[l]; df.nqg[l] = plan.lv[l].nqg; }
                *num_contacts = -1;
                return IBVH_OK;
            }
            IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
        }
        bool overflow = false;
        double worst = 0.0;
        for (int l = 0; l < nl; ++l) {
            unsigned long long c = hp[1 + l];
            need_cap[l] = c;
            if (c > cap[l]) overflow = true;
            double r = (double)c / (double)plan.lv[l].nqg;
            if (l < nl - 1 && r > worst) worst = r;
        }
        if (h->cfg.debug) {
            fprintf(stderr, "[ibvh debug] pyramid levels=%d:", nl);
            for (int l = nl - 1; l >= 0; --l) fprintf(stderr, " k=%d pairs=%llu/%llu", plan.lv[l].k, hp[1 + l], cap[l]);
            fprintf(stderr, " contacts=%llu%s\n", hp[0], overflow ? " OVERFLOW -> retry" : "");
        }
        if (overflow) { factor = worst * 1.15 + 2.0; continue; }
        h->pyr_factor = worst * 1.25 + 4.0;
        {   // test counts of this traversal, from the pair-list sizes (ibvh_last_traversal_stats)
            const int64_t F2 = int64_t(1) << (2 * kPyrFan), G2 = int64_t(1) << (2 * kPyrLeafLog);
            int64_t box = plan.lv[nl - 1].nqg * plan.lv[nl - 1].ntg;
            for (int l = nl - 1; l >= 1; --l) box += (int64_t)std::min(hp[1 + l], cap[l]) * F2;
            h->last_stats[0] = box;
            h->last_stats[1] = (int64_t)std::min(hp[1], cap[0]) * G2;
            h->last_stats[2] = (int64_t)hp[1];
            h->last_stats[3] = nl;
        }
        *num_contacts = (int64_t)hp[0];
        if (!unordered) h->stash_hint = *num_contacts;
        if (unordered) return *num_contacts > capacity ? IBVH_ERR_CAPACITY : IBVH_OK;
        if (count_only || *num_contacts == 0) return IBVH_OK;
        if (*num_contacts > capacity) return IBVH_ERR_CAPACITY;
        // ordered write: per-query cursors (from the stash of the single tile pass, or — counts valid from an earlier
        // call — a write pass of the tile kernel), then sort each query's handful of hits by target position
        IBVH_CUDA_TRY(h, cudaMemsetAsync(cursors, 0, (size_t)ta.q_count * 4, st));
        if (stash_mode && !counts_valid && *num_contacts <= stash_cap) {
            { ProfScope _ps(h, st, "pyr_scatter_kernel");
            pyr_scatter_kernel<I><<<h->sm_count * 16, 256, 0, st>>>(stash, d_total, stash_cap, counts, cursors, (IndexPair<I>*)d_contacts);
            }
            IBVH_LAUNCH_CHECK(h, "pyr_scatter_kernel");
        } else {
            { ProfScope _ps(h, st, "pyr_leaf_tile_kernel");
            pyr_leaf_tile_kernel<KIND, kWrite, 2, LQ, LT, I><<<grid, kPyrWarps * 32, 0, st>>>(qleaves, q_begin, qe, bvh, lists[0], pflip, capacity, d_total, counts, cursors, (IndexPair<I>*)d_contacts, 0, d_tick + 16 + (tile_launch++), PQ, PT, QIDX, TIDX);
            }
            IBVH_LAUNCH_CHECK(h, "pyr_leaf_tile_kernel");
        }
        // per-query sort of the hits by target position; segments longer than kFixupInsertion are queued and sorted
        // cooperatively (one warp / one block each) by the two launches that follow
        uint32_t* long_counts = (uint32_t*)(h->d_small + kSmallFixup);
        IBVH_CUDA_TRY(h, cudaMemsetAsync(long_counts, 0, 8, st));
        { ProfScope _ps(h, st, "pyr_fixup_kernel");
        pyr_fixup_kernel<KIND, LQ, LT, I><<<(unsigned)((ta.q_count + 255) / 256), 256, 0, st>>>(qleaves, QIDX, q_begin, ta.q_count, pflip, counts, (IndexPair<I>*)d_contacts, long_lists, long_counts, long_cap);
        }
        IBVH_LAUNCH_CHECK(h, "pyr_fixup_kernel");
        { ProfScope _ps(h, st, "pyr_fixup_long_kernel");
        pyr_fixup_long_kernel<KIND, true, LQ, I><<<h->sm_count * 2, 256, 0, st>>>(qleaves, q_begin, pflip, counts, (IndexPair<I>*)d_contacts, long_lists, long_counts, long_cap);
        }
        { ProfScope _ps(h, st, "pyr_fixup_long_kernel");
        pyr_fixup_long_kernel<KIND, false, LQ, I><<<h->sm_count * 2, 256, 0, st>>>(qleaves, q_begin, pflip, counts, (IndexPair<I>*)d_contacts, long_lists + long_cap, long_counts + 1, long_cap);
        }
        IBVH_LAUNCH_CHECK(h, "pyr_fixup_long_kernel");
        return IBVH_OK;
    }
    h->set_error("pyramid pair lists kept overflowing");
    return IBVH_ERR_ALLOC;
}

// Schedule choice for leaf queries: tiled (BBox nodes, tall enough tree, start level above the groups),
// else the packet schedule; IBVH_TRAVERSE_REFERENCE_SHAPED / IBVH_TRAVERSE_PACKET force the others.
template <int KIND, class LQ, class LT, class N, class I>
int traverse_leaf_queries(ibvh_handle* h, const LQ* qleaves, int64_t n_query_total, const DBvh<LT, N>& d, int64_t built_level, const TraverseArgs& a,
                          uint32_t flags, void* d_counts, void* d_contacts, int64_t capacity, int64_t* num_contacts, cudaStream_t st) {
    if (a.peer) {
        // fused traversal + all-gather: pyramid schedule with multicast output only; everything else is
        // "traverse locally, then ibvh_allgather_pairs"
        bool ok = false;
        if constexpr (std::is_same<N, BBox<typename LT::value_type>>::value) {
            PyrPlan plan;
            ok = peer_ok(a.peer) && a.peer->multicast && a.peer->fused_seq > 0 && (flags & IBVH_TRAVERSE_UNORDERED) &&
                 !(flags & (IBVH_TRAVERSE_REFERENCE_SHAPED | IBVH_TRAVERSE_PACKET | IBVH_TRAVERSE_STATS | IBVH_TRAVERSE_WALK)) &&
                 a.start_level <= d.ti.levels - 1 && d.ti.n < (int64_t(1) << 29) && n_query_total < (int64_t(1) << 29) &&
                 make_pyr_plan(d.ti, built_level, 0, n_query_total, &plan);       // same verdict on every rank
            if (ok) {
                if (a.q_count > 0 && !make_pyr_plan(d.ti, built_level, a.q_begin, a.q_count, &plan)) ok = false;
                if (ok) return traverse_pyramid<KIND, LQ, LT, I>(h, qleaves, n_query_total, d, plan, a, flags, d_counts, d_contacts, capacity, num_contacts, st);
            }
        }
        h->set_error("fused multi-GPU traversal needs BBox nodes, the pyramid schedule, IBVH_TRAVERSE_UNORDERED and a multicast alias");
        return IBVH_ERR_UNSUPPORTED;
    }
    if (flags & IBVH_TRAVERSE_REFERENCE_SHAPED)
        return traverse_impl<KIND, false, LQ, LT, N, I>(h, qleaves, nullptr, nullptr, d, a, flags, d_counts, d_contacts, capacity, num_contacts, st);
    if constexpr (std::is_same<N, BBox<typename LT::value_type>>::value) {
        // pyramid refinement: needs start_level above the leaves (so that the leaf-parent test is part of the
        // reference predicate) and n < 2^31 positions in 32-bit pair entries
        if (!(flags & (IBVH_TRAVERSE_PACKET | IBVH_TRAVERSE_STATS | IBVH_TRAVERSE_WALK)) && a.start_level <= d.ti.levels - 1 &&
            d.ti.n < (int64_t(1) << 29) && n_query_total < (int64_t(1) << 29)) {
            PyrPlan plan;
            if (make_pyr_plan(d.ti, built_level, a.q_begin, a.q_count, &plan))
                return traverse_pyramid<KIND, LQ, LT, I>(h, qleaves, n_query_total, d, plan, a, flags, d_counts, d_contacts, capacity, num_contacts, st);
        }
        if (!(flags & (IBVH_TRAVERSE_PACKET | IBVH_TRAVERSE_STATS)) && tiled_applicable<LT>(d.ti, a.start_level)) {
            int rc = traverse_tiled<KIND, LQ, LT, I>(h, qleaves, n_query_total, d, a, flags, d_counts, d_contacts, capacity, num_contacts, st);
            if (rc != IBVH_ERR_UNSUPPORTED) return rc;
        }
    }
    return traverse_impl<KIND, true, LQ, LT, N, I>(h, qleaves, nullptr, nullptr, d, a, flags, d_counts, d_contacts, capacity, num_contacts, st);
}

// ---- BFS traversal driver (traverse_bfs.cuh) -----------------------------------------------------------------------------
// Level-synchronous: one kernel per BVTT level, one 8-byte read-back per level (the reference reads `dst_offsets[level]`
// the same way, traverse_single_gpu.jl:24), the two entry lists ping-pong between two library-owned buffers.
template <class F> int bfs_dispatch_volume(int kind, int fbytes, F&& f) {
    if (fbytes == 4) {
        if (kind == IBVH_BSPHERE) return f(Tag<BSphere<float>>{});
        if (kind == IBVH_BBOX) return f(Tag<BBox<float>>{});
    }
#ifdef IBVH_ENABLE_F64
    if (fbytes == 8) {
        if (kind == IBVH_BSPHERE) return f(Tag<BSphere<double>>{});
        if (kind == IBVH_BBOX) return f(Tag<BBox<double>>{});
    }
#endif
    return IBVH_ERR_UNSUPPORTED;
}
template <class F> int bfs_dispatch_index(int ibytes, F&& f) {
    if (ibytes == 4) return f(Tag<int32_t>{});
    if (ibytes == 8) return f(Tag<int64_t>{});
    return IBVH_ERR_UNSUPPORTED;
}
inline int bfs_node_fbytes(const ibvh_types_t& t) { return t.node_float_bytes ? t.node_float_bytes : t.float_bytes; }
// sizeof(BoundingVolume{V,I,M}) and offsetof(index) under natural alignment (bounding_volumes.jl:55-59)
inline void bfs_leaf_layout(const ibvh_types_t& t, uint32_t* stride, uint32_t* index_offset) {
    auto up = [](uint32_t v, uint32_t a) { return (v + a - 1) / a * a; };
    const uint32_t vol = (uint32_t)(t.leaf_kind == IBVH_BSPHERE ? 4 : 6) * (uint32_t)t.float_bytes;
    const uint32_t io = up(vol, (uint32_t)t.index_bytes);
    const uint32_t mo = up(io + (uint32_t)t.index_bytes, (uint32_t)t.morton_bytes);
    const uint32_t al = std::max((uint32_t)t.float_bytes, std::max((uint32_t)t.index_bytes, (uint32_t)t.morton_bytes));
    *stride = up(mo + (uint32_t)t.morton_bytes, al);
    *index_offset = io;
}
// BSphere nodes exist only over BSphere leaves (merge.jl); node floats are the leaf floats or Float32 over Float64 leaves
inline bool bfs_types_ok(const ibvh_types_t& t) {
    if (!types_ok(&t)) return false;
    if (t.node_kind == IBVH_BSPHERE && t.leaf_kind != IBVH_BSPHERE) return false;
    return true;
}
inline uint32_t bfs_vec(const void* base, uint32_t stride) {
    const uintptr_t m = reinterpret_cast<uintptr_t>(base) | (uintptr_t)stride;
    return (m & 15u) == 0 ? 16u : (m & 7u) == 0 ? 8u : 4u;
}
inline BfsSide bfs_node_side(const ibvh_bvh_t* b, const ibvh_tree_t& tree, int64_t level) {       // level < levels
    BfsSide s{};
    s.base = b->d_nodes;
    s.sub = (uint32_t)(skip_of_level(tree, level) + 1);
    s.stride = (uint32_t)((b->types.node_kind == IBVH_BSPHERE ? 4 : 6) * bfs_node_fbytes(b->types));
    s.child_first = (uint32_t)(int64_t(1) << level);
    s.child_nreal = (uint32_t)((int64_t(1) << level) - shr64(tree.virtual_leaves, tree.levels - (level + 1)));
    s.vec = bfs_vec(s.base, s.stride);
    return s;
}
inline BfsSide bfs_leaf_side(const ibvh_bvh_t* b, const ibvh_tree_t& tree) {
    BfsSide s{};
    uint32_t stride, io;
    bfs_leaf_layout(b->types, &stride, &io);
    s.base = b->d_leaves;
    s.sub = (uint32_t)(int64_t(1) << (tree.levels - 1));
    s.stride = stride;
    s.child_first = 0; s.child_nreal = 0;
    s.vec = bfs_vec(s.base, s.stride);
    return s;
}

struct BfsRun {
    ibvh_handle* h; cudaStream_t st;
    int cur = 0;                       // buffer that holds the current list
    unsigned long long count = 0;      // entries in it
    long long checks = 0;              // BVHTraversal.num_checks so far
    unsigned long long* d_counter() const { return (unsigned long long*)(h->d_small + kSmallBfs); }
    uint2* list(int which) const { return (uint2*)h->bfs_buf[which]; }
    int grid() const { return (int)std::min<unsigned long long>((count + kBfsTile - 1) / kBfsTile, (unsigned long long)h->sm_count * 8ull); }
    int reserve(int which, unsigned long long entries) {
        if (entries > (1ull << 40)) { h->set_error("BFS traversal: the BVTT would need more than 2^40 entries"); return IBVH_ERR_ALLOC; }
        return h->reserve_bfs(which, (size_t)(entries + 4) * sizeof(uint2));
    }
    int begin_step(unsigned long long bound) {
        int rc = reserve(cur ^ 1, bound);
        if (rc != IBVH_OK) return rc;
        IBVH_CUDA_TRY(h, cudaMemsetAsync(d_counter(), 0, 8, st));
        return IBVH_OK;
    }
    int read_counter(unsigned long long* out) {
        unsigned long long* hp = (unsigned long long*)h->h_pinned;
        IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, d_counter(), 8, cudaMemcpyDeviceToHost, st));
        IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
        *out = *hp;
        return IBVH_OK;
    }
    int end_step() {
        unsigned long long c;
        int rc = read_counter(&c);
        if (rc != IBVH_OK) return rc;
        count = c; checks += (long long)c; cur ^= 1;
        return IBVH_OK;
    }
};

// one node level: VA / VB are the volume types read on the two sides
template <int MODE, class VA, class VB>
int bfs_nodes_step(BfsRun& r, const BfsSide& sa, const BfsSide& sb, int self_checks) {
    if (r.count == 0) return IBVH_OK;
    const unsigned long long mult = (MODE == kBfsLeft || MODE == kBfsRight) ? 2ull : 4ull;
    int rc = r.begin_step(mult * r.count);
    if (rc != IBVH_OK) return rc;
    { ProfScope _ps(r.h, r.st, "bfs_nodes_kernel");
    bfs_nodes_kernel<MODE, VA, VB><<<r.grid(), kBfsThreads, 0, r.st>>>(r.list(r.cur), r.count, sa, sb, self_checks, r.list(r.cur ^ 1), r.d_counter());
    }
    IBVH_LAUNCH_CHECK(r.h, "bfs_nodes_kernel");
    return r.end_step();
}

// Leaf level of the single / pair traversals: `fused` = r holds the list of the LAST NODE levels (both trees one above their
// leaves) and bfs_last_kernel tests nodes and leaves in one pass; else r holds leaf pairs (traversal started at / reached a leaf
// level on one side first) and bfs_leaves_kernel tests them. r.checks comes back complete. MODE: kBfsSingle / kBfsBoth.
template <int MODE>
int bfs_leaf_level(ibvh_handle* h, cudaStream_t st, BfsRun& r, bool fused, const ibvh_bvh_t* b1, const ibvh_tree_t& t1, const ibvh_bvh_t* b2,
                   const ibvh_tree_t& t2, uint32_t flags, void* d_contacts, int64_t capacity, unsigned long long* total, long long* checks_out) {
    *total = 0;
    *checks_out = r.checks;
    if (r.count == 0) return IBVH_OK;
    const ibvh_types_t& ty = b1->types;
    const BfsSide ls1 = bfs_leaf_side(b1, t1), ls2 = bfs_leaf_side(b2, t2);
    uint32_t stride, io;
    bfs_leaf_layout(ty, &stride, &io);
    const int positions = (flags & IBVH_TRAVERSE_POSITIONS) ? 1 : 0;
    const unsigned long long cap = d_contacts ? (unsigned long long)std::max<int64_t>(capacity, 0) : 0ull;
    unsigned long long* d_cnt = r.d_counter();
    IBVH_CUDA_TRY(h, cudaMemsetAsync(d_cnt, 0, 16, st));
    int rc = bfs_dispatch_volume(ty.leaf_kind, ty.float_bytes, [&](auto vtag) -> int {
        using V = typename decltype(vtag)::type;
        return bfs_dispatch_index(ty.index_bytes, [&](auto itag) -> int {
            using I = typename decltype(itag)::type;
            if (!fused) {
                { ProfScope _ps(h, st, "bfs_leaves_kernel");
                bfs_leaves_kernel<MODE == kBfsSingle, V, I><<<r.grid(), kBfsThreads, 0, st>>>(r.list(r.cur), r.count, ls1, ls2, io, positions, (IndexPair<I>*)d_contacts, cap, d_cnt);
                }
                IBVH_LAUNCH_CHECK(h, "bfs_leaves_kernel");
                return IBVH_OK;
            }
            return bfs_dispatch_volume(ty.node_kind, bfs_node_fbytes(ty), [&](auto ntag) -> int {
                using N = typename decltype(ntag)::type;
                if constexpr (N::kind == IBVH_BSPHERE && V::kind != IBVH_BSPHERE) return (int)IBVH_ERR_ARGUMENT;
                else {
                    const BfsSide n1 = bfs_node_side(b1, t1, t1.levels - 1), n2 = bfs_node_side(b2, t2, t2.levels - 1);
                    { ProfScope _ps(h, st, "bfs_last_kernel");
                    bfs_last_kernel<MODE, MODE == kBfsSingle, N, V, I><<<r.grid(), kBfsThreads, 0, st>>>(r.list(r.cur), r.count, n1, n2, ls1, ls2, io, positions,
                                                                                                       (IndexPair<I>*)d_contacts, cap, d_cnt, d_cnt + 1);
                    }
                    IBVH_LAUNCH_CHECK(h, "bfs_last_kernel");
                    return IBVH_OK;
                }
            });
        });
    });
    if (rc != IBVH_OK) return rc;
    unsigned long long* hp = (unsigned long long*)h->h_pinned;
    IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, d_cnt, 16, cudaMemcpyDeviceToHost, st));
    IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
    *total = hp[0];
    if (fused) *checks_out = r.checks + (long long)hp[1];
    return IBVH_OK;
}

// What a BFS call that ended in IBVH_ERR_CAPACITY leaves behind, so that the repeat call with a larger cache1
// (IBVH_TRAVERSE_COUNTS_VALID) only redoes the leaf level.
inline bool bfs_resume(ibvh_handle* h, uint32_t flags, int kind, const void* l1, int64_t n1, const void* l2, int64_t n2, int64_t s1, int64_t s2, BfsRun* r, bool* fused = nullptr) {
    auto& p = h->bfs_pending;
    const bool ok = (flags & IBVH_TRAVERSE_COUNTS_VALID) && p.valid && p.kind == kind && p.leaves1 == l1 && p.n1 == n1 && p.leaves2 == l2 && p.n2 == n2 &&
                    p.start1 == s1 && p.start2 == s2;
    p.valid = false;
    if (!ok) return false;
    r->cur = p.cur; r->count = p.count; r->checks = p.checks;
    if (fused) *fused = p.fused;
    return true;
}
inline void bfs_remember(ibvh_handle* h, int kind, const void* l1, int64_t n1, const void* l2, int64_t n2, int64_t s1, int64_t s2, const BfsRun& r, bool fused = false) {
    auto& p = h->bfs_pending;
    p.valid = true; p.fused = fused; p.kind = kind; p.leaves1 = l1; p.n1 = n1; p.leaves2 = l2; p.n2 = n2; p.start1 = s1; p.start2 = s2;
    p.cur = r.cur; p.count = r.count; p.checks = r.checks;
}

int check_bvh(const ibvh_bvh_t* b, ibvh_tree_t* tree) {
    if (!b || !types_ok(&b->types)) return IBVH_ERR_ARGUMENT;
    int rc = make_tree(b->n, tree);
    if (rc != IBVH_OK) return rc;
    if (tree->levels > 32) return IBVH_ERR_ARGUMENT;                                 // traverse_single.jl:10
    if (!b->d_leaves) return IBVH_ERR_ARGUMENT;
    if (tree->real_nodes - tree->real_leaves > 0 && !b->d_nodes) return IBVH_ERR_ARGUMENT;
    return IBVH_OK;
}

void shard_range(const ibvh_traverse_params_t* p, int64_t nq, int64_t* begin, int64_t* count) {
    int64_t b = p->query_begin < 0 ? 0 : p->query_begin;
    if (b > nq) b = nq;
    int64_t c = p->query_count < 0 ? nq - b : p->query_count;
    if (b + c > nq) c = nq - b;
    *begin = b; *count = c;
}

}  // namespace IBVH_NS
using namespace IBVH_NS;

// =================================================================================================================
extern "C" {

#if defined(IBVH_PART_HOST) || defined(IBVH_PART_ALL)

int ibvh_version(void) { return IBVH_VERSION; }

const char* ibvh_status_string(int s) {
    switch (s) {
        case IBVH_OK: return "ok";
        case IBVH_ERR_ARGUMENT: return "ArgumentError";
        case IBVH_ERR_DOMAIN: return "DomainError";
        case IBVH_ERR_UNSUPPORTED: return "unsupported type combination";
        case IBVH_ERR_CUDA: return "CUDA error";
        case IBVH_ERR_CAPACITY: return "contacts capacity too small";
        case IBVH_ERR_ALLOC: return "workspace allocation failed";
        case IBVH_ERR_PEER: return "peer GPU did not arrive at the shard exchange";
        case IBVH_ERR_AGAIN: return "deferred traversal must be repeated (scratch pair lists were too small)";
    }
    return "unknown";
}
const char* ibvh_last_error(const ibvh_handle_t* h) { return h ? h->err : "null handle"; }

// ---- host-only --------------------------------------------------------------------------------------------
int ibvh_tree_shape(int64_t n, ibvh_tree_t* tree, int64_t* skips) {
    if (!tree) return IBVH_ERR_ARGUMENT;
    int rc = make_tree(n, tree);
    if (rc != IBVH_OK) return rc;
    if (skips) for (int64_t l = 1; l <= tree->levels; ++l) skips[l - 1] = skip_of_level(*tree, l);
    return IBVH_OK;
}
int64_t ibvh_memory_index(const ibvh_tree_t* t, int64_t implicit_index) {                 // implicit_tree.jl:128-148
    int64_t level = ilog2_floor_u64((uint64_t)implicit_index) + 1;
    return implicit_index - skip_of_level(*t, level);
}
int ibvh_level_indices(const ibvh_tree_t* t, int64_t level, int64_t* start, int64_t* stop) {   // :156-171
    if (level < 1 || level > t->levels) return IBVH_ERR_ARGUMENT;
    *start = ibvh_memory_index(t, int64_t(1) << (level - 1));
    int64_t nreal = (int64_t(1) << (level - 1)) - shr64(t->virtual_leaves, t->levels - level);
    *stop = *start + nreal - 1;
    return IBVH_OK;
}
int ibvh_isvirtual(const ibvh_tree_t* t, int64_t implicit_index) {                         // :191-199
    int64_t level = ilog2_floor_u64((uint64_t)implicit_index) + 1;
    int64_t level_first = int64_t(1) << (level - 1);
    int64_t nreal = level_first - shr64(t->virtual_leaves, t->levels - level);
    return implicit_index - level_first + 1 > nreal ? 1 : 0;
}
int ibvh_compute_build_level(int64_t levels, int is_float, int64_t ilevel, double flevel, int64_t* out) {   // build.jl:309-325
    if (!out) return IBVH_ERR_ARGUMENT;
    if (!is_float) {
        if (ilevel < 1 || ilevel > levels) return IBVH_ERR_ARGUMENT;
        *out = ilevel;
    } else {
        if (!(flevel >= 0.0 && flevel <= 1.0)) return IBVH_ERR_ARGUMENT;
        *out = (int64_t)nearbyint((double)levels + (double)(1 - levels) * flevel);   // round half to even
    }
    return IBVH_OK;
}
int64_t ibvh_volume_bytes(int32_t kind, int32_t float_bytes) {
    if (float_bytes != 4 && float_bytes != 8) return -1;
    if (kind == IBVH_BSPHERE) return 4 * float_bytes;
    if (kind == IBVH_BBOX) return 6 * float_bytes;
    return -1;
}
int64_t ibvh_leaf_bytes(const ibvh_types_t* t) {
    if (!types_ok(t)) return -1;
    int64_t v = ibvh_volume_bytes(t->leaf_kind, t->float_bytes);
    int64_t align = t->float_bytes;
    if (t->index_bytes > align) align = t->index_bytes;
    if (t->morton_bytes > align) align = t->morton_bytes;
    int64_t off = (v + t->index_bytes - 1) / t->index_bytes * t->index_bytes + t->index_bytes;
    off = (off + t->morton_bytes - 1) / t->morton_bytes * t->morton_bytes + t->morton_bytes;
    return (off + align - 1) / align * align;
}
int64_t ibvh_num_nodes(int64_t n) {
    ibvh_tree_t t;
    if (make_tree(n, &t) != IBVH_OK) return -1;
    return t.real_nodes - t.real_leaves;
}

// ---- handle -------------------------------------------------------------------------------------------------
int ibvh_create(ibvh_handle_t** out, int device) {
    if (!out) return IBVH_ERR_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return IBVH_ERR_CUDA; }
    ibvh_handle* h = new (std::nothrow) ibvh_handle();
    if (!h) return IBVH_ERR_ALLOC;
    h->device = device;
    h->cfg.parse();                      // environment knobs: read once here, never on the call path
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (!g.ok || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; cudaGetLastError(); return IBVH_ERR_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    if (cudaMalloc((void**)&h->d_small, ibvh_handle::kSmallBytes) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_pinned, ibvh_handle::kPinnedBytes) != cudaSuccess) {
        if (h->d_small) cudaFree(h->d_small);
        delete h; cudaGetLastError();
        return IBVH_ERR_ALLOC;
    }
    cudaMemset(h->d_small, 0, ibvh_handle::kSmallBytes);
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_defer, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(h->d_small); cudaFreeHost(h->h_pinned); delete h;
        return IBVH_ERR_CUDA;
    }
    *out = h;
    return IBVH_OK;
}
int ibvh_destroy(ibvh_handle_t* h) {
    if (!h) return IBVH_OK;
    DeviceGuard g(h->device);
    if (h->ws) cudaFree(h->ws);
    if (h->aux) cudaFree(h->aux);
    h->free_bfs();
    for (int k = 0; k < 2; ++k) if (h->sidecars[k].buf) cudaFree(h->sidecars[k].buf);
    if (h->d_small) cudaFree(h->d_small);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_defer) cudaEventDestroy(h->ev_defer);
    delete h;
    return IBVH_OK;
}
int64_t ibvh_workspace_bytes(const ibvh_handle_t* h) { return h ? (int64_t)h->ws_bytes : -1; }
int ibvh_release_workspace(ibvh_handle_t* h) {
    if (!h) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    if (h->ws) { IBVH_CUDA_TRY(h, cudaFree(h->ws)); h->ws = nullptr; h->ws_bytes = 0; }
    if (h->aux) { IBVH_CUDA_TRY(h, cudaFree(h->aux)); h->aux = nullptr; h->aux_bytes = 0; }
    h->free_bfs();
    for (int k = 0; k < 2; ++k) {
        ibvh_handle::Sidecar& sc = h->sidecars[k];
        if (sc.buf) { IBVH_CUDA_TRY(h, cudaFree(sc.buf)); sc.buf = nullptr; sc.bytes = 0; }
        sc.id = 0;
    }
    return IBVH_OK;
}

int ibvh_profile_enable(ibvh_handle_t* h, int on) {
    if (!h) return IBVH_ERR_ARGUMENT;
    h->prof_on = on != 0;
    h->prof_n = 0;
    return IBVH_OK;
}
int ibvh_profile_count(ibvh_handle_t* h) { return h ? h->prof_n : -1; }
int ibvh_profile_get(ibvh_handle_t* h, int i, char* name, int name_cap, float* ms) {
    if (!h || i < 0 || i >= h->prof_n || !ms) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    IBVH_CUDA_TRY(h, cudaEventSynchronize(h->prof_ev[i][1]));
    IBVH_CUDA_TRY(h, cudaEventElapsedTime(ms, h->prof_ev[i][0], h->prof_ev[i][1]));
    if (name && name_cap > 0) { snprintf(name, (size_t)name_cap, "%s", h->prof_name[i]); }
    return IBVH_OK;
}
int ibvh_profile_reset(ibvh_handle_t* h) { if (!h) return IBVH_ERR_ARGUMENT; h->prof_n = 0; return IBVH_OK; }

int ibvh_traverse_finish(ibvh_handle_t* h, int64_t* num_contacts) {
    if (!h || !num_contacts) return IBVH_ERR_ARGUMENT;
    ibvh_handle::Deferred& df = h->deferred;
    if (!df.active) { h->set_error("ibvh_traverse_finish: no deferred traversal outstanding"); return IBVH_ERR_ARGUMENT; }
    DeviceGuard g(h->device);
    df.active = false;
    cudaError_t e = cudaEventSynchronize(h->ev_defer);
    if (e != cudaSuccess) { h->set_cuda_error(e, "cudaEventSynchronize(deferred traversal)"); return IBVH_ERR_CUDA; }
    const unsigned long long* hp = (const unsigned long long*)h->h_pinned;
    bool overflow = false;
    double worst = 0.0;
    for (int l = 0; l < df.nl; ++l) {
        const unsigned long long c = hp[1 + l];
        if (c > df.cap[l]) overflow = true;
        const double r = (double)c / (double)df.nqg[l];
        if (l < df.nl - 1 && r > worst) worst = r;
    }
    if (overflow) { h->pyr_factor = worst * 1.15 + 2.0; *num_contacts = 0; return IBVH_ERR_AGAIN; }
    h->pyr_factor = worst * 1.25 + 4.0;
    long long box = df.top_pairs;
    for (int l = df.nl - 1; l >= 1; --l) box += (long long)hp[1 + l] * df.fan2;
    h->last_stats[0] = box; h->last_stats[1] = (long long)hp[1] * df.leaf2; h->last_stats[2] = (long long)hp[1]; h->last_stats[3] = df.nl;
    *num_contacts = (int64_t)hp[0];
    return *num_contacts > df.capacity ? IBVH_ERR_CAPACITY : IBVH_OK;
}

int ibvh_traverse_cancel(ibvh_handle_t* h) {
    if (!h) return IBVH_ERR_ARGUMENT;
    ibvh_handle::Deferred& df = h->deferred;
    if (!df.active) return IBVH_OK;
    DeviceGuard g(h->device);
    df.active = false;
    // the enqueued work still writes the handle's read-back slots: wait for it so that the next call owns them
    cudaError_t e = cudaEventSynchronize(h->ev_defer);
    if (e != cudaSuccess) { h->set_cuda_error(e, "cudaEventSynchronize(cancelled deferred traversal)"); return IBVH_ERR_CUDA; }
    return IBVH_OK;
}

uint64_t ibvh_last_build_id(ibvh_handle_t* h) { return h ? h->last_build_id : 0; }

int ibvh_peer_last_counts(ibvh_handle_t* h, int64_t* counts, int32_t world) {
    if (!h || !counts || world < 1 || world > IBVH_MAX_PEERS) return IBVH_ERR_ARGUMENT;
    for (int r = 0; r < world; ++r) counts[r] = h->last_peer_counts[r];
    return IBVH_OK;
}

int ibvh_last_traversal_stats(ibvh_handle_t* h, int64_t out[4]) {
    if (!h || !out) return IBVH_ERR_ARGUMENT;
    for (int k = 0; k < 4; ++k) out[k] = h->last_stats[k];
    return IBVH_OK;
}

#endif  // IBVH_PART_HOST

#if defined(IBVH_PART_BUILD) || defined(IBVH_PART_ALL)
int64_t ibvh_workspace_query(const ibvh_types_t* types, int64_t n) {
    if (!types_ok(types) || n < 1) return -1;
    int64_t out = -1;
    dispatch_leaf(*types, [&](auto tag) -> int { using L = typename decltype(tag)::type; out = (int64_t)build_workspace_bytes<L>(n, true); return IBVH_OK; });
    return out;
}
// ---- build stages ---------------------------------------------------------------------------------------------
int ibvh_wrap(ibvh_handle_t* h, const void* d_volumes, int64_t n, const ibvh_types_t* types, void* d_leaves, void* stream) {
    if (!h || !types_ok(types) || n < 0) return IBVH_ERR_ARGUMENT;
    if (n == 0) return IBVH_OK;
    if (!d_volumes || !d_leaves) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(*types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type;
        { ProfScope _ps(h, st, "wrap_kernel");
        wrap_kernel<L><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const typename L::vol_t*)d_volumes, n, (L*)d_leaves);
        }
        IBVH_LAUNCH_CHECK(h, "wrap_kernel");
        return (int)IBVH_OK;
    });
}

int ibvh_volumes_from_triangles(ibvh_handle_t* h, const void* d_triangles, int64_t n, int32_t kind, int32_t float_bytes, void* d_volumes, void* stream) {
    if (!h || n < 0 || (kind != IBVH_BSPHERE && kind != IBVH_BBOX) || (float_bytes != 4 && float_bytes != 8)) return IBVH_ERR_ARGUMENT;
    if (n == 0) return IBVH_OK;
    if (!d_triangles || !d_volumes) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    { ProfScope _ps(h, st, "triangles_kernel");
    if (float_bytes == 4) {
        if (kind == IBVH_BSPHERE) triangles_kernel<BSphere<float>><<<blocks, 256, 0, st>>>((const float*)d_triangles, n, (BSphere<float>*)d_volumes);
        else triangles_kernel<BBox<float>><<<blocks, 256, 0, st>>>((const float*)d_triangles, n, (BBox<float>*)d_volumes);
    } else {
        if (kind == IBVH_BSPHERE) triangles_kernel<BSphere<double>><<<blocks, 256, 0, st>>>((const double*)d_triangles, n, (BSphere<double>*)d_volumes);
        else triangles_kernel<BBox<double>><<<blocks, 256, 0, st>>>((const double*)d_triangles, n, (BBox<double>*)d_volumes);
    }
    }
    IBVH_LAUNCH_CHECK(h, "triangles_kernel");
    return IBVH_OK;
}

int ibvh_morton_encode(ibvh_handle_t* h, void* d_leaves, int64_t n, const ibvh_types_t* types, int compute_extrema,
                       const double* mins, const double* maxs, double* out_mins, double* out_maxs, void* stream) {
    if (!h || !types_ok(types) || n < 0) return IBVH_ERR_ARGUMENT;
    if (n == 0) return IBVH_OK;                                              // default.jl:49
    if (!d_leaves) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(*types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type; using M = typename L::mor_t; using T = typename L::value_type;
        constexpr int P = radix_passes<M>();
        int rc = h->reserve(ibvh_handle::padded((size_t)n * sizeof(M)) + ibvh_handle::padded(P * radix_bins<M>() * 4) + 4096);
        if (rc != IBVH_OK) return rc;
        h->reset();
        M* keys = h->alloc<M>(n);
        uint32_t* hist = h->alloc<uint32_t>(P * radix_bins<M>());
        rc = launch_encode<L, L, false>(h, (const L*)d_leaves, n, compute_extrema, mins, maxs, keys, nullptr, hist, (uint32_t*)(h->d_small + kSmallTickets), st);
        if (rc != IBVH_OK) return rc;
        { ProfScope _ps(h, st, "scatter_morton_kernel");
        scatter_morton_kernel<L><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((L*)d_leaves, n, keys);
        }
        IBVH_LAUNCH_CHECK(h, "scatter_morton_kernel");
        return read_used_bounds<T>(h, out_mins, out_maxs, st);
    });
}

int ibvh_sort_leaves(ibvh_handle_t* h, void* d_leaves, int64_t n, const ibvh_types_t* types, void* stream) {
    if (!h || !types_ok(types) || n < 0) return IBVH_ERR_ARGUMENT;
    if (n == 0) return IBVH_OK;
    if (!d_leaves) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(*types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type; using M = typename L::mor_t; using T = typename L::value_type;
        constexpr int P = radix_passes<M>();
        SortScratch s;
        int rc = carve_sort_scratch<L>(h, n, true, &s, st);
        if (rc != IBVH_OK) return rc;
        { ProfScope _ps(h, st, "init_build_kernel");
        init_build_kernel<T><<<8, 256, 0, st>>>((typename OrdOf<T>::type*)(h->d_small + kSmallBounds), s.hist, P * radix_bins<M>(), s.tickets, 16);
        }
        IBVH_LAUNCH_CHECK(h, "init_build_kernel");
        { ProfScope _ps(h, st, "extract_keys_kernel");
        extract_keys_kernel<L><<<grid_for(n, 256, 4, h->sm_count * 8), 256, 0, st>>>((const L*)d_leaves, n, (M*)s.keysA, (L*)s.copy, s.hist);
        }
        IBVH_LAUNCH_CHECK(h, "extract_keys_kernel");
        M* keys_sorted; uint32_t* perm;
        rc = sort_pairs<M>(h, (M*)s.keysA, (M*)s.keysB, s.valsA, s.valsB, n, s.hist, s.lookback, s.tickets, st, &keys_sorted, &perm);
        if (rc != IBVH_OK) return rc;
        ibvh_tree_t tree; make_tree(n, &tree);
        TreeInfo ti = make_tree_info(tree);
        // gather only: stop_level above every level => no node is written
        return launch_gather_merge<L, L, BBox<T>, true>(h, (const L*)s.copy, perm, keys_sorted, (L*)d_leaves, (BBox<T>*)nullptr, ti, INT_MAX, st);
    });
}

int ibvh_aggregate(ibvh_handle_t* h, const void* d_leaves, int64_t n, const ibvh_types_t* types, void* d_nodes, int64_t built_level, void* stream) {
    if (!h || !types_ok(types)) return IBVH_ERR_ARGUMENT;
    ibvh_tree_t tree;
    int rc = make_tree(n, &tree);
    if (rc != IBVH_OK) return rc;
    if (built_level < 1 || built_level > tree.levels) return IBVH_ERR_ARGUMENT;
    if (tree.real_nodes < 2) return IBVH_OK;                               // build.jl:266
    if (!d_leaves || !d_nodes) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(*types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type;
        return dispatch_node<L>(*types, [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type;
            TreeInfo ti = make_tree_info(tree);
            int stop_level = (int)(built_level < tree.levels - 1 ? built_level : tree.levels - 1);
            int r = launch_gather_merge<L, L, N, false>(h, (const L*)nullptr, nullptr, nullptr, (L*)d_leaves, (N*)d_nodes, ti, stop_level, st);
            if (r != IBVH_OK) return r;
            return merge_upper_levels<N>(h, (N*)d_nodes, ti, stop_level, st);
        });
    });
}

int ibvh_build(ibvh_handle_t* h, const void* d_volumes, void* d_leaves, int64_t n, const ibvh_types_t* types, void* d_nodes,
               int64_t built_level, int compute_extrema, const double* mins, const double* maxs, void* stream) {
    if (!h || !types_ok(types)) return IBVH_ERR_ARGUMENT;
    if (n < 1) return IBVH_ERR_DOMAIN;                                     // implicit_tree.jl:78-80
    if (!d_leaves) return IBVH_ERR_ARGUMENT;
    if (ibvh_num_nodes(n) > 0 && !d_nodes) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(*types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type;
        return dispatch_node<L>(*types, [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type;
            return build_impl<L, N>(h, d_volumes, d_leaves, n, d_nodes, built_level, compute_extrema, mins, maxs, st);
        });
    });
}

// ---- reference-shaped proxy build (reference_shaped.cuh): what BVH(...) launches through AcceleratedKernels ----------
int ibvh_build_reference_shaped(ibvh_handle_t* h, const void* d_volumes, void* d_leaves, int64_t n, const ibvh_types_t* types,
                                void* d_nodes, int64_t built_level, void* stream) {
    namespace rs = ibvh::refshaped;
    if (!h || !types_ok(types)) return IBVH_ERR_ARGUMENT;
    if (n < 1) return IBVH_ERR_DOMAIN;
    if (!d_volumes || !d_leaves) return IBVH_ERR_ARGUMENT;
    if (!(types->leaf_kind == IBVH_BSPHERE && types->node_kind == IBVH_BBOX && types->float_bytes == 4 && types->index_bytes == 4 &&
          types->morton_bytes == 4 && (types->node_float_bytes == 0 || types->node_float_bytes == 4))) {
        h->set_error("the reference-shaped proxy build exists for BSphere{Float32} / Int32 / UInt32 / BBox{Float32} only");
        return IBVH_ERR_UNSUPPORTED;
    }
    ibvh_tree_t tree;
    int rc = make_tree(n, &tree);
    if (rc != IBVH_OK) return rc;
    if (built_level < 1 || built_level > tree.levels) return IBVH_ERR_ARGUMENT;
    if (ibvh_num_nodes(n) > 0 && !d_nodes) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    rs::RLeaf* leaves = (rs::RLeaf*)d_leaves;
    rs::RNode* nodes = (rs::RNode*)d_nodes;
    const int rblocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
    rc = h->reserve(ibvh_handle::padded((size_t)n * sizeof(rs::RLeaf)) + ibvh_handle::padded((size_t)rblocks * 3 * 4) + 4096);
    if (rc != IBVH_OK) return rc;
    h->reset();
    rs::RLeaf* tmp = h->alloc<rs::RLeaf>(n);
    float* partial = h->alloc<float>((size_t)rblocks * 3);
    if (!tmp || !partial) { h->set_error("workspace carve failed"); return IBVH_ERR_ALLOC; }
    const unsigned nb = (unsigned)((n + 255) / 256);
    // wrap_bounding_volumes (build.jl:340)
    { ProfScope _ps(h, st, "ref_wrap_kernel");
    wrap_kernel<rs::RLeaf><<<nb, 256, 0, st>>>((const BSphere<float>*)d_volumes, n, leaves);
    }
    IBVH_LAUNCH_CHECK(h, "ref_wrap_kernel");
    // _compute_extrema: two mapreduce passes, each ending in a scalar read-back on the host (morton/utils.jl:24-44)
    float* d_ext = (float*)(h->d_small + kSmallBoundsF + 256);
    float* h_ext = (float*)(h->h_pinned + 2048 + 256);
    { ProfScope _ps(h, st, "ref_extrema_kernel");
    rs::extrema_partial_kernel<false><<<rblocks, 256, 0, st>>>(leaves, n, partial);
    rs::extrema_final_kernel<false><<<1, 32, 0, st>>>(partial, rblocks, d_ext);
    }
    IBVH_LAUNCH_CHECK(h, "ref_extrema_kernel(min)");
    IBVH_CUDA_TRY(h, cudaMemcpyAsync(h_ext, d_ext, 12, cudaMemcpyDeviceToHost, st));
    IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
    { ProfScope _ps(h, st, "ref_extrema_kernel");
    rs::extrema_partial_kernel<true><<<rblocks, 256, 0, st>>>(leaves, n, partial);
    rs::extrema_final_kernel<true><<<1, 32, 0, st>>>(partial, rblocks, d_ext + 3);
    }
    IBVH_LAUNCH_CHECK(h, "ref_extrema_kernel(max)");
    IBVH_CUDA_TRY(h, cudaMemcpyAsync(h_ext + 3, d_ext + 3, 12, cudaMemcpyDeviceToHost, st));
    IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
    rs::Bounds6 b;
    for (int k = 0; k < 3; ++k) { b.mins[k] = h_ext[k]; b.maxs[k] = h_ext[3 + k]; }
    pad_extrema(b.mins, b.maxs);                                   // on the host, as bounding_volumes_extrema does (morton/utils.jl:63-69)
    // _morton_encode! (default.jl:66)
    { ProfScope _ps(h, st, "ref_encode_kernel");
    rs::encode_kernel<<<nb, 256, 0, st>>>(leaves, n, b);
    }
    IBVH_LAUNCH_CHECK(h, "ref_encode_kernel");
    // AK.sort!(leaves, by = morton): struct-moving merge sort (build.jl:248-253)
    rs::RLeaf* cur = tmp; rs::RLeaf* other = leaves;
    { ProfScope _ps(h, st, "ref_block_sort_kernel");
    rs::block_sort_kernel<<<(unsigned)((n + rs::kBlockSort - 1) / rs::kBlockSort), rs::kBlockSort, 0, st>>>(leaves, tmp, n);
    }
    IBVH_LAUNCH_CHECK(h, "ref_block_sort_kernel");
    for (int64_t width = rs::kBlockSort; width < n; width *= 2) {
        const int64_t threads = (n + rs::kMergePerThread - 1) / rs::kMergePerThread;
        { ProfScope _ps(h, st, "ref_merge_pass_kernel");
        rs::merge_pass_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(cur, other, n, width);
        }
        IBVH_LAUNCH_CHECK(h, "ref_merge_pass_kernel");
        rs::RLeaf* t = cur; cur = other; other = t;
    }
    if (cur != leaves) IBVH_CUDA_TRY(h, cudaMemcpyAsync(leaves, cur, (size_t)n * sizeof(rs::RLeaf), cudaMemcpyDeviceToDevice, st));
    // aggregate_oibvh!: one launch per level (build.jl:366-523)
    if (tree.real_nodes >= 2) {
        TreeInfo ti = make_tree_info(tree);
        int lvl = (int)tree.levels - 1;
        { ProfScope _ps(h, st, "ref_aggregate_kernel");
        rs::aggregate_last_level_kernel<<<(unsigned)((ti.level_nreal[lvl] + 255) / 256), 256, 0, st>>>(leaves, nodes, n, ti.level_start[lvl], ti.level_nreal[lvl]);
        }
        IBVH_LAUNCH_CHECK(h, "ref_aggregate_kernel(last level)");
        for (lvl -= 1; lvl >= built_level && lvl >= 1; --lvl) {
            { ProfScope _ps(h, st, "ref_aggregate_kernel");
            rs::aggregate_level_kernel<<<(unsigned)((ti.level_nreal[lvl] + 255) / 256), 256, 0, st>>>(nodes, ti.level_start[lvl], ti.level_nreal[lvl], ti.level_start[lvl + 1], ti.level_nreal[lvl + 1]);
            }
            IBVH_LAUNCH_CHECK(h, "ref_aggregate_kernel");
        }
    }
    return IBVH_OK;
}

#endif  // IBVH_PART_BUILD

// ---- traversals ---------------------------------------------------------------------------------------------------
#if defined(IBVH_PART_SINGLE) || defined(IBVH_PART_ALL)
int ibvh_traverse_single(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const ibvh_traverse_params_t* p, void* d_counts, void* d_contacts,
                         int64_t capacity, int64_t* num_contacts, void* stream) {
    if (!h || !p || !num_contacts) return IBVH_ERR_ARGUMENT;
    if (h->deferred.active) { h->set_error("a deferred traversal is outstanding on this handle: call ibvh_traverse_finish (or ibvh_traverse_cancel) first"); return IBVH_ERR_ARGUMENT; }
    for (int k = 0; k < 4; ++k) h->last_stats[k] = 0;
    ibvh_tree_t tree;
    int rc = check_bvh(bvh, &tree);
    if (rc != IBVH_OK) return rc;
    if (!(bvh->built_level <= p->start_level && p->start_level <= tree.levels)) return IBVH_ERR_ARGUMENT;   // traverse_single.jl:9-11
    if (d_contacts && (reinterpret_cast<uintptr_t>(d_contacts) % (2u * (unsigned)bvh->types.index_bytes)) != 0) { h->set_error("d_contacts must be aligned to the size of one IndexPair (2 * index_bytes)"); return IBVH_ERR_ARGUMENT; }
    *num_contacts = 0;
    if (tree.real_nodes <= 1) return IBVH_OK;                               // traverse_single.jl:17-21
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(bvh->types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type; using I = typename L::idx_t;
        return dispatch_node<L>(bvh->types, [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type;
            DBvh<L, N> d{(const L*)bvh->d_leaves, (const N*)bvh->d_nodes, make_tree_info(tree)};
            TraverseArgs a{};
            shard_range(p, bvh->n, &a.q_begin, &a.q_count);
            a.start_level = (int32_t)p->start_level;
            a.flip = 0;
            a.peer = p->peer;
            a.positions = (p->flags & IBVH_TRAVERSE_POSITIONS) ? 1 : 0;
            a.t_build_id = a.q_build_id = bvh->build_id;
            a.t_built_level = bvh->built_level;
            return traverse_leaf_queries<kSingle, L, L, N, I>(h, d.leaves, bvh->n, d, bvh->built_level, a, p->flags, d_counts, d_contacts, capacity, num_contacts, st);
        });
    });
}

#endif  // IBVH_PART_SINGLE

#if defined(IBVH_PART_PAIR) || defined(IBVH_PART_ALL)
int ibvh_traverse_pair(ibvh_handle_t* h, const ibvh_bvh_t* queries, const ibvh_bvh_t* target, const ibvh_traverse_params_t* p,
                       void* d_counts, void* d_contacts, int64_t capacity, int64_t* num_contacts, void* stream) {
    if (!h || !p || !num_contacts || !queries || !target) return IBVH_ERR_ARGUMENT;
    if (h->deferred.active) { h->set_error("a deferred traversal is outstanding on this handle: call ibvh_traverse_finish (or ibvh_traverse_cancel) first"); return IBVH_ERR_ARGUMENT; }
    for (int k = 0; k < 4; ++k) h->last_stats[k] = 0;
    ibvh_tree_t tq, tt;
    if (!types_ok(&queries->types) || queries->n < 1 || !queries->d_leaves) return IBVH_ERR_ARGUMENT;
    int rc = make_tree(queries->n, &tq);
    if (rc != IBVH_OK) return rc;
    rc = check_bvh(target, &tt);
    if (rc != IBVH_OK) return rc;
    if (tq.levels > 32) return IBVH_ERR_ARGUMENT;
    // both trees must share leaf / index / node types (traverse_pair.jl:50-52)
    const ibvh_types_t &a1 = queries->types, &a2 = target->types;
    if (a1.leaf_kind != a2.leaf_kind || a1.float_bytes != a2.float_bytes || a1.index_bytes != a2.index_bytes ||
        a1.morton_bytes != a2.morton_bytes || a1.node_kind != a2.node_kind || a1.node_float_bytes != a2.node_float_bytes) return IBVH_ERR_ARGUMENT;
    if (!(target->built_level <= p->start_level && p->start_level <= tt.levels)) return IBVH_ERR_ARGUMENT;   // traverse_pair.jl:11-12
    if (d_contacts && (reinterpret_cast<uintptr_t>(d_contacts) % (2u * (unsigned)target->types.index_bytes)) != 0) { h->set_error("d_contacts must be aligned to the size of one IndexPair (2 * index_bytes)"); return IBVH_ERR_ARGUMENT; }
    *num_contacts = 0;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(target->types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type; using I = typename L::idx_t;
        return dispatch_node<L>(target->types, [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type;
            DBvh<L, N> d{(const L*)target->d_leaves, (const N*)target->d_nodes, make_tree_info(tt)};
            TraverseArgs a{};
            shard_range(p, queries->n, &a.q_begin, &a.q_count);
            a.start_level = (int32_t)p->start_level;
            a.flip = p->flip ? 1 : 0;
            a.peer = p->peer;
            a.positions = (p->flags & IBVH_TRAVERSE_POSITIONS) ? 1 : 0;
            a.t_build_id = target->build_id; a.q_build_id = queries->build_id;
            a.t_built_level = target->built_level;
            return traverse_leaf_queries<kPair, L, L, N, I>(h, (const L*)queries->d_leaves, queries->n, d, target->built_level, a, p->flags, d_counts, d_contacts, capacity, num_contacts, st);
        });
    });
}

#endif  // IBVH_PART_PAIR

#if defined(IBVH_PART_RAYS) || defined(IBVH_PART_ALL)
int ibvh_traverse_rays(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const void* d_points, const void* d_directions, int64_t nrays,
                       const ibvh_traverse_params_t* p, void* d_counts, void* d_contacts, int64_t capacity, int64_t* num_contacts, void* stream) {
    if (!h || !p || !num_contacts || nrays < 0) return IBVH_ERR_ARGUMENT;
    // (the ray traversal's read-back uses the same pinned / counter slots as an outstanding deferred traversal)
    if (h->deferred.active) { h->set_error("a deferred traversal is outstanding on this handle: call ibvh_traverse_finish (or ibvh_traverse_cancel) first"); return IBVH_ERR_ARGUMENT; }
    for (int k = 0; k < 4; ++k) h->last_stats[k] = 0;
    ibvh_tree_t tree;
    int rc = check_bvh(bvh, &tree);
    if (rc != IBVH_OK) return rc;
    if (!(bvh->built_level <= p->start_level && p->start_level <= tree.levels)) return IBVH_ERR_ARGUMENT;   // leaf_vs_tree.jl:11-15
    if (d_contacts && (reinterpret_cast<uintptr_t>(d_contacts) % (2u * (unsigned)bvh->types.index_bytes)) != 0) { h->set_error("d_contacts must be aligned to the size of one IndexPair (2 * index_bytes)"); return IBVH_ERR_ARGUMENT; }
    *num_contacts = 0;
    if (nrays == 0 && !p->peer) return IBVH_OK;                               // leaf_vs_tree.jl:22-26
    if (nrays > 0 && (!d_points || !d_directions)) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_leaf(bvh->types, [&](auto tag) -> int {
        using L = typename decltype(tag)::type; using I = typename L::idx_t; using T = typename L::value_type;
        // the reference's ray tests take the volume and the ray in ONE float type (isintersection.jl:1-5, 36-40): a tree whose
        // node float type differs from its leaves' has no ray traversal there either (MethodError) -> ArgumentError here
        if (bvh->types.node_float_bytes != 0 && bvh->types.node_float_bytes != bvh->types.float_bytes) return (int)IBVH_ERR_ARGUMENT;
        return dispatch_node_kind<L, T>(bvh->types, [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type;
            DBvh<L, N> d{(const L*)bvh->d_leaves, (const N*)bvh->d_nodes, make_tree_info(tree)};
            TraverseArgs a{};
            shard_range(p, nrays, &a.q_begin, &a.q_count);
            a.start_level = (int32_t)p->start_level;
            a.flip = 0;
            a.id_base = p->id_base;
            a.peer = p->peer;
            a.positions = (p->flags & IBVH_TRAVERSE_POSITIONS) ? 1 : 0;
            return traverse_impl<kRays, false, L, L, N, I>(h, (const L*)nullptr, (const T*)d_points, (const T*)d_directions, d, a, p->flags,
                                                           d_counts, d_contacts, capacity, num_contacts, st);
        });
    });
}

#endif  // IBVH_PART_RAYS

#if defined(IBVH_PART_BFS) || defined(IBVH_PART_ALL)
int64_t ibvh_bfs_default_start_level(int64_t levels, int64_t built_level) {     // breadth_first/breadth_first.jl:4-6
    const int64_t half = levels / 2;
    return half > built_level ? half : built_level;
}

// common argument checks of the three BFS entry points
static int bfs_common_checks(ibvh_handle_t* h, const ibvh_bvh_t* b, void* d_contacts, int64_t* num_contacts, int64_t* num_checks) {
    if (!h || !b || !num_contacts) return IBVH_ERR_ARGUMENT;
    if (h->deferred.active) { h->set_error("a deferred traversal is outstanding on this handle: call ibvh_traverse_finish (or ibvh_traverse_cancel) first"); return IBVH_ERR_ARGUMENT; }
    if (!bfs_types_ok(b->types)) return IBVH_ERR_ARGUMENT;
    if (d_contacts && (reinterpret_cast<uintptr_t>(d_contacts) % (2u * (unsigned)b->types.index_bytes)) != 0) { h->set_error("d_contacts must be aligned to the size of one IndexPair (2 * index_bytes)"); return IBVH_ERR_ARGUMENT; }
    *num_contacts = 0;
    if (num_checks) *num_checks = 0;
    return IBVH_OK;
}

int ibvh_traverse_bfs_single(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const ibvh_traverse_params_t* p, void* d_contacts, int64_t capacity,
                             int64_t* num_contacts, int64_t* num_checks, void* stream) {
    if (!p) return IBVH_ERR_ARGUMENT;
    int rc = bfs_common_checks(h, bvh, d_contacts, num_contacts, num_checks);
    if (rc != IBVH_OK) return rc;
    ibvh_tree_t tree;
    rc = check_bvh(bvh, &tree);
    if (rc != IBVH_OK) return rc;
    if (!(bvh->built_level <= p->start_level && p->start_level <= tree.levels)) return IBVH_ERR_ARGUMENT;   // breadth_first/traverse_single.jl:10
    if (tree.real_nodes <= 1) return IBVH_OK;                                                             // :17-21
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    BfsRun r{h, st};
    const int64_t levels = tree.levels, sl = p->start_level;
    bool fused = sl < levels;                 // the last node level runs fused with the leaf level (a start at the leaf level has none)
    if (!bfs_resume(h, p->flags, 1, bvh->d_leaves, bvh->n, bvh->d_leaves, bvh->n, sl, sl, &r, &fused)) {
        // initial_bvtt, traverse_single.jl:69-157: every pair (i <= j) of the real nodes of the start level
        const int64_t first = int64_t(1) << (sl - 1);
        const int64_t nreal = first - shr64(tree.virtual_leaves, levels - sl);
        const bool with_self = sl != levels;
        const unsigned long long n0 = (unsigned long long)nreal * (unsigned long long)(nreal - 1) / 2 + (with_self ? (unsigned long long)nreal : 0ull);
        rc = r.reserve(0, n0);
        if (rc != IBVH_OK) return rc;
        if (n0) {
            const unsigned long long threads = (unsigned long long)nreal * (unsigned long long)nreal;
            { ProfScope _ps(h, st, "bfs_init_kernel");
            bfs_init_single_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(r.list(0), (uint32_t)first, (uint32_t)nreal, with_self ? 1 : 0);
            }
            IBVH_LAUNCH_CHECK(h, "bfs_init_single_kernel");
        }
        r.cur = 0; r.count = n0; r.checks = (long long)n0;
        rc = bfs_dispatch_volume(bvh->types.node_kind, bfs_node_fbytes(bvh->types), [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type;
            for (int64_t level = sl; level < levels - 1; ++level) {     // (level levels - 1 is the fused one; self-checks sprout on all of these)
                const BfsSide s = bfs_node_side(bvh, tree, level);
                int rc2 = bfs_nodes_step<kBfsSingle, N, N>(r, s, s, 1);
                if (rc2 != IBVH_OK) return rc2;
            }
            return IBVH_OK;
        });
        if (rc != IBVH_OK) return rc;
    }
    // traverse_leaves!, traverse_single_gpu.jl:114-211 (+ the last traverse_nodes! level when fused)
    unsigned long long total = 0;
    long long checks = 0;
    rc = bfs_leaf_level<kBfsSingle>(h, st, r, fused, bvh, tree, bvh, tree, p->flags, d_contacts, capacity, &total, &checks);
    if (rc != IBVH_OK) return rc;
    if (num_checks) *num_checks = checks;
    *num_contacts = (int64_t)total;
    if ((d_contacts && (int64_t)total > capacity) || (!d_contacts && total > 0)) bfs_remember(h, 1, bvh->d_leaves, bvh->n, bvh->d_leaves, bvh->n, sl, sl, r, fused);
    return d_contacts && (int64_t)total > capacity ? IBVH_ERR_CAPACITY : IBVH_OK;
}

int ibvh_traverse_bfs_pair(ibvh_handle_t* h, const ibvh_bvh_t* bvh1, const ibvh_bvh_t* bvh2, int64_t start_level1, int64_t start_level2,
                           uint32_t flags, void* d_contacts, int64_t capacity, int64_t* num_contacts, int64_t* num_checks, void* stream) {
    if (!bvh2) return IBVH_ERR_ARGUMENT;
    int rc = bfs_common_checks(h, bvh1, d_contacts, num_contacts, num_checks);
    if (rc != IBVH_OK) return rc;
    ibvh_tree_t t1, t2;
    rc = check_bvh(bvh1, &t1);
    if (rc != IBVH_OK) return rc;
    rc = check_bvh(bvh2, &t2);
    if (rc != IBVH_OK) return rc;
    const ibvh_types_t &a1 = bvh1->types, &a2 = bvh2->types;
    if (a1.leaf_kind != a2.leaf_kind || a1.float_bytes != a2.float_bytes || a1.index_bytes != a2.index_bytes ||
        a1.morton_bytes != a2.morton_bytes || a1.node_kind != a2.node_kind || bfs_node_fbytes(a1) != bfs_node_fbytes(a2)) return IBVH_ERR_ARGUMENT;
    if (!(bvh1->built_level <= start_level1 && start_level1 <= t1.levels)) return IBVH_ERR_ARGUMENT;     // breadth_first/traverse_pair.jl:10-11
    if (!(bvh2->built_level <= start_level2 && start_level2 <= t2.levels)) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    BfsRun r{h, st};
    bool fused = false;
    if (!bfs_resume(h, flags, 2, bvh1->d_leaves, bvh1->n, bvh2->d_leaves, bvh2->n, start_level1, start_level2, &r, &fused)) {
        // initial_bvtt, traverse_pair.jl:161-221: the product of the real nodes of the two start levels
        const int64_t f1 = int64_t(1) << (start_level1 - 1), f2 = int64_t(1) << (start_level2 - 1);
        const int64_t n1 = f1 - shr64(t1.virtual_leaves, t1.levels - start_level1), n2 = f2 - shr64(t2.virtual_leaves, t2.levels - start_level2);
        const unsigned long long n0 = (unsigned long long)n1 * (unsigned long long)n2;
        rc = r.reserve(0, n0);
        if (rc != IBVH_OK) return rc;
        { ProfScope _ps(h, st, "bfs_init_kernel");
        bfs_init_product_kernel<<<(unsigned)std::min<unsigned long long>((n0 + 255) / 256, 1u << 20), 256, 0, st>>>(r.list(0), (uint32_t)f1, (unsigned long long)n1, (uint32_t)f2, (unsigned long long)n2);
        }
        IBVH_LAUNCH_CHECK(h, "bfs_init_product_kernel");
        r.cur = 0; r.count = n0; r.checks = (long long)n0;
        rc = bfs_dispatch_volume(a1.node_kind, bfs_node_fbytes(a1), [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type;
            return bfs_dispatch_volume(a1.leaf_kind, a1.float_bytes, [&](auto vtag) -> int {
                using V = typename decltype(vtag)::type;
                if constexpr (N::kind == IBVH_BSPHERE && V::kind != IBVH_BSPHERE) return (int)IBVH_ERR_ARGUMENT;
                else {
                    // the stages of traverse_pair.jl:39-140
                    int64_t l1 = start_level1, l2 = start_level2;
                    int rc2 = IBVH_OK;
                    while (l1 < t1.levels - 1 && l2 < t2.levels - 1) {
                        if ((rc2 = bfs_nodes_step<kBfsBoth, N, N>(r, bfs_node_side(bvh1, t1, l1), bfs_node_side(bvh2, t2, l2), 0)) != IBVH_OK) return rc2;
                        ++l1; ++l2;
                    }
                    while (l1 < t1.levels - 1 && l2 == t2.levels - 1) {
                        if ((rc2 = bfs_nodes_step<kBfsLeft, N, N>(r, bfs_node_side(bvh1, t1, l1), bfs_node_side(bvh2, t2, l2), 0)) != IBVH_OK) return rc2;
                        ++l1;
                    }
                    while (l2 < t2.levels - 1 && l1 == t1.levels - 1) {
                        if ((rc2 = bfs_nodes_step<kBfsRight, N, N>(r, bfs_node_side(bvh1, t1, l1), bfs_node_side(bvh2, t2, l2), 0)) != IBVH_OK) return rc2;
                        ++l2;
                    }
                    while (l2 == t2.levels && l1 < t1.levels) {          // bvh2 already at its leaves: node (bvh1) against leaf volume (bvh2)
                        if ((rc2 = bfs_nodes_step<kBfsLeft, N, V>(r, bfs_node_side(bvh1, t1, l1), bfs_leaf_side(bvh2, t2), 0)) != IBVH_OK) return rc2;
                        ++l1;
                    }
                    while (l1 == t1.levels && l2 < t2.levels) {
                        if ((rc2 = bfs_nodes_step<kBfsRight, V, N>(r, bfs_leaf_side(bvh1, t1), bfs_node_side(bvh2, t2, l2), 0)) != IBVH_OK) return rc2;
                        ++l2;
                    }
                    // both one above their leaves: that node level runs fused with the leaf level (bfs_leaf_level)
                    fused = (l1 == t1.levels - 1 && l2 == t2.levels - 1);
                    return IBVH_OK;
                }
            });
        });
        if (rc != IBVH_OK) return rc;
    }
    unsigned long long total = 0;
    long long checks = 0;
    rc = bfs_leaf_level<kBfsBoth>(h, st, r, fused, bvh1, t1, bvh2, t2, flags, d_contacts, capacity, &total, &checks);
    if (rc != IBVH_OK) return rc;
    if (num_checks) *num_checks = checks;
    *num_contacts = (int64_t)total;
    if ((d_contacts && (int64_t)total > capacity) || (!d_contacts && total > 0)) bfs_remember(h, 2, bvh1->d_leaves, bvh1->n, bvh2->d_leaves, bvh2->n, start_level1, start_level2, r, fused);
    return d_contacts && (int64_t)total > capacity ? IBVH_ERR_CAPACITY : IBVH_OK;
}

int ibvh_traverse_bfs_rays(ibvh_handle_t* h, const ibvh_bvh_t* bvh, const void* d_points, const void* d_directions, int64_t nrays,
                           const ibvh_traverse_params_t* p, void* d_contacts, int64_t capacity, int64_t* num_contacts, int64_t* num_checks, void* stream) {
    if (!p || nrays < 0) return IBVH_ERR_ARGUMENT;
    int rc = bfs_common_checks(h, bvh, d_contacts, num_contacts, num_checks);
    if (rc != IBVH_OK) return rc;
    ibvh_tree_t tree;
    rc = check_bvh(bvh, &tree);
    if (rc != IBVH_OK) return rc;
    if (!(bvh->built_level <= p->start_level && p->start_level <= tree.levels)) return IBVH_ERR_ARGUMENT;   // raytrace/breadth_first/breadth_first.jl:12
    if (bfs_node_fbytes(bvh->types) != bvh->types.float_bytes) return IBVH_ERR_ARGUMENT;                    // isintersection.jl:1-5: one float type
    if (nrays == 0) return IBVH_OK;                                                                         // :25-29
    if (!d_points || !d_directions) return IBVH_ERR_ARGUMENT;
    if (nrays >= (int64_t(1) << 32)) { h->set_error("BFS ray traversal: ray ids are kept in 32 bits (nrays < 2^32)"); return IBVH_ERR_UNSUPPORTED; }
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    BfsRun r{h, st};
    const int64_t levels = tree.levels, sl = p->start_level;
    if (!bfs_resume(h, p->flags, 3, bvh->d_leaves, bvh->n, d_points, nrays, sl, sl, &r)) {
        const int64_t first = int64_t(1) << (sl - 1);
        const int64_t nreal = first - shr64(tree.virtual_leaves, levels - sl);
        const unsigned long long n0 = (unsigned long long)nreal * (unsigned long long)nrays;
        rc = r.reserve(0, n0);
        if (rc != IBVH_OK) return rc;
        { ProfScope _ps(h, st, "bfs_init_kernel");
        bfs_init_product_kernel<<<(unsigned)std::min<unsigned long long>((n0 + 255) / 256, 1u << 20), 256, 0, st>>>(r.list(0), (uint32_t)first, (unsigned long long)nreal, 1u, (unsigned long long)nrays);
        }
        IBVH_LAUNCH_CHECK(h, "bfs_init_product_kernel");
        r.cur = 0; r.count = n0; r.checks = (long long)n0;
        rc = bfs_dispatch_volume(bvh->types.node_kind, bvh->types.float_bytes, [&](auto ntag) -> int {
            using N = typename decltype(ntag)::type; using T = typename N::value_type;
            for (int64_t level = sl; level < levels && r.count; ++level) {
                int rc2 = r.begin_step(2ull * r.count);
                if (rc2 != IBVH_OK) return rc2;
                { ProfScope _ps(h, st, "bfs_rays_nodes_kernel");
                bfs_rays_nodes_kernel<N><<<r.grid(), kBfsThreads, 0, st>>>(r.list(r.cur), r.count, bfs_node_side(bvh, tree, level), (const T*)d_points, (const T*)d_directions,
                                                                            r.list(r.cur ^ 1), r.d_counter());
                }
                IBVH_LAUNCH_CHECK(h, "bfs_rays_nodes_kernel");
                if ((rc2 = r.end_step()) != IBVH_OK) return rc2;
            }
            return IBVH_OK;
        });
        if (rc != IBVH_OK) return rc;
    }
    if (num_checks) *num_checks = r.checks;
    if (r.count == 0) return IBVH_OK;
    const BfsSide ls = bfs_leaf_side(bvh, tree);
    uint32_t stride, io;
    bfs_leaf_layout(bvh->types, &stride, &io);
    IBVH_CUDA_TRY(h, cudaMemsetAsync(r.d_counter(), 0, 8, st));
    rc = bfs_dispatch_volume(bvh->types.leaf_kind, bvh->types.float_bytes, [&](auto vtag) -> int {
        using V = typename decltype(vtag)::type; using T = typename V::value_type;
        return bfs_dispatch_index(bvh->types.index_bytes, [&](auto itag) -> int {
            using I = typename decltype(itag)::type;
            { ProfScope _ps(h, st, "bfs_rays_leaves_kernel");
            bfs_rays_leaves_kernel<V, I><<<r.grid(), kBfsThreads, 0, st>>>(r.list(r.cur), r.count, ls, (const T*)d_points, (const T*)d_directions, io,
                                                                           (p->flags & IBVH_TRAVERSE_POSITIONS) ? 1 : 0, (long long)p->id_base,
                                                                           (IndexPair<I>*)d_contacts, d_contacts ? (unsigned long long)std::max<int64_t>(capacity, 0) : 0ull, r.d_counter());
            }
            IBVH_LAUNCH_CHECK(h, "bfs_rays_leaves_kernel");
            return IBVH_OK;
        });
    });
    if (rc != IBVH_OK) return rc;
    unsigned long long total;
    rc = r.read_counter(&total);
    if (rc != IBVH_OK) return rc;
    *num_contacts = (int64_t)total;
    if ((d_contacts && (int64_t)total > capacity) || (!d_contacts && total > 0)) bfs_remember(h, 3, bvh->d_leaves, bvh->n, d_points, nrays, sl, sl, r);
    return d_contacts && (int64_t)total > capacity ? IBVH_ERR_CAPACITY : IBVH_OK;
}
// Opt-in post-processing of a contact list on the device: ascending by (a, b), optionally unique.
int ibvh_sort_contacts(ibvh_handle_t* h, void* d_contacts, int64_t count, int32_t index_bytes, int unique, int64_t* out_count, void* stream) {
    if (!h || count < 0 || (index_bytes != 4 && index_bytes != 8) || (count > 0 && !d_contacts)) return IBVH_ERR_ARGUMENT;
    if (h->deferred.active) { h->set_error("a deferred traversal is outstanding on this handle: call ibvh_traverse_finish (or ibvh_traverse_cancel) first"); return IBVH_ERR_ARGUMENT; }
    if (out_count) *out_count = count;
    if (count <= 1) return IBVH_OK;
    if (count > (int64_t(1) << 31)) { h->set_error("ibvh_sort_contacts: more than 2^31 pairs"); return IBVH_ERR_UNSUPPORTED; }
    if (reinterpret_cast<uintptr_t>(d_contacts) % (2u * (unsigned)index_bytes) != 0) return IBVH_ERR_ARGUMENT;
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    using M = uint64_t;
    constexpr int P = radix_passes<M>();
    const int64_t n = count;
    const size_t lb_bytes = (sort_wide_lookback(n) || h->cfg.force_wide_lookback) ? 8 : 4;
    const int64_t tiles = (n + sort_tile<M>() - 1) / sort_tile<M>();
    const int64_t scan_blocks = (n + kScanTile - 1) / kScanTile;
    size_t need = 2 * ibvh_handle::padded((size_t)n * 8) + 2 * ibvh_handle::padded((size_t)n * 4) + ibvh_handle::padded((size_t)P * radix_bins<M>() * 4) +
                  ibvh_handle::padded((size_t)P * tiles * radix_bins<M>() * lb_bytes) + ibvh_handle::padded((size_t)scan_blocks * 8) + 4096;
    int rc = h->reserve(need);
    if (rc != IBVH_OK) return rc;
    h->reset();
    M* keysA = h->alloc<M>(n); M* keysB = h->alloc<M>(n);
    uint32_t* valsA = h->alloc<uint32_t>(n); uint32_t* valsB = h->alloc<uint32_t>(n);
    uint32_t* hist = h->alloc<uint32_t>((size_t)P * radix_bins<M>());
    unsigned char* lookback = h->alloc<unsigned char>((size_t)P * tiles * radix_bins<M>() * lb_bytes);
    long long* block_sums = h->alloc<long long>((size_t)scan_blocks);
    uint32_t* tickets = (uint32_t*)(h->d_small + kSmallTickets);
    if (!keysA || !keysB || !valsA || !valsB || !hist || !lookback || !block_sums) { h->set_error("workspace carve failed"); return IBVH_ERR_ALLOC; }
    IBVH_CUDA_TRY(h, cudaMemsetAsync(lookback, 0, (size_t)P * tiles * radix_bins<M>() * lb_bytes, st));
    IBVH_CUDA_TRY(h, cudaMemsetAsync(hist, 0, (size_t)P * radix_bins<M>() * 4, st));
    IBVH_CUDA_TRY(h, cudaMemsetAsync(tickets, 0, 16 * 4, st));
    uint32_t* d_bad = (uint32_t*)(h->d_small + kSmallBfs + 16);
    unsigned long long* d_total = (unsigned long long*)(h->d_small + kSmallBfs);
    IBVH_CUDA_TRY(h, cudaMemsetAsync(d_total, 0, 24, st));
    const int grid = grid_for(n, 256, 4, h->sm_count * 8);
    { ProfScope _ps(h, st, "contact_keys_kernel");
    if (index_bytes == 4) contact_keys_kernel<int32_t><<<grid, 256, 0, st>>>((const IndexPair<int32_t>*)d_contacts, n, keysA, hist, d_bad);
    else contact_keys_kernel<int64_t><<<grid, 256, 0, st>>>((const IndexPair<int64_t>*)d_contacts, n, keysA, hist, d_bad);
    }
    IBVH_LAUNCH_CHECK(h, "contact_keys_kernel");
    unsigned long long* hp = (unsigned long long*)h->h_pinned;
    IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, d_bad, 4, cudaMemcpyDeviceToHost, st));
    IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
    if (((const uint32_t*)hp)[0] != 0) {
        h->set_error("ibvh_sort_contacts: an index outside [0, 2^31) x [0, 2^32) (the keys are a << 32 | b); the list is untouched");
        return IBVH_ERR_UNSUPPORTED;
    }
    M* sorted; uint32_t* perm;
    rc = sort_pairs<M>(h, keysA, keysB, valsA, valsB, n, hist, lookback, tickets, st, &sorted, &perm);
    if (rc != IBVH_OK) return rc;
    int32_t* slots = nullptr;
    if (unique) {
        slots = (int32_t*)(perm == valsA ? valsB : valsA);          // the permutation is not needed: its other buffer holds the flags
        { ProfScope _ps(h, st, "contact_flags_kernel");
        contact_flags_kernel<<<grid, 256, 0, st>>>(sorted, n, slots);
        }
        IBVH_LAUNCH_CHECK(h, "contact_flags_kernel");
        rc = scan_counts<int32_t>(h, slots, n, block_sums, d_total, st);
        if (rc != IBVH_OK) return rc;
    }
    { ProfScope _ps(h, st, "contact_unpack_kernel");
    if (index_bytes == 4) contact_unpack_kernel<int32_t><<<grid, 256, 0, st>>>(sorted, n, slots, (IndexPair<int32_t>*)d_contacts);
    else contact_unpack_kernel<int64_t><<<grid, 256, 0, st>>>(sorted, n, slots, (IndexPair<int64_t>*)d_contacts);
    }
    IBVH_LAUNCH_CHECK(h, "contact_unpack_kernel");
    if (unique) {
        IBVH_CUDA_TRY(h, cudaMemcpyAsync(hp, d_total, 8, cudaMemcpyDeviceToHost, st));
        IBVH_CUDA_TRY(h, cudaStreamSynchronize(st));
        if (out_count) *out_count = (int64_t)hp[0];
    }
    return IBVH_OK;
}
#endif  // IBVH_PART_BFS

}  // extern "C"
