// ibvh_oracle_capi.cpp — extern "C" surface of the CPU oracle (TEST INFRASTRUCTURE ONLY; see the
// header of ibvh_oracle.hpp). Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs. Never linked into libibvh_b200.so.
//
// Type codes (same as include/ibvh.h): vol kind 0 = BSphere, 1 = BBox; float bytes 4 / 8;
// index bytes 4 / 8 (Int32 / Int64); morton bytes 2 / 4 / 8 (UInt16 / UInt32 / UInt64).
#include "ibvh_oracle.hpp"

using namespace orc;

namespace {

template <class X> struct Tag { using type = X; };

struct LeafDesc { int kind, fbytes, ibytes, mbytes; };
struct NodeDesc { int kind, fbytes; };

template <class V, class F> int dispatch_im(const LeafDesc& d, F&& f) {
    if (d.ibytes == 4) {
        if (d.mbytes == 2) return f(Tag<Leaf<V, int32_t, uint16_t>>{});
        if (d.mbytes == 4) return f(Tag<Leaf<V, int32_t, uint32_t>>{});
        if (d.mbytes == 8) return f(Tag<Leaf<V, int32_t, uint64_t>>{});
    } else if (d.ibytes == 8) {
        if (d.mbytes == 2) return f(Tag<Leaf<V, int64_t, uint16_t>>{});
        if (d.mbytes == 4) return f(Tag<Leaf<V, int64_t, uint32_t>>{});
        if (d.mbytes == 8) return f(Tag<Leaf<V, int64_t, uint64_t>>{});
    }
    return -1;
}
template <class F> int dispatch_leaf(const LeafDesc& d, F&& f) {
    if (d.kind == 0 && d.fbytes == 4) return dispatch_im<BSphere<float>>(d, f);
    if (d.kind == 0 && d.fbytes == 8) return dispatch_im<BSphere<double>>(d, f);
    if (d.kind == 1 && d.fbytes == 4) return dispatch_im<BBox<float>>(d, f);
    if (d.kind == 1 && d.fbytes == 8) return dispatch_im<BBox<double>>(d, f);
    return -1;
}
// Allowed (leaf volume, node) pairs: node float == leaf float, or Float64 leaves -> Float32 nodes;
// BBox leaves cannot feed BSphere nodes (the reference has no BSphere(::BBox) constructor).
template <class L, class F> int dispatch_node(const NodeDesc& nd, F&& f) {
    using V = typename L::vol_t;
    using T = typename V::value_type;
    constexpr bool leaf_is_sphere = std::is_same<V, BSphere<T>>::value;
    if (nd.kind == 1) {
        if (nd.fbytes == (int)sizeof(T)) return f(Tag<BBox<T>>{});
        if constexpr (std::is_same<T, double>::value) { if (nd.fbytes == 4) return f(Tag<BBox<float>>{}); }
    } else if (nd.kind == 0) {
        if constexpr (leaf_is_sphere) {
            if (nd.fbytes == (int)sizeof(T)) return f(Tag<BSphere<T>>{});
            if constexpr (std::is_same<T, double>::value) { if (nd.fbytes == 4) return f(Tag<BSphere<float>>{}); }
        }
    }
    return -1;
}

void skips_vec(const Tree& t, std::vector<int64_t>& s) { s.resize((size_t)t.levels); compute_skips(t, s.data()); }

}  // namespace

extern "C" {

// ---- implicit tree (implicit_tree.jl) -------------------------------------------------------
int orc_tree_shape(int64_t n, int64_t* tree5, int64_t* skips /* >= 64 slots or NULL */) {
    if (n < 1) return -2;                         // DomainError, implicit_tree.jl:78-80
    Tree t = make_tree(n);
    tree5[0] = t.levels; tree5[1] = t.real_leaves; tree5[2] = t.real_nodes;
    tree5[3] = t.virtual_leaves; tree5[4] = t.virtual_nodes;
    if (skips) compute_skips(t, skips);
    return 0;
}
int64_t orc_memory_index(int64_t n, int64_t implicit_index) { return memory_index(make_tree(n), implicit_index); }
int orc_level_indices(int64_t n, int64_t level, int64_t* start, int64_t* stop) { level_indices(make_tree(n), level, start, stop); return 0; }
int orc_isvirtual(int64_t n, int64_t implicit_index) { return isvirtual(make_tree(n), implicit_index) ? 1 : 0; }
int64_t orc_build_level_float(int64_t levels, double f) { return compute_build_level_float(levels, f); }

// ---- scalar pieces for the known-answer tests ------------------------------------------------
uint16_t orc_morton_split3_u16(uint16_t v) { return morton_split3(v); }
uint32_t orc_morton_split3_u32(uint32_t v) { return morton_split3(v); }
uint64_t orc_morton_split3_u64(uint64_t v) { return morton_split3(v); }

int orc_ray_box_f64(const double* box6, const double* p, const double* d) {
    BBox<double> b; for (int k = 0; k < 3; ++k) { b.lo[k] = box6[k]; b.up[k] = box6[3 + k]; }
    return isintersection(b, p, d) ? 1 : 0;
}
int orc_ray_sphere_f64(const double* s4, const double* p, const double* d) {
    BSphere<double> s; for (int k = 0; k < 3; ++k) s.x[k] = s4[k]; s.r = s4[3];
    return isintersection(s, p, d) ? 1 : 0;
}
int orc_ray_box_f32(const float* box6, const float* p, const float* d) {
    BBox<float> b; for (int k = 0; k < 3; ++k) { b.lo[k] = box6[k]; b.up[k] = box6[3 + k]; }
    return isintersection(b, p, d) ? 1 : 0;
}
int orc_ray_sphere_f32(const float* s4, const float* p, const float* d) {
    BSphere<float> s; for (int k = 0; k < 3; ++k) s.x[k] = s4[k]; s.r = s4[3];
    return isintersection(s, p, d) ? 1 : 0;
}
void orc_merge_sphere_f64(const double* a4, const double* b4, double* out4) {
    BSphere<double> a, b; for (int k = 0; k < 3; ++k) { a.x[k] = a4[k]; b.x[k] = b4[k]; } a.r = a4[3]; b.r = b4[3];
    BSphere<double> c = merge_sphere<double>(a, b);
    for (int k = 0; k < 3; ++k) out4[k] = c.x[k]; out4[3] = c.r;
}
void orc_merge_sphere_f32(const float* a4, const float* b4, float* out4) {
    BSphere<float> a, b; for (int k = 0; k < 3; ++k) { a.x[k] = a4[k]; b.x[k] = b4[k]; } a.r = a4[3]; b.r = b4[3];
    BSphere<float> c = merge_sphere<float>(a, b);
    for (int k = 0; k < 3; ++k) out4[k] = c.x[k]; out4[3] = c.r;
}
void orc_merge_box_f64(const double* a6, const double* b6, double* out6) {
    BBox<double> a, b; for (int k = 0; k < 3; ++k) { a.lo[k] = a6[k]; a.up[k] = a6[3 + k]; b.lo[k] = b6[k]; b.up[k] = b6[3 + k]; }
    BBox<double> c = merge_box<double>(a, b);
    for (int k = 0; k < 3; ++k) { out6[k] = c.lo[k]; out6[3 + k] = c.up[k]; }
}
void orc_merge_spheres_to_box_f64(const double* a4, const double* b4, double* out6) {
    BSphere<double> a, b; for (int k = 0; k < 3; ++k) { a.x[k] = a4[k]; b.x[k] = b4[k]; } a.r = a4[3]; b.r = b4[3];
    BBox<double> c = merge_box<double>(a, b);
    for (int k = 0; k < 3; ++k) { out6[k] = c.lo[k]; out6[3 + k] = c.up[k]; }
}
void orc_merge_spheres_to_box_f32(const float* a4, const float* b4, float* out6) {
    BSphere<float> a, b; for (int k = 0; k < 3; ++k) { a.x[k] = a4[k]; b.x[k] = b4[k]; } a.r = a4[3]; b.r = b4[3];
    BBox<float> c = merge_box<float>(a, b);
    for (int k = 0; k < 3; ++k) { out6[k] = c.lo[k]; out6[3 + k] = c.up[k]; }
}
int orc_contact_sphere_f32(const float* a4, const float* b4) {
    BSphere<float> a, b; for (int k = 0; k < 3; ++k) { a.x[k] = a4[k]; b.x[k] = b4[k]; } a.r = a4[3]; b.r = b4[3];
    return iscontact(a, b) ? 1 : 0;
}
int orc_contact_sphere_f64(const double* a4, const double* b4) {
    BSphere<double> a, b; for (int k = 0; k < 3; ++k) { a.x[k] = a4[k]; b.x[k] = b4[k]; } a.r = a4[3]; b.r = b4[3];
    return iscontact(a, b) ? 1 : 0;
}

// ---- leaf volumes from triangles: tri = T[n][3][3] (three vertices of three coordinates) -------
int orc_volumes_from_triangles(const void* tri, int64_t n, int kind, int fbytes, void* out) {
    auto run = [&](auto ttag) {
        using T = typename decltype(ttag)::type;
        const T* t = (const T*)tri;
        for (int64_t i = 0; i < n; ++i) {
            const T* a = t + 9 * i; const T* b = a + 3; const T* c = a + 6;
            if (kind == 0) ((BSphere<T>*)out)[i] = sphere_from_triangle<T>(a, b, c);
            else ((BBox<T>*)out)[i] = box_from_triangle<T>(a, b, c);
        }
    };
    if (kind != 0 && kind != 1) return -1;
    if (fbytes == 4) run(Tag<float>{}); else if (fbytes == 8) run(Tag<double>{}); else return -1;
    return 0;
}

// ---- size of a wrapped leaf ------------------------------------------------------------------
int64_t orc_leaf_bytes(int kind, int fbytes, int ibytes, int mbytes) {
    int64_t out = -1;
    dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) { using L = typename decltype(tag)::type; out = sizeof(L); return 0; });
    return out;
}

// ---- wrap (build.jl:328-352): BoundingVolume(bv[i], I(i), M(0)), i 1-based --------------------
int orc_wrap(const void* volumes, int64_t n, int kind, int fbytes, int ibytes, int mbytes, void* leaves_out) {
    return dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using V = typename L::vol_t;
        const V* v = (const V*)volumes; L* o = (L*)leaves_out;
        for (int64_t i = 0; i < n; ++i) {
            std::memset((void*)&o[i], 0, sizeof(L));       // deterministic padding bytes
            o[i].volume = v[i]; o[i].index = (typename L::idx_t)(i + 1); o[i].morton = 0;
        }
        return 0;
    });
}

// ---- Morton (morton/utils.jl, morton/default.jl) ---------------------------------------------
// mins/maxs: 3 values of the leaf float type; written when compute_extrema != 0, read otherwise.
int orc_morton_encode(void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes,
                      int compute_extrema_flag, void* mins, void* maxs, int num_threads, int64_t min_elems) {
    return dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using T = typename L::vol_t::value_type;
        morton_encode((L*)leaves, n, compute_extrema_flag != 0, (T*)mins, (T*)maxs, num_threads, min_elems);
        return 0;
    });
}
int orc_sort_leaves(void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes, int num_threads, int64_t min_elems) {
    return dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type;
        sort_leaves((L*)leaves, n, num_threads, min_elems);
        return 0;
    });
}
int orc_aggregate(const void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes,
                  void* nodes, int node_kind, int node_fbytes, int64_t built_level, int num_threads, int64_t min_elems) {
    return dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            Tree t = make_tree(n);
            if (built_level < 1 || built_level > t.levels) return -3;
            if (t.real_nodes >= 2) aggregate<L, N>((N*)nodes, (const L*)leaves, t, built_level, num_threads, min_elems);
            return 0;
        });
    });
}
// build.jl:198-271 (encode + sort + aggregate; leaves modified in place)
int orc_build(void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes,
              void* nodes, int node_kind, int node_fbytes, int64_t built_level,
              int compute_extrema_flag, void* mins, void* maxs,
              int num_threads, int64_t min_mortons, int64_t min_sorts, int64_t min_boundings) {
    return dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using T = typename L::vol_t::value_type;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            if (n < 1) return -2;
            Tree t = make_tree(n);
            if (built_level < 1 || built_level > t.levels) return -3;
            build<L, N>((L*)leaves, n, (N*)nodes, built_level, compute_extrema_flag != 0, (T*)mins, (T*)maxs,
                        num_threads, min_mortons, min_sorts, min_boundings);
            return 0;
        });
    });
}

// ---- LVT traversals ---------------------------------------------------------------------------
// contacts: IndexPair{I} array of `capacity` pairs or NULL (count only). counts: I[n queries] or
// NULL — receives the inclusive scan of per-query counts (GPU-backend form of cache2).
// Returns the total number of contacts, or a negative error.
int64_t orc_traverse_single(const void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes,
                            const void* nodes, int node_kind, int node_fbytes, int64_t built_level,
                            int64_t start_level, void* contacts, int64_t capacity, void* counts,
                            int num_threads, int64_t min_elems) {
    int64_t total = -1;
    int rc = dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using I = typename L::idx_t;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            Tree t = make_tree(n);
            if (!(built_level <= start_level && start_level <= t.levels && t.levels <= 32)) return -3;
            if (t.real_nodes <= 1) { total = 0; return 0; }                 // traverse_single.jl:17-21
            std::vector<int64_t> sk; skips_vec(t, sk);
            BVHView<L, N> bvh{t, sk.data(), (const N*)nodes, (const L*)leaves, built_level};
            total = two_pass<I>(n, num_threads, min_elems, (IndexPair<I>*)contacts, capacity, (I*)counts,
                                [&](int64_t q, Emit<I>& em) { traverse_lvt_single<L, N, I>(bvh.leaves[q], q + 1, bvh, start_level, em); });
            return 0;
        });
    });
    return rc < 0 ? rc : total;
}

// traverse(bvh1, bvh2): the caller passes them in user order; the swap to "larger tree queries"
// and the flip flag follow traverse_pair.jl:16-36.
int64_t orc_traverse_pair(const void* leaves1, int64_t n1, const void* nodes1, int64_t built_level1, int64_t start_level1,
                          const void* leaves2, int64_t n2, const void* nodes2, int64_t built_level2, int64_t start_level2,
                          int kind, int fbytes, int ibytes, int mbytes, int node_kind, int node_fbytes,
                          void* contacts, int64_t capacity, void* counts, int num_threads, int64_t min_elems) {
    int64_t total = -1;
    int rc = dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using I = typename L::idx_t;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            Tree t1 = make_tree(n1), t2 = make_tree(n2);
            if (!(built_level1 <= start_level1 && start_level1 <= t1.levels && t1.levels <= 32)) return -3;
            if (!(built_level2 <= start_level2 && start_level2 <= t2.levels && t2.levels <= 32)) return -3;
            bool flip = !(n1 >= n2);
            const L* ql = (const L*)(flip ? leaves2 : leaves1);
            int64_t nq = flip ? n2 : n1;
            const Tree& tt = flip ? t1 : t2;
            std::vector<int64_t> sk; skips_vec(tt, sk);
            BVHView<L, N> target{tt, sk.data(), (const N*)(flip ? nodes1 : nodes2), (const L*)(flip ? leaves1 : leaves2),
                                 flip ? built_level1 : built_level2};
            int64_t sl = flip ? start_level1 : start_level2;
            total = two_pass<I>(nq, num_threads, min_elems, (IndexPair<I>*)contacts, capacity, (I*)counts,
                                [&](int64_t q, Emit<I>& em) { traverse_lvt_pair<L, L, N, I>(ql[q], target, sl, flip, em); });
            (void)nodes1;
            return 0;
        });
    });
    return rc < 0 ? rc : total;
}

// points/directions: column-major 3 x R of `ray_fbytes` floats; converted to the leaf float type per
// ray (leaf_vs_tree.jl:116-125).
int64_t orc_traverse_rays(const void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes,
                          const void* nodes, int node_kind, int node_fbytes, int64_t built_level,
                          int64_t start_level, const void* points, const void* directions, int ray_fbytes, int64_t nrays,
                          void* contacts, int64_t capacity, void* counts, int num_threads, int64_t min_elems) {
    int64_t total = -1;
    int rc = dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using I = typename L::idx_t; using T = typename L::vol_t::value_type;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            if constexpr (!std::is_same<typename N::value_type, T>::value) return -1;   // isintersection needs one T
            else {
                Tree t = make_tree(n);
                if (!(built_level <= start_level && start_level <= t.levels && t.levels <= 32)) return -3;
                if (nrays == 0) { total = 0; return 0; }
                std::vector<int64_t> sk; skips_vec(t, sk);
                BVHView<L, N> bvh{t, sk.data(), (const N*)nodes, (const L*)leaves, built_level};
                auto getp = [&](const void* base, int64_t i, T out[3]) {
                    if (ray_fbytes == 4) { const float* p = (const float*)base + 3 * i; out[0] = (T)p[0]; out[1] = (T)p[1]; out[2] = (T)p[2]; }
                    else { const double* p = (const double*)base + 3 * i; out[0] = (T)p[0]; out[1] = (T)p[1]; out[2] = (T)p[2]; }
                };
                total = two_pass<I>(nrays, num_threads, min_elems, (IndexPair<I>*)contacts, capacity, (I*)counts,
                                    [&](int64_t q, Emit<I>& em) {
                                        T p[3], d[3]; getp(points, q, p); getp(directions, q, d);
                                        traverse_ray_lvt<L, N, I, T>(p, d, q + 1, bvh, start_level, em);
                                    });
                return 0;
            }
        });
    });
    return rc < 0 ? rc : total;
}

// ---- BFS traversals (src/traverse/breadth_first, src/raytrace/breadth_first) ------------------------
// Each call runs the whole traversal and keeps the list in a process-wide stash; orc_bfs_fetch copies it
// out (the list length is the call's return value). *num_checks receives BVHTraversal.num_checks.
static std::vector<char> g_bfs_stash;
#define bfs_stash(v)                                                                        \
    do {                                                                                    \
        g_bfs_stash.resize((v).size() * sizeof((v)[0]));                                    \
        if (!(v).empty()) std::memcpy(g_bfs_stash.data(), (v).data(), g_bfs_stash.size());  \
    } while (0)
int64_t orc_bfs_fetch(void* out, int64_t bytes) {
    if ((int64_t)g_bfs_stash.size() != bytes) return -1;
    if (bytes) std::memcpy(out, g_bfs_stash.data(), (size_t)bytes);
    return 0;
}
int64_t orc_traverse_bfs_single(const void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes,
                                const void* nodes, int node_kind, int node_fbytes, int64_t built_level,
                                int64_t start_level, int positions, int64_t* num_checks) {
    int64_t total = -1;
    int rc = dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using I = typename L::idx_t;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            Tree t = make_tree(n);
            if (!(built_level <= start_level && start_level <= t.levels)) return -3;      // traverse_single.jl:10
            std::vector<int64_t> sk; skips_vec(t, sk);
            BVHView<L, N> bvh{t, sk.data(), (const N*)nodes, (const L*)leaves, built_level};
            std::vector<IndexPair<I>> out;
            total = traverse_bfs_single<L, N, I>(bvh, start_level, positions != 0, out, num_checks);
            bfs_stash(out);
            return 0;
        });
    });
    return rc < 0 ? rc : total;
}
int64_t orc_traverse_bfs_pair(const void* leaves1, int64_t n1, const void* nodes1, int64_t built_level1, int64_t start_level1,
                              const void* leaves2, int64_t n2, const void* nodes2, int64_t built_level2, int64_t start_level2,
                              int kind, int fbytes, int ibytes, int mbytes, int node_kind, int node_fbytes,
                              int positions, int64_t* num_checks) {
    int64_t total = -1;
    int rc = dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using I = typename L::idx_t;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            Tree t1 = make_tree(n1), t2 = make_tree(n2);
            if (!(built_level1 <= start_level1 && start_level1 <= t1.levels)) return -3;  // traverse_pair.jl:10-11
            if (!(built_level2 <= start_level2 && start_level2 <= t2.levels)) return -3;
            std::vector<int64_t> s1, s2; skips_vec(t1, s1); skips_vec(t2, s2);
            BVHView<L, N> b1{t1, s1.data(), (const N*)nodes1, (const L*)leaves1, built_level1};
            BVHView<L, N> b2{t2, s2.data(), (const N*)nodes2, (const L*)leaves2, built_level2};
            std::vector<IndexPair<I>> out;
            total = traverse_bfs_pair<L, N, I>(b1, b2, start_level1, start_level2, positions != 0, out, num_checks);
            bfs_stash(out);
            return 0;
        });
    });
    return rc < 0 ? rc : total;
}
int64_t orc_traverse_bfs_rays(const void* leaves, int64_t n, int kind, int fbytes, int ibytes, int mbytes,
                              const void* nodes, int node_kind, int node_fbytes, int64_t built_level,
                              int64_t start_level, const void* points, const void* directions, int64_t nrays,
                              int positions, int64_t* num_checks) {
    int64_t total = -1;
    int rc = dispatch_leaf({kind, fbytes, ibytes, mbytes}, [&](auto tag) {
        using L = typename decltype(tag)::type; using I = typename L::idx_t; using T = typename L::vol_t::value_type;
        return dispatch_node<L>({node_kind, node_fbytes}, [&](auto ntag) {
            using N = typename decltype(ntag)::type;
            if constexpr (!std::is_same<typename N::value_type, T>::value) return -1;   // isintersection needs one T
            else {
                Tree t = make_tree(n);
                if (!(built_level <= start_level && start_level <= t.levels)) return -3;
                std::vector<int64_t> sk; skips_vec(t, sk);
                BVHView<L, N> bvh{t, sk.data(), (const N*)nodes, (const L*)leaves, built_level};
                std::vector<IndexPair<I>> out;
                total = traverse_bfs_rays<L, N, I, T>(bvh, (const T*)points, (const T*)directions, nrays, start_level,
                                                      positions != 0, out, num_checks);
                bfs_stash(out);
                return 0;
            }
        });
    });
    return rc < 0 ? rc : total;
}

// Brute force O(n^2) single-set contacts in *input order* (test/runtests.jl:851-859): pairs (i, j),
// i < j, 1-based, iscontact on the raw volumes. Returns count; writes up to capacity pairs (int64).
int64_t orc_brute_single(const void* volumes, int64_t n, int kind, int fbytes, int64_t* pairs, int64_t capacity) {
    int64_t c = 0;
    auto run = [&](auto vtag) {
        using V = typename decltype(vtag)::type; const V* v = (const V*)volumes;
        for (int64_t i = 0; i < n; ++i) for (int64_t j = i + 1; j < n; ++j)
            if (iscontact(v[i], v[j])) { if (c < capacity) { pairs[2 * c] = i + 1; pairs[2 * c + 1] = j + 1; } ++c; }
    };
    if (kind == 0 && fbytes == 4) run(Tag<BSphere<float>>{}); else if (kind == 0 && fbytes == 8) run(Tag<BSphere<double>>{});
    else if (kind == 1 && fbytes == 4) run(Tag<BBox<float>>{}); else if (kind == 1 && fbytes == 8) run(Tag<BBox<double>>{});
    else return -1;
    return c;
}
int64_t orc_brute_pair(const void* v1, int64_t n1, const void* v2, int64_t n2, int kind, int fbytes, int64_t* pairs, int64_t capacity) {
    int64_t c = 0;
    auto run = [&](auto vtag) {
        using V = typename decltype(vtag)::type; const V* a = (const V*)v1; const V* b = (const V*)v2;
        for (int64_t i = 0; i < n1; ++i) for (int64_t j = 0; j < n2; ++j)
            if (iscontact(a[i], b[j])) { if (c < capacity) { pairs[2 * c] = i + 1; pairs[2 * c + 1] = j + 1; } ++c; }
    };
    if (kind == 0 && fbytes == 4) run(Tag<BSphere<float>>{}); else if (kind == 0 && fbytes == 8) run(Tag<BSphere<double>>{});
    else if (kind == 1 && fbytes == 4) run(Tag<BBox<float>>{}); else if (kind == 1 && fbytes == 8) run(Tag<BBox<double>>{});
    else return -1;
    return c;
}
int64_t orc_brute_rays(const void* volumes, int64_t n, int kind, int fbytes, const void* points, const void* dirs, int64_t nrays,
                       int64_t* pairs, int64_t capacity) {
    int64_t c = 0;
    auto run = [&](auto vtag) {
        using V = typename decltype(vtag)::type; using T = typename V::value_type;
        const V* v = (const V*)volumes; const T* p = (const T*)points; const T* d = (const T*)dirs;
        for (int64_t r = 0; r < nrays; ++r) for (int64_t i = 0; i < n; ++i)
            if (isintersection(v[i], p + 3 * r, d + 3 * r)) { if (c < capacity) { pairs[2 * c] = i + 1; pairs[2 * c + 1] = r + 1; } ++c; }
    };
    if (kind == 0 && fbytes == 4) run(Tag<BSphere<float>>{}); else if (kind == 0 && fbytes == 8) run(Tag<BSphere<double>>{});
    else if (kind == 1 && fbytes == 4) run(Tag<BBox<float>>{}); else if (kind == 1 && fbytes == 8) run(Tag<BBox<double>>{});
    else return -1;
    return c;
}

}  // extern "C"
