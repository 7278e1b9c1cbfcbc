// ibvh_oracle.hpp — CPU restatement of the ImplicitBVH.jl hot path (TEST INFRASTRUCTURE ONLY).
//
// This file is the parity ORACLE for the B200 kernels. It is a from-scratch C++17 restatement of the
// reference's *algorithm* (Julia, /root/reference/src), function by function, with every function
// citing the reference file:line it follows. It is NOT part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Parity status: PINNED against the reference's own known-answer tests and doctests
// (tests/golden/reference_known_answers.json, extracted by hand from test/runtests.jl, README.md
// and the docstrings; the reference itself is Julia and cannot be executed in this image).
// One thing is unpinned by the reference itself: the order of equal Morton keys after AK.sort!
// (third-party AcceleratedKernels 0.4, not mounted). We adopt a STABLE ascending sort.
//
// Arithmetic rules reproduced here: no FMA contraction (compile with
// -ffp-contract=off), IEEE div/sqrt, denormals kept, left-to-right association exactly as written
// in the Julia source, Julia's `a < b ? a : b` NaN behaviour, 1-based indices on the API surface.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <thread>
#include <type_traits>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Geometry types — isbits layouts of the reference (natural alignment == Julia struct layout)
//   BSphere{T}        src/bounding_volumes/bsphere.jl:26-29
//   BBox{T}           src/bounding_volumes/bbox.jl:35-38
//   BoundingVolume    src/bounding_volumes/bounding_volumes.jl:55-59  (struct Leaf below)
//   IndexPair{I}      src/traverse/traverse.jl:6
// ---------------------------------------------------------------------------------------------
template <class T> struct BSphere { T x[3]; T r; using value_type = T; };
template <class T> struct BBox    { T lo[3]; T up[3]; using value_type = T; };
template <class I> struct IndexPair { I a, b; };

// src/utils.jl:177-181 — note the NaN behaviour: comparison false => second argument.
template <class T> inline T minimum2(T a, T b) { return a < b ? a : b; }
template <class T> inline T maximum2(T a, T b) { return a > b ? a : b; }

// src/utils.jl:168-172 (left-to-right association, no contraction)
template <class T> inline T dist3sq(const T* x, const T* y) {
    T d0 = (x[0] - y[0]) * (x[0] - y[0]);
    T d1 = (x[1] - y[1]) * (x[1] - y[1]);
    T d2 = (x[2] - y[2]) * (x[2] - y[2]);
    return (d0 + d1) + d2;
}
template <class T> inline T dist3(const T* x, const T* y) { return std::sqrt(dist3sq(x, y)); }
// src/utils.jl:163-165
template <class T> inline T dot3(const T* x, const T* y) { return (x[0] * y[0] + x[1] * y[1]) + x[2] * y[2]; }

// center(): bsphere.jl:142, bbox.jl:100-102
template <class T> inline void center(const BSphere<T>& b, T c[3]) { c[0] = b.x[0]; c[1] = b.x[1]; c[2] = b.x[2]; }
template <class T> inline void center(const BBox<T>& b, T c[3]) {
    c[0] = T(0.5) * (b.lo[0] + b.up[0]);
    c[1] = T(0.5) * (b.lo[1] + b.up[1]);
    c[2] = T(0.5) * (b.lo[2] + b.up[2]);
}

// ---------------------------------------------------------------------------------------------
// Merges / conversions — src/bounding_volumes/merge.jl. `TN` is the node float type; arithmetic
// happens in the *leaf* float type and the result tuple is converted to TN, as Julia's
// BBox{T}(lower, upper) convert does.
// ---------------------------------------------------------------------------------------------
// merge.jl:47-51  BBox{T}(a::BSphere)
template <class TN, class TL> inline BBox<TN> to_box(const BSphere<TL>& a) {
    BBox<TN> o;
    for (int k = 0; k < 3; ++k) { o.lo[k] = TN(a.x[k] - a.r); o.up[k] = TN(a.x[k] + a.r); }
    return o;
}
// bbox.jl:54  BBox{T}(x::BBox)
template <class TN, class TL> inline BBox<TN> to_box(const BBox<TL>& a) {
    BBox<TN> o;
    for (int k = 0; k < 3; ++k) { o.lo[k] = TN(a.lo[k]); o.up[k] = TN(a.up[k]); }
    return o;
}
// bsphere.jl:38  BSphere{T}(x::BSphere)
template <class TN, class TL> inline BSphere<TN> to_sphere(const BSphere<TL>& a) {
    BSphere<TN> o;
    for (int k = 0; k < 3; ++k) o.x[k] = TN(a.x[k]);
    o.r = TN(a.r);
    return o;
}

// merge.jl:30-43  BBox{T}(a::BBox, b::BBox)
template <class TN, class TL> inline BBox<TN> merge_box(const BBox<TL>& a, const BBox<TL>& b) {
    BBox<TN> o;
    for (int k = 0; k < 3; ++k) {
        o.lo[k] = TN(minimum2(a.lo[k], b.lo[k]));
        o.up[k] = TN(maximum2(a.up[k], b.up[k]));
    }
    return o;
}
// merge.jl:58-81  BBox{T}(a::BSphere, b::BSphere)
template <class TN, class TL> inline BBox<TN> merge_box(const BSphere<TL>& a, const BSphere<TL>& b) {
    TL length = dist3(a.x, b.x);
    if (length + a.r <= b.r) return to_box<TN>(b);      // a enclosed in b
    if (length + b.r <= a.r) return to_box<TN>(a);      // b enclosed in a
    BBox<TN> o;
    for (int k = 0; k < 3; ++k) {
        o.lo[k] = TN(minimum2(a.x[k] - a.r, b.x[k] - b.r));
        o.up[k] = TN(maximum2(a.x[k] + a.r, b.x[k] + b.r));
    }
    return o;
}
// merge.jl:2-26  BSphere{T}(a::BSphere, b::BSphere)
template <class TN, class TL> inline BSphere<TN> merge_sphere(const BSphere<TL>& a, const BSphere<TL>& b) {
    TL length = dist3(a.x, b.x);
    if (length + a.r <= b.r) return to_sphere<TN>(b);
    if (length + b.r <= a.r) return to_sphere<TN>(a);
    // T(0.5) etc. are literals of the *target* type T in the reference; with TN != TL Julia
    // promotes, so the arithmetic runs in the wider of the two. We only instantiate TN == TL or
    // (TL=double, TN=float) where promotion gives double == TL.
    using W = typename std::conditional<(sizeof(TL) > sizeof(TN)), TL, TN>::type;
    W frac = W(0.5) * ((W(b.r) - W(a.r)) / W(length) + W(1));
    BSphere<TN> o;
    for (int k = 0; k < 3; ++k) o.x[k] = TN(W(a.x[k]) + frac * (W(b.x[k]) - W(a.x[k])));
    o.r = TN(W(0.5) * ((W(length) + W(a.r)) + W(b.r)));
    return o;
}

// Node-type dispatch used by build/traverse: NodeType(leaf.volume), NodeType(l, r), node + node.
template <class N> struct NodeOps;
template <class TN> struct NodeOps<BBox<TN>> {
    template <class V> static BBox<TN> convert(const V& v) { return to_box<TN>(v); }
    template <class V> static BBox<TN> merge(const V& a, const V& b) { return merge_box<TN>(a, b); }
};
template <class TN> struct NodeOps<BSphere<TN>> {
    template <class TL> static BSphere<TN> convert(const BSphere<TL>& v) { return to_sphere<TN>(v); }
    template <class TL> static BSphere<TN> merge(const BSphere<TL>& a, const BSphere<TL>& b) { return merge_sphere<TN>(a, b); }
};

// ---------------------------------------------------------------------------------------------
// Leaf volumes from triangles — src/bounding_volumes/bsphere.jl:43-112, bbox.jl:59-70 (the step right
// before the hot path). minimum3 / maximum3: src/utils.jl:178-181.
// ---------------------------------------------------------------------------------------------
template <class T> inline T minimum3(T a, T b, T c) { return a < b ? minimum2(a, c) : minimum2(b, c); }
template <class T> inline T maximum3(T a, T b, T c) { return a > b ? maximum2(a, c) : maximum2(b, c); }
template <class T> inline BSphere<T> sphere_from_triangle(const T* a, const T* b, const T* c) {
    T abab = ((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1])) + (b[2] - a[2]) * (b[2] - a[2]);
    T abac = ((b[0] - a[0]) * (c[0] - a[0]) + (b[1] - a[1]) * (c[1] - a[1])) + (b[2] - a[2]) * (c[2] - a[2]);
    T acac = ((c[0] - a[0]) * (c[0] - a[0]) + (c[1] - a[1]) * (c[1] - a[1])) + (c[2] - a[2]) * (c[2] - a[2]);
    T d = T(2) * (abab * acac - abac * abac);
    BSphere<T> o;
    if (std::fabs(d) <= std::numeric_limits<T>::epsilon()) {
        T lo[3], up[3];
        for (int k = 0; k < 3; ++k) { lo[k] = minimum3(a[k], b[k], c[k]); up[k] = maximum3(a[k], b[k], c[k]); }
        for (int k = 0; k < 3; ++k) o.x[k] = T(0.5) * (lo[k] + up[k]);
        o.r = dist3(o.x, up);
        return o;
    }
    T s = (abab * acac - acac * abac) / d;
    T t = (acac * abab - abab * abac) / d;
    if (s <= T(0)) { for (int k = 0; k < 3; ++k) o.x[k] = T(0.5) * (a[k] + c[k]); o.r = dist3(o.x, a); }
    else if (t <= T(0)) { for (int k = 0; k < 3; ++k) o.x[k] = T(0.5) * (a[k] + b[k]); o.r = dist3(o.x, a); }
    else if (s + t >= T(1)) { for (int k = 0; k < 3; ++k) o.x[k] = T(0.5) * (b[k] + c[k]); o.r = dist3(o.x, b); }
    else { for (int k = 0; k < 3; ++k) o.x[k] = (a[k] + s * (b[k] - a[k])) + t * (c[k] - a[k]); o.r = dist3(o.x, a); }
    return o;
}
template <class T> inline BBox<T> box_from_triangle(const T* a, const T* b, const T* c) {
    BBox<T> o;
    for (int k = 0; k < 3; ++k) { o.lo[k] = minimum3(a[k], b[k], c[k]); o.up[k] = maximum3(a[k], b[k], c[k]); }
    return o;
}

// ---------------------------------------------------------------------------------------------
// iscontact — src/bounding_volumes/iscontact.jl:2-28
// ---------------------------------------------------------------------------------------------
template <class T> inline bool iscontact(const BSphere<T>& a, const BSphere<T>& b) {
    return dist3sq(a.x, b.x) <= (a.r + b.r) * (a.r + b.r);
}
template <class T> inline bool iscontact(const BBox<T>& a, const BBox<T>& b) {
    return (a.up[0] >= b.lo[0] && a.lo[0] <= b.up[0]) &&
           (a.up[1] >= b.lo[1] && a.lo[1] <= b.up[1]) &&
           (a.up[2] >= b.lo[2] && a.lo[2] <= b.up[2]);
}
template <class T> inline bool iscontact(const BSphere<T>& a, const BBox<T>& b) {
    BBox<T> ab;
    for (int k = 0; k < 3; ++k) { ab.lo[k] = a.x[k] - a.r; ab.up[k] = a.x[k] + a.r; }
    return iscontact(ab, b);
}
template <class T> inline bool iscontact(const BBox<T>& a, const BSphere<T>& b) { return iscontact(b, a); }

// ---------------------------------------------------------------------------------------------
// isintersection — src/bounding_volumes/isintersection.jl:1-65
// ---------------------------------------------------------------------------------------------
template <class T> inline bool isintersection(const BBox<T>& b, const T p[3], const T d[3]) {
    T inv_d[3] = {T(1) / d[0], T(1) / d[1], T(1) / d[2]};
    T t1 = (b.lo[0] - p[0]) * inv_d[0];
    T t2 = (b.up[0] - p[0]) * inv_d[0];
    T tmin = minimum2(t1, t2);
    T tmax = maximum2(t1, t2);
    t1 = (b.lo[1] - p[1]) * inv_d[1];
    t2 = (b.up[1] - p[1]) * inv_d[1];
    tmin = maximum2(tmin, minimum2(t1, t2));
    tmax = minimum2(tmax, maximum2(t1, t2));
    t1 = (b.lo[2] - p[2]) * inv_d[2];
    t2 = (b.up[2] - p[2]) * inv_d[2];
    tmin = maximum2(tmin, minimum2(t1, t2));
    tmax = minimum2(tmax, maximum2(t1, t2));
    return (tmin <= tmax) && (tmax >= T(0));
}
template <class T> inline bool isintersection(const BSphere<T>& s, const T p[3], const T d[3]) {
    T a = dot3(d, d);
    T b = T(2) * (((p[0] - s.x[0]) * d[0] + (p[1] - s.x[1]) * d[1]) + (p[2] - s.x[2]) * d[2]);
    T c = (((p[0] - s.x[0]) * (p[0] - s.x[0]) + (p[1] - s.x[1]) * (p[1] - s.x[1])) +
           (p[2] - s.x[2]) * (p[2] - s.x[2])) - s.r * s.r;
    T disc = b * b - (T(4) * a) * c;
    if (disc >= T(0)) {
        if (b <= T(0)) return true;
        return T(0) >= c;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Implicit tree — src/implicit_tree.jl. Julia precedence: `a - b >> c` == `a - (b >> c)`.
// ---------------------------------------------------------------------------------------------
inline int ilog2_floor(uint64_t x) { return 63 - __builtin_clzll(x); }                 // utils.jl:131-133
inline int ilog2_ceil(uint64_t x) { return (x & (x - 1)) == 0 ? ilog2_floor(x) : ilog2_floor(x) + 1; }  // utils.jl:120

struct Tree {                                     // implicit_tree.jl:52-67
    int64_t levels, real_leaves, real_nodes, virtual_leaves, virtual_nodes;
};
inline Tree make_tree(int64_t n) {                // implicit_tree.jl:77-90 (caller checks n >= 1)
    Tree t;
    t.real_leaves = n;
    t.levels = ilog2_ceil((uint64_t)n) + 1;
    int64_t lv = (int64_t(1) << (t.levels - 1)) - n;
    t.virtual_leaves = lv;
    t.virtual_nodes = 2 * lv - __builtin_popcountll((uint64_t)lv);
    t.real_nodes = 2 * n - 1 + __builtin_popcountll((uint64_t)lv);
    return t;
}
// Julia `>>` with a shift count >= bit width yields 0 (no UB); guard it the same way.
inline int64_t shr(int64_t v, int64_t s) { return s >= 63 ? 0 : (v >> s); }
inline void compute_skips(const Tree& t, int64_t* skips) {   // implicit_tree.jl:100-113 (i is 1-based)
    for (int64_t i = 1; i <= t.levels; ++i) {
        int64_t v = shr(t.virtual_leaves, t.levels - (i - 1));
        skips[i - 1] = 2 * v - __builtin_popcountll((uint64_t)v);
    }
}
inline int64_t memory_index(const Tree& t, int64_t implicit_index) {  // implicit_tree.jl:128-148
    int64_t level = ilog2_floor((uint64_t)implicit_index) + 1;
    int64_t v = shr(t.virtual_leaves, t.levels - (level - 1));
    return implicit_index - (2 * v - __builtin_popcountll((uint64_t)v));
}
inline void level_indices(const Tree& t, int64_t level, int64_t* start, int64_t* stop) {  // :156-171
    *start = memory_index(t, int64_t(1) << (level - 1));
    int64_t nreal = (int64_t(1) << (level - 1)) - shr(t.virtual_leaves, t.levels - level);
    *stop = *start + nreal - 1;
}
inline bool isvirtual(const Tree& t, int64_t implicit_index) {        // implicit_tree.jl:191-199
    int64_t level = ilog2_floor((uint64_t)implicit_index) + 1;
    int64_t level_first = int64_t(1) << (level - 1);
    int64_t nreal = level_first - shr(t.virtual_leaves, t.levels - level);
    return implicit_index - level_first + 1 > nreal;
}
// build.jl:309-325 — Float path: round(I, levels + (1 - levels) * f), ties to even.
inline int64_t compute_build_level_float(int64_t levels, double f) {
    return (int64_t)std::nearbyint(double(levels) + double(1 - levels) * f);
}

// ---------------------------------------------------------------------------------------------
// Static contiguous partition over tasks — restates the *shape* of AK.itask_partition /
// AK.foreachindex on CPU threads (third-party AcceleratedKernels 0.4, not mounted): at most
// `max_tasks` tasks, at least `min_elems` items each, contiguous near-equal ranges.
// ---------------------------------------------------------------------------------------------
inline int num_tasks_for(int64_t n, int max_tasks, int64_t min_elems) {
    if (n <= 0) return 0;
    int64_t t = n / std::max<int64_t>(min_elems, 1);
    t = std::max<int64_t>(1, std::min<int64_t>(t, max_tasks));
    return (int)t;
}
inline void task_range(int64_t n, int ntasks, int itask, int64_t* lo, int64_t* hi) {   // 0-based [lo, hi)
    int64_t q = n / ntasks, r = n % ntasks;
    *lo = itask * q + std::min<int64_t>(itask, r);
    *hi = *lo + q + (itask < r ? 1 : 0);
}
template <class F> inline void parallel_tasks(int ntasks, F&& f) {
    if (ntasks <= 1) { if (ntasks == 1) f(0); return; }
    std::vector<std::thread> th;
    th.reserve(ntasks - 1);
    for (int t = 1; t < ntasks; ++t) th.emplace_back([&f, t] { f(t); });
    f(0);
    for (auto& x : th) x.join();
}

// ---------------------------------------------------------------------------------------------
// Morton — src/morton/default.jl, src/morton/utils.jl
// ---------------------------------------------------------------------------------------------
inline uint16_t morton_split3(uint16_t v) {       // default.jl:118-127
    uint16_t s = v & 0x001f;
    s = (s | (uint16_t)(s << 8)) & 0x100f;
    s = (s | (uint16_t)(s << 4)) & 0x10c3;
    s = (s | (uint16_t)(s << 2)) & 0x1249;
    return s;
}
inline uint32_t morton_split3(uint32_t v) {       // default.jl:130-143
    uint32_t s = v & 0x000003ffu;
    s = (s | s << 16) & 0x030000ffu;
    s = (s | s << 8) & 0x0300f00fu;
    s = (s | s << 4) & 0x030c30c3u;
    s = (s | s << 2) & 0x09249249u;
    return s;
}
inline uint64_t morton_split3(uint64_t v) {       // default.jl:146-157
    uint64_t s = v & 0x00000000001fffffull;
    s = (s | s << 32) & 0x001f00000000ffffull;
    s = (s | s << 16) & 0x001f0000ff0000ffull;
    s = (s | s << 8) & 0x100f00f00f00f00full;
    s = (s | s << 4) & 0x10c30c30c30c30c3ull;
    s = (s | s << 2) & 0x1249249249249249ull;
    return s;
}
template <class M> inline int morton_scaling();   // default.jl:167-169
template <> inline int morton_scaling<uint16_t>() { return 1 << 5; }
template <> inline int morton_scaling<uint32_t>() { return 1 << 10; }
template <> inline int morton_scaling<uint64_t>() { return 1 << 21; }
template <class T> inline T relative_precision(); // default.jl:179-181
template <> inline float relative_precision<float>() { return float(1e-5); }
template <> inline double relative_precision<double>() { return 1e-14; }

// morton/utils.jl:55-72 — padding, left to right: (m - rp*abs(m)) - floatmin.
template <class T> inline void pad_extrema(T mins[3], T maxs[3]) {
    const T rp = relative_precision<T>();
    const T fm = std::numeric_limits<T>::min();
    for (int k = 0; k < 3; ++k) {
        mins[k] = (mins[k] - rp * std::fabs(mins[k])) - fm;
        maxs[k] = (maxs[k] + rp * std::fabs(maxs[k])) + fm;
    }
}

// default.jl:91-108 — unsafe_trunc == C cast for in-range values.
template <class M, class T> inline M morton_encode_single(const T c[3], const T mins[3], const T maxs[3]) {
    const T scaling = T(morton_scaling<M>());
    T s1 = (c[0] - mins[0]) / (maxs[0] - mins[0]);
    T s2 = (c[1] - mins[1]) / (maxs[1] - mins[1]);
    T s3 = (c[2] - mins[2]) / (maxs[2] - mins[2]);
    M i1 = (M)(s1 * scaling), i2 = (M)(s2 * scaling), i3 = (M)(s3 * scaling);
    return (M)((M)(morton_split3(i1) << 2) | (M)(morton_split3(i2) << 1) | morton_split3(i3));
}

// ---------------------------------------------------------------------------------------------
// The wrapped leaf. `vol_t` typedef lets templates recover the float type.
// ---------------------------------------------------------------------------------------------
template <class V, class I, class M> struct Leaf {
    V volume; I index; M morton;
    using vol_t = V; using idx_t = I; using mor_t = M;
};

// morton/utils.jl:1-47 — min seeded with floatmax, **max seeded with floatmin (smallest positive
// normal)**, comparisons `a < b ? a : b`. Sequential left fold (the reduction is order-independent
// for non-NaN input).
template <class L> inline void compute_extrema(const L* leaves, int64_t n,
                                               typename L::vol_t::value_type mins[3],
                                               typename L::vol_t::value_type maxs[3]) {
    using T = typename L::vol_t::value_type;
    for (int k = 0; k < 3; ++k) { mins[k] = std::numeric_limits<T>::max(); maxs[k] = std::numeric_limits<T>::min(); }
    for (int64_t i = 0; i < n; ++i) {
        T c[3];
        center(leaves[i].volume, c);
        for (int k = 0; k < 3; ++k) {
            mins[k] = mins[k] < c[k] ? mins[k] : c[k];
            maxs[k] = maxs[k] > c[k] ? maxs[k] : c[k];
        }
    }
}

// default.jl:43-82 — morton_encode!: in place, all leaves rewritten with the new code.
// compute_extrema=false: the reference reads `options.mins`, a field that does not exist
// (default.jl:55-56) and would throw; we treat caller bounds as "use as given, unpadded" and
// flag it as an extension.
template <class L>
inline void morton_encode(L* leaves, int64_t n, bool compute_ext,
                          typename L::vol_t::value_type mins[3], typename L::vol_t::value_type maxs[3],
                          int num_threads, int64_t min_elems) {
    using T = typename L::vol_t::value_type;
    using M = typename L::mor_t;
    if (n == 0) return;
    if (compute_ext) {
        int nt = num_tasks_for(n, num_threads, min_elems);
        std::vector<T> pm(3 * nt), pM(3 * nt);
        parallel_tasks(nt, [&](int t) {
            int64_t lo, hi; task_range(n, nt, t, &lo, &hi);
            compute_extrema(leaves + lo, hi - lo, &pm[3 * t], &pM[3 * t]);
        });
        for (int k = 0; k < 3; ++k) { mins[k] = pm[k]; maxs[k] = pM[k]; }
        for (int t = 1; t < nt; ++t)
            for (int k = 0; k < 3; ++k) {
                mins[k] = mins[k] < pm[3 * t + k] ? mins[k] : pm[3 * t + k];
                maxs[k] = maxs[k] > pM[3 * t + k] ? maxs[k] : pM[3 * t + k];
            }
        pad_extrema(mins, maxs);
    }
    int nt = num_tasks_for(n, num_threads, min_elems);
    parallel_tasks(nt, [&](int t) {
        int64_t lo, hi; task_range(n, nt, t, &lo, &hi);
        for (int64_t i = lo; i < hi; ++i) {
            T c[3];
            center(leaves[i].volume, c);
            leaves[i].morton = morton_encode_single<M>(c, mins, maxs);
        }
    });
}

// build.jl:248-253 — AK.sort!(by = morton). STABLE (tie order is unpinned by the reference).
// Parallel structure: per-task stable_sort of contiguous chunks + pairwise stable merges over
// whole structs (what a CPU merge/sample sort of 24-byte structs costs).
template <class L> inline void sort_leaves(L* leaves, int64_t n, int num_threads, int64_t min_elems) {
    auto cmp = [](const L& a, const L& b) { return a.morton < b.morton; };
    int nt = num_tasks_for(n, num_threads, min_elems);
    if (nt <= 1) { std::stable_sort(leaves, leaves + n, cmp); return; }
    std::vector<int64_t> bounds(nt + 1);
    for (int t = 0; t < nt; ++t) { int64_t lo, hi; task_range(n, nt, t, &lo, &hi); bounds[t] = lo; bounds[t + 1] = hi; }
    parallel_tasks(nt, [&](int t) { std::stable_sort(leaves + bounds[t], leaves + bounds[t + 1], cmp); });
    std::vector<L> tmp(n);
    L* src = leaves; L* dst = tmp.data();
    std::vector<int64_t> b = bounds;
    while ((int)b.size() > 2) {
        int runs = (int)b.size() - 1;
        int pairs = runs / 2;
        std::vector<int64_t> nb;
        for (int p = 0; p < pairs; ++p) nb.push_back(b[2 * p]);
        if (runs % 2) nb.push_back(b[runs - 1]);
        nb.push_back(b[runs]);
        parallel_tasks(pairs + (runs % 2), [&](int p) {
            if (p < pairs) std::merge(src + b[2 * p], src + b[2 * p + 1], src + b[2 * p + 1], src + b[2 * p + 2], dst + b[2 * p], cmp);
            else std::copy(src + b[runs - 1], src + b[runs], dst + b[runs - 1]);
        });
        std::swap(src, dst);
        b = nb;
    }
    if (src != leaves) std::memcpy((void*)leaves, (const void*)src, sizeof(L) * (size_t)n);
}

// ---------------------------------------------------------------------------------------------
// Aggregation — build.jl:366-523. All positions 1-based as in the reference.
// ---------------------------------------------------------------------------------------------
template <class L, class N> struct SameLeafNode { static constexpr bool value = std::is_same<typename L::vol_t, N>::value; };

template <class V, class N> inline N leaf_to_node(const V& v) {
    if constexpr (std::is_same<V, N>::value) return v; else return NodeOps<N>::convert(v);
}
template <class V, class N> inline N leaves_to_node(const V& a, const V& b) { return NodeOps<N>::merge(a, b); }

template <class L, class N>
inline void aggregate(N* nodes, const L* leaves, const Tree& tree, int64_t built_level,
                      int num_threads, int64_t min_elems) {
    using V = typename L::vol_t;
    // build.jl:381-457 — level above the leaves
    {
        int64_t level = tree.levels - 1;
        int64_t start_pos = memory_index(tree, int64_t(1) << (level - 1));
        int64_t num_nodes = (int64_t(1) << (level - 1)) - (tree.virtual_leaves >> 1);
        int64_t num_nodes_next = tree.real_leaves;
        int nt = num_tasks_for(num_nodes, num_threads, min_elems);
        parallel_tasks(nt, [&](int t) {
            int64_t lo, hi; task_range(num_nodes, nt, t, &lo, &hi);
            for (int64_t i = lo + 1; i <= hi; ++i) {
                int64_t l = 2 * i - 1, r = 2 * i;
                if (r > num_nodes_next) nodes[start_pos - 1 + i - 1] = leaf_to_node<V, N>(leaves[l - 1].volume);
                else nodes[start_pos - 1 + i - 1] = leaves_to_node<V, N>(leaves[l - 1].volume, leaves[r - 1].volume);
            }
        });
    }
    // build.jl:371-375, 460-523 — remaining levels up to built_level
    for (int64_t level = tree.levels - 2; level >= built_level; --level) {
        int64_t start_pos = memory_index(tree, int64_t(1) << (level - 1));
        int64_t num_nodes = (int64_t(1) << (level - 1)) - shr(tree.virtual_leaves, tree.levels - level);
        int64_t start_pos_next = memory_index(tree, int64_t(1) << level);
        int64_t num_nodes_next = (int64_t(1) << level) - shr(tree.virtual_leaves, tree.levels - (level + 1));
        int nt = num_tasks_for(num_nodes, num_threads, min_elems);
        parallel_tasks(nt, [&](int t) {
            int64_t lo, hi; task_range(num_nodes, nt, t, &lo, &hi);
            for (int64_t i = lo + 1; i <= hi; ++i) {
                int64_t l = start_pos_next + 2 * i - 2, r = start_pos_next + 2 * i - 1;
                if (r > start_pos_next + num_nodes_next - 1) nodes[start_pos - 1 + i - 1] = nodes[l - 1];
                else nodes[start_pos - 1 + i - 1] = NodeOps<N>::merge(nodes[l - 1], nodes[r - 1]);
            }
        });
    }
}

// build.jl:198-271 — the constructor minus allocation: encode, sort, aggregate.
template <class L, class N>
inline void build(L* leaves, int64_t n, N* nodes, int64_t built_level, bool compute_ext,
                  typename L::vol_t::value_type mins[3], typename L::vol_t::value_type maxs[3],
                  int num_threads, int64_t min_mortons, int64_t min_sorts, int64_t min_boundings) {
    Tree tree = make_tree(n);
    morton_encode(leaves, n, compute_ext, mins, maxs, num_threads, min_mortons);
    sort_leaves(leaves, n, num_threads, min_sorts);
    if (tree.real_nodes >= 2) aggregate<L, N>(nodes, leaves, tree, built_level, num_threads, min_boundings);
}

// ---------------------------------------------------------------------------------------------
// LVT traversals — src/traverse/leaf_vs_tree/*.jl, src/raytrace/leaf_vs_tree/leaf_vs_tree.jl
// A BVH "view": pointers + tree + skips + built_level (build.jl:155-166).
// ---------------------------------------------------------------------------------------------
template <class L, class N> struct BVHView {
    Tree tree; const int64_t* skips; const N* nodes; const L* leaves; int64_t built_level;
};

// Emission target shared by the count pass (contacts == nullptr) and the write pass.
template <class I> struct Emit {
    IndexPair<I>* contacts;   // nullptr on the counting pass
    int64_t iwrite;           // 0-based next write slot (write pass)
    int64_t count;            // contacts seen by this task (count pass)
};

// traverse_single.jl:136-208 — one query leaf (1-based sorted position `ileaf`) against its own tree.
template <class L, class N, class I>
inline void traverse_lvt_single(const L& bv, int64_t ileaf, const BVHView<L, N>& bvh, int64_t start_level, Emit<I>& em) {
    using V = typename L::vol_t;
    const Tree& tr = bvh.tree;
    int64_t stack[64];
    int64_t inode_start = int64_t(1) << (start_level - 1);
    int64_t level_num_real = (int64_t(1) << (start_level - 1)) - shr(tr.virtual_leaves, tr.levels - start_level);
    int64_t inode_end = inode_start + level_num_real - 1;
    N bv_node = leaf_to_node<V, N>(bv.volume);
    for (int64_t inode_root = inode_start; inode_root <= inode_end; ++inode_root) {
        int64_t istack = 0, inode = inode_root;
        while (true) {
            int64_t ilevel = ilog2_floor((uint64_t)inode) + 1;
            int64_t irightmost = ((inode + 1) << (tr.levels - ilevel)) - 1;
            bool descended = false;
            if (irightmost <= ileaf + (int64_t(1) << (tr.levels - 1)) - 1) {
                // subtree fully to the left of (or at) the query: skip
            } else if (ilevel == tr.levels) {
                const L& leaf = bvh.leaves[inode - (int64_t(1) << (tr.levels - 1)) + 1 - 1];
                if (iscontact(bv.volume, leaf.volume)) {
                    if (!em.contacts) em.count += 1;
                    else {
                        if (bv.index > leaf.index) em.contacts[em.iwrite] = {leaf.index, bv.index};
                        else em.contacts[em.iwrite] = {bv.index, leaf.index};
                        em.iwrite += 1;
                    }
                }
            } else {
                const N& node = bvh.nodes[inode - bvh.skips[ilevel - 1] - 1];
                if (iscontact(bv_node, node)) {
                    if (!isvirtual(tr, 2 * inode + 1)) stack[istack++] = 2 * inode + 1;
                    inode = 2 * inode;
                    descended = true;
                }
            }
            if (descended) continue;
            if (istack == 0) break;
            inode = stack[--istack];
        }
    }
}

// traverse_pair.jl:176-244 — one query leaf of bvh1 against bvh2's tree; no skip rule.
template <class L, class L2, class N, class I>
inline void traverse_lvt_pair(const L& bv, const BVHView<L2, N>& bvh, int64_t start_level, bool flip, Emit<I>& em) {
    using V = typename L::vol_t;
    const Tree& tr = bvh.tree;
    int64_t stack[64];
    int64_t inode_start = int64_t(1) << (start_level - 1);
    int64_t level_num_real = (int64_t(1) << (start_level - 1)) - shr(tr.virtual_leaves, tr.levels - start_level);
    int64_t inode_end = inode_start + level_num_real - 1;
    N bv_node = leaf_to_node<V, N>(bv.volume);
    for (int64_t inode_root = inode_start; inode_root <= inode_end; ++inode_root) {
        int64_t istack = 0, inode = inode_root;
        while (true) {
            int64_t ilevel = ilog2_floor((uint64_t)inode) + 1;
            bool descended = false;
            if (ilevel == tr.levels) {
                const L2& leaf = bvh.leaves[inode - (int64_t(1) << (tr.levels - 1)) + 1 - 1];
                if (iscontact(bv.volume, leaf.volume)) {
                    if (!em.contacts) em.count += 1;
                    else {
                        if (flip) em.contacts[em.iwrite] = {(I)leaf.index, (I)bv.index};
                        else em.contacts[em.iwrite] = {(I)bv.index, (I)leaf.index};
                        em.iwrite += 1;
                    }
                }
            } else {
                const N& node = bvh.nodes[inode - bvh.skips[ilevel - 1] - 1];
                if (iscontact(bv_node, node)) {
                    if (!isvirtual(tr, 2 * inode + 1)) stack[istack++] = 2 * inode + 1;
                    inode = 2 * inode;
                    descended = true;
                }
            }
            if (descended) continue;
            if (istack == 0) break;
            inode = stack[--istack];
        }
    }
}

// raytrace/leaf_vs_tree/leaf_vs_tree.jl:170-228 — one ray (1-based id `iray`).
template <class L, class N, class I, class T>
inline void traverse_ray_lvt(const T point[3], const T dir[3], int64_t iray, const BVHView<L, N>& bvh,
                             int64_t start_level, Emit<I>& em) {
    const Tree& tr = bvh.tree;
    int64_t stack[64];
    int64_t inode_start = int64_t(1) << (start_level - 1);
    int64_t level_num_real = (int64_t(1) << (start_level - 1)) - shr(tr.virtual_leaves, tr.levels - start_level);
    int64_t inode_end = inode_start + level_num_real - 1;
    for (int64_t inode_root = inode_start; inode_root <= inode_end; ++inode_root) {
        int64_t istack = 0, inode = inode_root;
        while (true) {
            int64_t ilevel = ilog2_floor((uint64_t)inode) + 1;
            bool descended = false;
            if (ilevel == tr.levels) {
                const L& leaf = bvh.leaves[inode - (int64_t(1) << (tr.levels - 1)) + 1 - 1];
                if (isintersection(leaf.volume, point, dir)) {
                    if (!em.contacts) em.count += 1;
                    else { em.contacts[em.iwrite] = {(I)leaf.index, (I)iray}; em.iwrite += 1; }
                }
            } else {
                const N& node = bvh.nodes[inode - bvh.skips[ilevel - 1] - 1];
                if (isintersection(node, point, dir)) {
                    if (!isvirtual(tr, 2 * inode + 1)) stack[istack++] = 2 * inode + 1;
                    inode = 2 * inode;
                    descended = true;
                }
            }
            if (descended) continue;
            if (istack == 0) break;
            inode = stack[--istack];
        }
    }
}

// Two-pass driver shared by the three traversals — leaf_vs_tree/traverse_single.jl:1-79,
// traverse_pair.jl:40-116, raytrace/leaf_vs_tree/leaf_vs_tree.jl:1-90 (CPU backend: one counter
// per task, count pass -> inclusive scan -> write pass; tasks are contiguous ascending ranges, so
// the output order is "ascending query, then DFS order", identical to the GPU backend's).
// `per_query(q0, emit)` runs query q0 (0-based). Returns the total; fills `contacts` (if non-null
// and capacity suffices) and `per_query_counts` (if non-null: inclusive scan of per-query counts,
// the GPU-backend form of cache2, traverse_single.jl:31,57).
template <class I, class F>
inline int64_t two_pass(int64_t nqueries, int num_threads, int64_t min_elems, IndexPair<I>* contacts,
                        int64_t capacity, I* per_query_counts, F&& per_query) {
    int nt = num_tasks_for(nqueries, num_threads, min_elems);
    if (nt == 0) return 0;
    std::vector<int64_t> task_counts(nt, 0);
    parallel_tasks(nt, [&](int t) {
        int64_t lo, hi; task_range(nqueries, nt, t, &lo, &hi);
        Emit<I> em{nullptr, 0, 0};
        for (int64_t q = lo; q < hi; ++q) {
            int64_t before = em.count;
            per_query(q, em);
            if (per_query_counts) per_query_counts[q] = (I)(em.count - before);
        }
        task_counts[t] = em.count;
    });
    for (int t = 1; t < nt; ++t) task_counts[t] += task_counts[t - 1];        // AK.accumulate!(+)
    int64_t total = task_counts[nt - 1];
    if (per_query_counts) for (int64_t q = 1; q < nqueries; ++q) per_query_counts[q] = (I)(per_query_counts[q] + per_query_counts[q - 1]);
    if (!contacts || total == 0 || total > capacity) return total;
    parallel_tasks(nt, [&](int t) {
        int64_t lo, hi; task_range(nqueries, nt, t, &lo, &hi);
        Emit<I> em{contacts, t == 0 ? 0 : task_counts[t - 1], 0};
        for (int64_t q = lo; q < hi; ++q) per_query(q, em);
    });
    return total;
}

// ---------------------------------------------------------------------------------------------
// BFS traversals — src/traverse/breadth_first/*.jl, src/raytrace/breadth_first/*.jl (CPU bodies).
// The reference's task-parallel form gives every task a contiguous range of the source list and
// compacts the tasks' outputs in task order (traverse_single_cpu.jl:28-56), so its lists are the
// ones this sequential loop produces, whatever the thread count. (The reference's GPU backend
// appends with atomics: same lists as sets, order unspecified.)
// A BVTT entry is a pair of implicit indices (or implicit index, ray id), 1-based as in Julia.
// ---------------------------------------------------------------------------------------------
struct BvttPair { int64_t a, b; };

// Node-vs-leaf-volume tests of the pair traversal's special cases (traverse_pair_cpu.jl, the
// `traverse_nodes_leaves_*` bodies): Julia promotes mixed float types operation by operation, which is
// the same as converting both operands to the wider type first (the conversion is exact).
template <class TA, class TB> inline bool iscontact_promoted(const BBox<TA>& a, const BBox<TB>& b) {
    using C = typename std::common_type<TA, TB>::type;
    return ((C)a.up[0] >= (C)b.lo[0] && (C)a.lo[0] <= (C)b.up[0]) &&
           ((C)a.up[1] >= (C)b.lo[1] && (C)a.lo[1] <= (C)b.up[1]) &&
           ((C)a.up[2] >= (C)b.lo[2] && (C)a.lo[2] <= (C)b.up[2]);
}
template <class TA, class TB> inline bool iscontact_promoted(const BSphere<TA>& a, const BSphere<TB>& b) {
    using C = typename std::common_type<TA, TB>::type;
    C ax[3] = {(C)a.x[0], (C)a.x[1], (C)a.x[2]}, bx[3] = {(C)b.x[0], (C)b.x[1], (C)b.x[2]};
    return dist3sq(ax, bx) <= ((C)a.r + (C)b.r) * ((C)a.r + (C)b.r);
}
template <class TA, class TB> inline bool iscontact_promoted(const BSphere<TA>& a, const BBox<TB>& b) {   // iscontact.jl:17-23
    BBox<TA> ab;
    for (int k = 0; k < 3; ++k) { ab.lo[k] = a.x[k] - a.r; ab.up[k] = a.x[k] + a.r; }
    return iscontact_promoted(ab, b);
}
template <class TA, class TB> inline bool iscontact_promoted(const BBox<TA>& a, const BSphere<TB>& b) { return iscontact_promoted(b, a); }

// traverse_single.jl:79-157 (fill_initial_bvtt_single!, CPU branch)
inline void bfs_initial_single(const Tree& tr, int64_t start_level, std::vector<BvttPair>& out) {
    int64_t level_nodes = int64_t(1) << (start_level - 1);
    int64_t num_real = level_nodes - shr(tr.virtual_leaves, tr.levels - start_level);
    out.clear();
    for (int64_t i = level_nodes; i <= level_nodes + num_real - 1; ++i) {
        if (start_level != tr.levels) out.push_back({i, i});
        for (int64_t j = i + 1; j <= level_nodes + num_real - 1; ++j) out.push_back({i, j});
    }
}

// traverse_single_cpu.jl:62-128 (traverse_nodes_range!)
template <class N>
inline void bfs_nodes_single(const Tree& tr, const N* nodes, const std::vector<BvttPair>& src, std::vector<BvttPair>& dst,
                             int64_t level, bool self_checks) {
    int64_t vl = shr(tr.virtual_leaves, tr.levels - (level - 1));
    int64_t num_skips = 2 * vl - __builtin_popcountll((uint64_t)vl);
    dst.clear();
    for (const BvttPair& e : src) {
        int64_t i1 = e.a, i2 = e.b;
        if (i1 == i2) {
            if (isvirtual(tr, 2 * i1 + 1)) {
                if (self_checks) dst.push_back({2 * i1, 2 * i1});
            } else if (self_checks) {
                dst.push_back({2 * i1, 2 * i1});
                dst.push_back({2 * i1, 2 * i1 + 1});
                dst.push_back({2 * i1 + 1, 2 * i1 + 1});
            } else {
                dst.push_back({2 * i1, 2 * i1 + 1});
            }
        } else if (iscontact(nodes[i1 - num_skips - 1], nodes[i2 - num_skips - 1])) {
            if (isvirtual(tr, 2 * i2 + 1)) {
                dst.push_back({2 * i1, 2 * i2});
                dst.push_back({2 * i1 + 1, 2 * i2});
            } else {
                dst.push_back({2 * i1, 2 * i2});
                dst.push_back({2 * i1, 2 * i2 + 1});
                dst.push_back({2 * i1 + 1, 2 * i2});
                dst.push_back({2 * i1 + 1, 2 * i2 + 1});
            }
        }
    }
}

// traverse(bvh, BFSTraversal()) — traverse_single.jl:1-66, traverse_single_cpu.jl:179-219.
// positions: emit (leaf position 1, leaf position 2) instead of the sorted index pair.
template <class L, class N, class I>
inline int64_t traverse_bfs_single(const BVHView<L, N>& bvh, int64_t start_level, bool positions,
                                   std::vector<IndexPair<I>>& contacts, int64_t* num_checks) {
    const Tree& tr = bvh.tree;
    contacts.clear();
    *num_checks = 0;
    if (tr.real_nodes <= 1) return 0;
    std::vector<BvttPair> a, b;
    bfs_initial_single(tr, start_level, a);
    int64_t checks = (int64_t)a.size();
    for (int64_t level = start_level; level < tr.levels; ++level) {
        bool self_checks = level < tr.levels - 1;
        bfs_nodes_single<N>(tr, bvh.nodes, a, b, level, self_checks);
        checks += (int64_t)b.size();
        a.swap(b);
    }
    int64_t num_above = (int64_t(1) << (tr.levels - 1)) - 1;
    for (const BvttPair& e : a) {
        const L& l1 = bvh.leaves[e.a - num_above - 1];
        const L& l2 = bvh.leaves[e.b - num_above - 1];
        if (iscontact(l1.volume, l2.volume)) {
            if (positions) contacts.push_back({(I)(e.a - num_above), (I)(e.b - num_above)});
            else if (l1.index > l2.index) contacts.push_back({l2.index, l1.index});
            else contacts.push_back({l1.index, l2.index});
        }
    }
    *num_checks = checks;
    return (int64_t)contacts.size();
}

// traverse(bvh1, bvh2, BFSTraversal()) — traverse_pair.jl:1-158, traverse_pair_cpu.jl. mode: which side sprouts.
enum BfsPairMode { kBfsBoth, kBfsLeft, kBfsRight, kBfsLeafLeft, kBfsLeafRight };
template <class L, class N>
inline void bfs_nodes_pair(const BVHView<L, N>& b1, const BVHView<L, N>& b2, const std::vector<BvttPair>& src,
                           std::vector<BvttPair>& dst, int64_t level1, int64_t level2, BfsPairMode mode) {
    const Tree &t1 = b1.tree, &t2 = b2.tree;
    int64_t v1 = shr(t1.virtual_leaves, t1.levels - (level1 - 1)), v2 = shr(t2.virtual_leaves, t2.levels - (level2 - 1));
    int64_t skips1 = 2 * v1 - __builtin_popcountll((uint64_t)v1), skips2 = 2 * v2 - __builtin_popcountll((uint64_t)v2);
    int64_t above1 = (int64_t(1) << (t1.levels - 1)) - 1, above2 = (int64_t(1) << (t2.levels - 1)) - 1;
    dst.clear();
    for (const BvttPair& e : src) {
        int64_t i1 = e.a, i2 = e.b;
        bool hit;
        if (mode == kBfsLeafLeft) hit = iscontact_promoted(b1.nodes[i1 - skips1 - 1], b2.leaves[i2 - above2 - 1].volume);
        else if (mode == kBfsLeafRight) hit = iscontact_promoted(b1.leaves[i1 - above1 - 1].volume, b2.nodes[i2 - skips2 - 1]);
        else hit = iscontact(b1.nodes[i1 - skips1 - 1], b2.nodes[i2 - skips2 - 1]);
        if (!hit) continue;
        bool r1 = !isvirtual(t1, 2 * i1 + 1), r2 = !isvirtual(t2, 2 * i2 + 1);
        if (mode == kBfsBoth) {
            dst.push_back({2 * i1, 2 * i2});
            if (r2) dst.push_back({2 * i1, 2 * i2 + 1});
            if (r1) dst.push_back({2 * i1 + 1, 2 * i2});
            if (r1 && r2) dst.push_back({2 * i1 + 1, 2 * i2 + 1});
        } else if (mode == kBfsLeft || mode == kBfsLeafLeft) {
            dst.push_back({2 * i1, i2});
            if (r1) dst.push_back({2 * i1 + 1, i2});
        } else {
            dst.push_back({i1, 2 * i2});
            if (r2) dst.push_back({i1, 2 * i2 + 1});
        }
    }
}

template <class L, class N, class I>
inline int64_t traverse_bfs_pair(const BVHView<L, N>& b1, const BVHView<L, N>& b2, int64_t start_level1, int64_t start_level2,
                                 bool positions, std::vector<IndexPair<I>>& contacts, int64_t* num_checks) {
    const Tree &t1 = b1.tree, &t2 = b2.tree;
    contacts.clear();
    std::vector<BvttPair> a, b;
    // initial_bvtt / fill_initial_bvtt_pair!, traverse_pair.jl:161-221
    int64_t ln1 = int64_t(1) << (start_level1 - 1), ln2 = int64_t(1) << (start_level2 - 1);
    int64_t nr1 = ln1 - shr(t1.virtual_leaves, t1.levels - start_level1), nr2 = ln2 - shr(t2.virtual_leaves, t2.levels - start_level2);
    for (int64_t i = ln1; i <= ln1 + nr1 - 1; ++i) for (int64_t j = ln2; j <= ln2 + nr2 - 1; ++j) a.push_back({i, j});
    int64_t checks = (int64_t)a.size();
    int64_t level1 = start_level1, level2 = start_level2;
    auto step = [&](BfsPairMode m) { bfs_nodes_pair<L, N>(b1, b2, a, b, level1, level2, m); checks += (int64_t)b.size(); a.swap(b); };
    while (level1 < t1.levels - 1 && level2 < t2.levels - 1) { step(kBfsBoth); ++level1; ++level2; }
    while (level1 < t1.levels - 1 && level2 == t2.levels - 1) { step(kBfsLeft); ++level1; }
    while (level2 < t2.levels - 1 && level1 == t1.levels - 1) { step(kBfsRight); ++level2; }
    while (level2 == t2.levels && level1 < t1.levels) { step(kBfsLeafLeft); ++level1; }
    while (level1 == t1.levels && level2 < t2.levels) { step(kBfsLeafRight); ++level2; }
    if (level1 == t1.levels - 1 && level2 == t2.levels - 1) { step(kBfsBoth); ++level1; ++level2; }
    int64_t above1 = (int64_t(1) << (t1.levels - 1)) - 1, above2 = (int64_t(1) << (t2.levels - 1)) - 1;
    for (const BvttPair& e : a) {                                                      // traverse_leaves_pair!
        const L& l1 = b1.leaves[e.a - above1 - 1];
        const L& l2 = b2.leaves[e.b - above2 - 1];
        if (iscontact(l1.volume, l2.volume)) {
            if (positions) contacts.push_back({(I)(e.a - above1), (I)(e.b - above2)});
            else contacts.push_back({l1.index, l2.index});
        }
    }
    *num_checks = checks;
    return (int64_t)contacts.size();
}

// traverse_rays(bvh, points, directions, BFSTraversal()) — raytrace/breadth_first/breadth_first.jl:1-66, raytrace_cpu.jl
template <class L, class N, class I, class T>
inline int64_t traverse_bfs_rays(const BVHView<L, N>& bvh, const T* points, const T* dirs, int64_t nrays, int64_t start_level,
                                 bool positions, std::vector<IndexPair<I>>& hits, int64_t* num_checks) {
    const Tree& tr = bvh.tree;
    hits.clear();
    *num_checks = 0;
    if (nrays == 0) return 0;
    std::vector<BvttPair> a, b;
    int64_t ln = int64_t(1) << (start_level - 1);
    int64_t nr = ln - shr(tr.virtual_leaves, tr.levels - start_level);
    for (int64_t i = ln; i <= ln + nr - 1; ++i) for (int64_t j = 1; j <= nrays; ++j) a.push_back({i, j});
    int64_t checks = (int64_t)a.size();
    for (int64_t level = start_level; level < tr.levels; ++level) {
        int64_t vl = shr(tr.virtual_leaves, tr.levels - (level - 1));
        int64_t num_skips = 2 * vl - __builtin_popcountll((uint64_t)vl);
        b.clear();
        for (const BvttPair& e : a) {
            if (isintersection(bvh.nodes[e.a - num_skips - 1], points + 3 * (e.b - 1), dirs + 3 * (e.b - 1))) {
                b.push_back({2 * e.a, e.b});
                if (!isvirtual(tr, 2 * e.a + 1)) b.push_back({2 * e.a + 1, e.b});
            }
        }
        checks += (int64_t)b.size();
        a.swap(b);
    }
    int64_t num_above = (int64_t(1) << (tr.levels - 1)) - 1;
    for (const BvttPair& e : a) {
        const L& leaf = bvh.leaves[e.a - num_above - 1];
        if (isintersection(leaf.volume, points + 3 * (e.b - 1), dirs + 3 * (e.b - 1))) {
            if (positions) hits.push_back({(I)(e.a - num_above), (I)e.b});
            else hits.push_back({(I)leaf.index, (I)e.b});
        }
    }
    *num_checks = checks;
    return (int64_t)hits.size();
}

}  // namespace orc
