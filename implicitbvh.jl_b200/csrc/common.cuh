// common.cuh — device-side restatement of the reference's scalar layer (L1–L3 in SURVEY.md §1):
// isbits layouts, merges, contact / ray predicates and implicit-tree index math.
//
// Bit-exactness contract (SURVEY.md §7 "hard parts"): this translation unit set is compiled with
// -fmad=false (Julia never contracts a*b+c), default -prec-div=true -prec-sqrt=true -ftz=false, and
// every expression keeps the reference's left-to-right association.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ibvh.h"

#define IBVH_HD __host__ __device__ __forceinline__
#define IBVH_D __device__ __forceinline__

namespace ibvh {

// ---- layouts (bsphere.jl:26-29, bbox.jl:35-38, bounding_volumes.jl:55-59, traverse.jl:6) ----
template <class T> struct BSphere { T x[3]; T r; using value_type = T; static constexpr int kind = IBVH_BSPHERE; };
template <class T> struct BBox    { T lo[3]; T up[3]; using value_type = T; static constexpr int kind = IBVH_BBOX; };
template <class V, class I, class M> struct Leaf {
    V volume; I index; M morton;
    using vol_t = V; using idx_t = I; using mor_t = M; using value_type = typename V::value_type;
};
// aligned to its size (8 / 16 bytes): one vector store per contact. Every element of a cudaMalloc'ed / CuArray / torch
// IndexPair array is; the entry points reject a misaligned cache1 (IBVH_ERR_ARGUMENT) instead of faulting.
template <class I> struct alignas(2 * sizeof(I)) IndexPair { I a, b; };

static_assert(sizeof(Leaf<BSphere<float>, int32_t, uint32_t>) == 24, "layout");
static_assert(sizeof(Leaf<BSphere<float>, int32_t, uint16_t>) == 24, "layout");
static_assert(sizeof(Leaf<BSphere<float>, int64_t, uint64_t>) == 32, "layout");
static_assert(sizeof(Leaf<BBox<float>, int32_t, uint32_t>) == 32, "layout");
static_assert(sizeof(Leaf<BBox<float>, int64_t, uint64_t>) == 40, "layout");

// ---- raw-word view of a leaf: moving leaves as words keeps every byte (padding included) intact ----
template <class L> struct Words {
    static_assert(sizeof(L) % 8 == 0, "leaf structs are 8-byte multiples");
    static constexpr int N8 = sizeof(L) / 8;
    uint2 w[N8];
};
template <class L> IBVH_D Words<L> load_words(const L* p) {
    Words<L> o;
    const uint2* s = reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int k = 0; k < Words<L>::N8; ++k) o.w[k] = s[k];
    return o;
}
template <class L> IBVH_D void store_words(L* p, const Words<L>& v) {
    uint2* d = reinterpret_cast<uint2*>(p);
#pragma unroll
    for (int k = 0; k < Words<L>::N8; ++k) d[k] = v.w[k];
}
template <class L> IBVH_D Words<L> zero_words() {
    Words<L> o;
#pragma unroll
    for (int k = 0; k < Words<L>::N8; ++k) o.w[k] = make_uint2(0u, 0u);
    return o;
}
template <class L> IBVH_D typename L::vol_t words_volume(const Words<L>& v) {
    typename L::vol_t o;
    memcpy(&o, v.w, sizeof(o));
    return o;
}
template <class L> IBVH_D typename L::idx_t words_index(const Words<L>& v) {
    typename L::idx_t o;
    memcpy(&o, reinterpret_cast<const char*>(v.w) + offsetof(L, index), sizeof(o));
    return o;
}
template <class L> IBVH_D typename L::mor_t words_morton(const Words<L>& v) {
    typename L::mor_t o;
    memcpy(&o, reinterpret_cast<const char*>(v.w) + offsetof(L, morton), sizeof(o));
    return o;
}
template <class L> IBVH_D void words_set_volume(Words<L>& v, const typename L::vol_t& x) { memcpy(v.w, &x, sizeof(x)); }
template <class L> IBVH_D void words_set_index(Words<L>& v, typename L::idx_t x) {
    memcpy(reinterpret_cast<char*>(v.w) + offsetof(L, index), &x, sizeof(x));
}
template <class L> IBVH_D void words_set_morton(Words<L>& v, typename L::mor_t x) {
    memcpy(reinterpret_cast<char*>(v.w) + offsetof(L, morton), &x, sizeof(x));
}

// One raw volume with the widest loads its size and the array's alignment allow (`vec` = 16, 8 or 4, chosen by the
// host from the pointer): a struct of floats only promises 4-byte alignment, and copying it into the leaf words
// through memcpy made nvcc emit one LDG.U8 per byte — 16 loads per sphere.
template <class V> IBVH_D V load_volume(const V* p, int vec) {
    alignas(16) V x;
    if (sizeof(V) % 16 == 0 && vec == 16) {
        const uint4* s = reinterpret_cast<const uint4*>(p);
        uint4* d = reinterpret_cast<uint4*>(&x);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / 16); ++k) d[k] = __ldg(s + k);
    } else if (sizeof(V) % 8 == 0 && vec >= 8) {
        const uint2* s = reinterpret_cast<const uint2*>(p);
        uint2* d = reinterpret_cast<uint2*>(&x);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / 8); ++k) d[k] = __ldg(s + k);
    } else {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(p);
        uint32_t* d = reinterpret_cast<uint32_t*>(&x);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(V) / 4); ++k) d[k] = __ldg(s + k);
    }
    return x;
}

// ---- utils.jl:163-181 -------------------------------------------------------------------------
template <class T> IBVH_HD T minimum2(T a, T b) { return a < b ? a : b; }
template <class T> IBVH_HD T maximum2(T a, T b) { return a > b ? a : b; }
template <class T> IBVH_HD T dist3sq(const T* x, const T* y) {
    T d0 = (x[0] - y[0]) * (x[0] - y[0]);
    T d1 = (x[1] - y[1]) * (x[1] - y[1]);
    T d2 = (x[2] - y[2]) * (x[2] - y[2]);
    return (d0 + d1) + d2;
}
IBVH_HD float ibvh_sqrt(float x) { return sqrtf(x); }
IBVH_HD double ibvh_sqrt(double x) { return sqrt(x); }
IBVH_HD float ibvh_abs(float x) { return fabsf(x); }
IBVH_HD double ibvh_abs(double x) { return fabs(x); }

// ---- center (bsphere.jl:142, bbox.jl:100-102) ---------------------------------------------------
template <class T> IBVH_HD void center(const BSphere<T>& b, T c[3]) { c[0] = b.x[0]; c[1] = b.x[1]; c[2] = b.x[2]; }
template <class T> IBVH_HD void center(const BBox<T>& b, T c[3]) {
    c[0] = T(0.5) * (b.lo[0] + b.up[0]);
    c[1] = T(0.5) * (b.lo[1] + b.up[1]);
    c[2] = T(0.5) * (b.lo[2] + b.up[2]);
}

// ---- merge.jl -----------------------------------------------------------------------------------
template <class T> IBVH_HD BBox<T> to_box(const BSphere<T>& a) {          // merge.jl:47-51
    BBox<T> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = a.x[k] - a.r; o.up[k] = a.x[k] + a.r; }
    return o;
}
template <class T> IBVH_HD BBox<T> merge(const BBox<T>& a, const BBox<T>& b) {   // merge.jl:30-43
    BBox<T> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = minimum2(a.lo[k], b.lo[k]); o.up[k] = maximum2(a.up[k], b.up[k]); }
    return o;
}
template <class T> IBVH_HD BBox<T> merge_to_box(const BSphere<T>& a, const BSphere<T>& b) {  // merge.jl:58-81
    T length = ibvh_sqrt(dist3sq(a.x, b.x));
    if (length + a.r <= b.r) return to_box(b);
    if (length + b.r <= a.r) return to_box(a);
    BBox<T> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.lo[k] = minimum2(a.x[k] - a.r, b.x[k] - b.r);
        o.up[k] = maximum2(a.x[k] + a.r, b.x[k] + b.r);
    }
    return o;
}
template <class T> IBVH_HD BSphere<T> merge(const BSphere<T>& a, const BSphere<T>& b) {     // merge.jl:2-26
    T length = ibvh_sqrt(dist3sq(a.x, b.x));
    if (length + a.r <= b.r) return b;
    if (length + b.r <= a.r) return a;
    T frac = T(0.5) * ((b.r - a.r) / length + T(1));
    BSphere<T> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) o.x[k] = a.x[k] + frac * (b.x[k] - a.x[k]);
    o.r = T(0.5) * ((length + a.r) + b.r);
    return o;
}

// ---- leaf volumes from triangles (bsphere.jl:43-112, bbox.jl:59-70; utils.jl:178-181) ---------------
template <class T> IBVH_HD T minimum3(T a, T b, T c) { return a < b ? minimum2(a, c) : minimum2(b, c); }
template <class T> IBVH_HD T maximum3(T a, T b, T c) { return a > b ? maximum2(a, c) : maximum2(b, c); }
template <class T> struct FloatEps;
template <> struct FloatEps<float> { static IBVH_HD float eps() { return 1.1920929e-07f; } };
template <> struct FloatEps<double> { static IBVH_HD double eps() { return 2.220446049250313e-16; } };
template <class T> IBVH_HD BSphere<T> sphere_from_triangle(const T* a, const T* b, const T* c) {
    T abab = ((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1])) + (b[2] - a[2]) * (b[2] - a[2]);
    T abac = ((b[0] - a[0]) * (c[0] - a[0]) + (b[1] - a[1]) * (c[1] - a[1])) + (b[2] - a[2]) * (c[2] - a[2]);
    T acac = ((c[0] - a[0]) * (c[0] - a[0]) + (c[1] - a[1]) * (c[1] - a[1])) + (c[2] - a[2]) * (c[2] - a[2]);
    T d = T(2) * (abab * acac - abac * abac);
    BSphere<T> o;
    if (ibvh_abs(d) <= FloatEps<T>::eps()) {
        T up[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { T lo = minimum3(a[k], b[k], c[k]); up[k] = maximum3(a[k], b[k], c[k]); o.x[k] = T(0.5) * (lo + up[k]); }
        o.r = ibvh_sqrt(dist3sq(o.x, up));
        return o;
    }
    T s = (abab * acac - acac * abac) / d;
    T t = (acac * abab - abab * abac) / d;
    if (s <= T(0)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) o.x[k] = T(0.5) * (a[k] + c[k]);
        o.r = ibvh_sqrt(dist3sq(o.x, a));
    } else if (t <= T(0)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) o.x[k] = T(0.5) * (a[k] + b[k]);
        o.r = ibvh_sqrt(dist3sq(o.x, a));
    } else if (s + t >= T(1)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) o.x[k] = T(0.5) * (b[k] + c[k]);
        o.r = ibvh_sqrt(dist3sq(o.x, b));
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) o.x[k] = (a[k] + s * (b[k] - a[k])) + t * (c[k] - a[k]);
        o.r = ibvh_sqrt(dist3sq(o.x, a));
    }
    return o;
}
template <class T> IBVH_HD BBox<T> box_from_triangle(const T* a, const T* b, const T* c) {
    BBox<T> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = minimum3(a[k], b[k], c[k]); o.up[k] = maximum3(a[k], b[k], c[k]); }
    return o;
}

// NodeType(leaf.volume) / NodeType(l.volume, r.volume) — build.jl:438-453.
// The node float type TN may differ from the leaf float type TL (the reference's default call builds BBox{Float32} nodes
// over Float64 leaves, README.md:38-46): the converting constructors (merge.jl:47-81) do their arithmetic in the LEAF
// type and convert the resulting tuples to TN (round to nearest), which is what the casts below restate.
template <class TN, class TL> IBVH_HD BBox<TN> to_box_as(const BSphere<TL>& a) {
    BBox<TN> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = (TN)(a.x[k] - a.r); o.up[k] = (TN)(a.x[k] + a.r); }
    return o;
}
template <class TN, class TL> IBVH_HD BBox<TN> to_box_as(const BBox<TL>& a) {
    BBox<TN> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = (TN)a.lo[k]; o.up[k] = (TN)a.up[k]; }
    return o;
}
template <class TN, class TL> IBVH_HD BBox<TN> merge_to_box_as(const BSphere<TL>& a, const BSphere<TL>& b) {
    TL length = ibvh_sqrt(dist3sq(a.x, b.x));
    if (length + a.r <= b.r) return to_box_as<TN>(b);
    if (length + b.r <= a.r) return to_box_as<TN>(a);
    BBox<TN> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.lo[k] = (TN)minimum2(a.x[k] - a.r, b.x[k] - b.r);
        o.up[k] = (TN)maximum2(a.x[k] + a.r, b.x[k] + b.r);
    }
    return o;
}
template <class TN, class TL> IBVH_HD BBox<TN> merge_box_as(const BBox<TL>& a, const BBox<TL>& b) {
    BBox<TN> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = (TN)minimum2(a.lo[k], b.lo[k]); o.up[k] = (TN)maximum2(a.up[k], b.up[k]); }
    return o;
}
template <class TN, class TL> IBVH_HD BSphere<TN> to_sphere_as(const BSphere<TL>& a) {
    BSphere<TN> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) o.x[k] = (TN)a.x[k];
    o.r = (TN)a.r;
    return o;
}
// BSphere{TN}(a::BSphere{TL}, b::BSphere{TL}): the literals are of type TN, the operands of type TL — Julia promotes, so
// the arithmetic runs in the wider of the two (only TL = double, TN = float is instantiated: the wider type is TL)
template <class TN, class TL> IBVH_HD BSphere<TN> merge_sphere_as(const BSphere<TL>& a, const BSphere<TL>& b) {
    TL length = ibvh_sqrt(dist3sq(a.x, b.x));
    if (length + a.r <= b.r) return to_sphere_as<TN>(b);
    if (length + b.r <= a.r) return to_sphere_as<TN>(a);
    TL frac = TL(0.5) * ((b.r - a.r) / length + TL(1));
    BSphere<TN> o;
#pragma unroll
    for (int k = 0; k < 3; ++k) o.x[k] = (TN)(a.x[k] + frac * (b.x[k] - a.x[k]));
    o.r = (TN)(TL(0.5) * ((length + a.r) + b.r));
    return o;
}
template <class N> struct NodeOps;
template <class T> struct NodeOps<BBox<T>> {
    static IBVH_HD BBox<T> convert(const BBox<T>& v) { return v; }
    static IBVH_HD BBox<T> convert(const BSphere<T>& v) { return to_box(v); }
    static IBVH_HD BBox<T> merge_leaves(const BBox<T>& a, const BBox<T>& b) { return merge(a, b); }
    static IBVH_HD BBox<T> merge_leaves(const BSphere<T>& a, const BSphere<T>& b) { return merge_to_box(a, b); }
    // leaf float type != node float type
    template <class TL> static IBVH_HD BBox<T> convert(const BBox<TL>& v) { return to_box_as<T>(v); }
    template <class TL> static IBVH_HD BBox<T> convert(const BSphere<TL>& v) { return to_box_as<T>(v); }
    template <class TL> static IBVH_HD BBox<T> merge_leaves(const BBox<TL>& a, const BBox<TL>& b) { return merge_box_as<T>(a, b); }
    template <class TL> static IBVH_HD BBox<T> merge_leaves(const BSphere<TL>& a, const BSphere<TL>& b) { return merge_to_box_as<T>(a, b); }
};
template <class T> struct NodeOps<BSphere<T>> {
    static IBVH_HD BSphere<T> convert(const BSphere<T>& v) { return v; }
    static IBVH_HD BSphere<T> merge_leaves(const BSphere<T>& a, const BSphere<T>& b) { return merge(a, b); }
    template <class TL> static IBVH_HD BSphere<T> convert(const BSphere<TL>& v) { return to_sphere_as<T>(v); }
    template <class TL> static IBVH_HD BSphere<T> merge_leaves(const BSphere<TL>& a, const BSphere<TL>& b) { return merge_sphere_as<T>(a, b); }
};

// ---- iscontact.jl:2-28 -----------------------------------------------------------------------------
template <class T> IBVH_HD bool iscontact(const BSphere<T>& a, const BSphere<T>& b) {
    return dist3sq(a.x, b.x) <= (a.r + b.r) * (a.r + b.r);
}
// Device form of the box test: the six closed comparisons of the reference as ONE predicate chain
// (setp.and), returning `bit` or 0. nvcc turns the C++ form into 6 FSETP + 6 SEL per test; the chain is
// 6 FSETP + 1 SEL, and the box tests are what the issue-bound traversal kernels spend their slots on.
// Ordered comparisons: any NaN makes the test false, exactly like `>=` / `<=`.
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t box_contact_bit(const BBox<float>& a, const BBox<float>& b, uint32_t bit) {
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.f32 p, %2, %3;\n\t"
        "setp.le.and.f32 p, %4, %5, p;\n\t"
        "setp.ge.and.f32 p, %6, %7, p;\n\t"
        "setp.le.and.f32 p, %8, %9, p;\n\t"
        "setp.ge.and.f32 p, %10, %11, p;\n\t"
        "setp.le.and.f32 p, %12, %13, p;\n\t"
        "selp.u32 %0, %1, 0, p;\n\t}"
        : "=r"(r) : "r"(bit), "f"(a.up[0]), "f"(b.lo[0]), "f"(a.lo[0]), "f"(b.up[0]), "f"(a.up[1]), "f"(b.lo[1]), "f"(a.lo[1]), "f"(b.up[1]),
          "f"(a.up[2]), "f"(b.lo[2]), "f"(a.lo[2]), "f"(b.up[2]));
    return r;
}
__device__ __forceinline__ uint32_t box_contact_bit(const BBox<double>& a, const BBox<double>& b, uint32_t bit) {
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.f64 p, %2, %3;\n\t"
        "setp.le.and.f64 p, %4, %5, p;\n\t"
        "setp.ge.and.f64 p, %6, %7, p;\n\t"
        "setp.le.and.f64 p, %8, %9, p;\n\t"
        "setp.ge.and.f64 p, %10, %11, p;\n\t"
        "setp.le.and.f64 p, %12, %13, p;\n\t"
        "selp.u32 %0, %1, 0, p;\n\t}"
        : "=r"(r) : "r"(bit), "d"(a.up[0]), "d"(b.lo[0]), "d"(a.lo[0]), "d"(b.up[0]), "d"(a.up[1]), "d"(b.lo[1]), "d"(a.lo[1]), "d"(b.up[1]),
          "d"(a.up[2]), "d"(b.lo[2]), "d"(a.lo[2]), "d"(b.up[2]));
    return r;
}
#endif
template <class T> IBVH_HD bool iscontact(const BBox<T>& a, const BBox<T>& b) {
#ifdef __CUDA_ARCH__
    return box_contact_bit(a, b, 1u) != 0u;
#else
    // same six closed comparisons as the reference, evaluated without short-circuit branches
    return ((a.up[0] >= b.lo[0]) & (a.lo[0] <= b.up[0])) &
           ((a.up[1] >= b.lo[1]) & (a.lo[1] <= b.up[1])) &
           ((a.up[2] >= b.lo[2]) & (a.lo[2] <= b.up[2]));
#endif
}

// ---- isintersection.jl:1-65 --------------------------------------------------------------------------
template <class T> IBVH_HD bool isintersection(const BBox<T>& b, const T p[3], const T d[3]) {
    T inv0 = T(1) / d[0], inv1 = T(1) / d[1], inv2 = T(1) / d[2];
    T t1 = (b.lo[0] - p[0]) * inv0;
    T t2 = (b.up[0] - p[0]) * inv0;
    T tmin = minimum2(t1, t2);
    T tmax = maximum2(t1, t2);
    t1 = (b.lo[1] - p[1]) * inv1;
    t2 = (b.up[1] - p[1]) * inv1;
    tmin = maximum2(tmin, minimum2(t1, t2));
    tmax = minimum2(tmax, maximum2(t1, t2));
    t1 = (b.lo[2] - p[2]) * inv2;
    t2 = (b.up[2] - p[2]) * inv2;
    tmin = maximum2(tmin, minimum2(t1, t2));
    tmax = minimum2(tmax, maximum2(t1, t2));
    return (tmin <= tmax) && (tmax >= T(0));
}
// the same slab test with inv = 1 / d computed once per ray by the caller (the same three IEEE divisions, hoisted)
template <class T> IBVH_HD bool isintersection_inv(const BBox<T>& b, const T p[3], const T inv[3]) {
    T t1 = (b.lo[0] - p[0]) * inv[0];
    T t2 = (b.up[0] - p[0]) * inv[0];
    T tmin = minimum2(t1, t2);
    T tmax = maximum2(t1, t2);
    t1 = (b.lo[1] - p[1]) * inv[1];
    t2 = (b.up[1] - p[1]) * inv[1];
    tmin = maximum2(tmin, minimum2(t1, t2));
    tmax = minimum2(tmax, maximum2(t1, t2));
    t1 = (b.lo[2] - p[2]) * inv[2];
    t2 = (b.up[2] - p[2]) * inv[2];
    tmin = maximum2(tmin, minimum2(t1, t2));
    tmax = minimum2(tmax, maximum2(t1, t2));
    return (tmin <= tmax) && (tmax >= T(0));
}
// node test of the ray kernels: BBox nodes take the hoisted reciprocals, BSphere nodes the direction itself
template <class T> IBVH_HD bool ray_hits_node(const BBox<T>& b, const T p[3], const T d[3], const T inv[3]) { (void)d; return isintersection_inv(b, p, inv); }
template <class T> IBVH_HD bool isintersection(const BSphere<T>& s, const T p[3], const T d[3]);
template <class T> IBVH_HD bool ray_hits_node(const BSphere<T>& s, const T p[3], const T d[3], const T inv[3]) { (void)inv; return isintersection(s, p, d); }
template <class T> IBVH_HD bool isintersection(const BSphere<T>& s, const T p[3], const T d[3]) {
    T a = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
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

// ---- implicit tree (implicit_tree.jl) — 0-based helpers used by the kernels ---------------------------
// Levels are 1-based (root = level 1, leaves = level `levels`). Within a level, nodes are numbered
// 0-based: implicit index = 2^(level-1) + i. Julia precedence: a - b >> c == a - (b >> c).
struct TreeInfo {
    int32_t levels;
    int64_t n;                 // real leaves
    int64_t virtual_leaves;
    int64_t level_start[34];   // level_start[l] = 0-based memory position of the first node of level l (l = 1..levels-1)
    int64_t level_nreal[34];   // real nodes on level l (l = 1..levels)
    int64_t skips[34];         // skips[l] as in compute_skips! (1-based level), implicit_tree.jl:100-113
};

IBVH_HD int ilog2_floor_u64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return 63 - __clzll((long long)x);
#else
    return 63 - __builtin_clzll(x);
#endif
}
IBVH_HD int64_t shr64(int64_t v, int64_t s) { return s >= 63 ? 0 : (v >> s); }

inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

inline int make_tree(int64_t n, ibvh_tree_t* t) {        // implicit_tree.jl:77-90
    if (n < 1) return IBVH_ERR_DOMAIN;
    int fl = 63 - __builtin_clzll((uint64_t)n);
    int cl = ((n & (n - 1)) == 0) ? fl : fl + 1;
    t->levels = cl + 1;
    t->real_leaves = n;
    int64_t lv = (int64_t(1) << (t->levels - 1)) - n;
    t->virtual_leaves = lv;
    t->virtual_nodes = 2 * lv - popc64((uint64_t)lv);
    t->real_nodes = 2 * n - 1 + popc64((uint64_t)lv);
    return IBVH_OK;
}
inline int64_t skip_of_level(const ibvh_tree_t& t, int64_t level /*1-based*/) {   // implicit_tree.jl:107-108
    int64_t v = shr64(t.virtual_leaves, t.levels - (level - 1));
    return 2 * v - popc64((uint64_t)v);
}
inline TreeInfo make_tree_info(const ibvh_tree_t& t) {
    TreeInfo ti{};
    ti.levels = (int32_t)t.levels;
    ti.n = t.real_leaves;
    ti.virtual_leaves = t.virtual_leaves;
    for (int64_t l = 1; l <= t.levels; ++l) {
        ti.skips[l] = skip_of_level(t, l);
        ti.level_nreal[l] = (int64_t(1) << (l - 1)) - shr64(t.virtual_leaves, t.levels - l);
        ti.level_start[l] = (int64_t(1) << (l - 1)) - ti.skips[l] - 1;    // memory_index(2^(l-1)) - 1
    }
    return ti;
}

// ---- 16-byte aligned records of the pyramid traversal (written by the build's sidecar and by the pack kernels) ----
#ifdef __CUDACC__
template <class T> IBVH_D BBox<T> empty_box() {
    BBox<T> b;
    const T inf = T(1) / T(0);
#pragma unroll
    for (int k = 0; k < 3; ++k) { b.lo[k] = inf; b.up[k] = -inf; }
    return b;
}

// Records the pyramid kernels load with 128-bit accesses (the refine / tile kernels are bound by L1 wavefronts,
// not by DRAM: AoS structs read as 8-byte pieces at a 24-byte stride cost 3-7x the wavefronts of aligned
// 16-byte loads). UBox = one query-pyramid box, Packed<V> = one leaf volume, both padded to 16 bytes.
template <class T> struct alignas(16) UBox { BBox<T> b; };
template <class V> struct alignas(16) Packed { V v; };
template <class R> IBVH_D R load16(const R* p) {
    static_assert(sizeof(R) % 16 == 0, "16-byte records");
    alignas(16) R out;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&out);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(R) / 16); ++k) d[k] = __ldg(s + k);
    return out;
}
template <class R> IBVH_D void store16(R* p, const R& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(R) / 16); ++k) d[k] = s[k];
}

#endif  // __CUDACC__

// CUDA error plumbing
#define IBVH_CUDA_TRY(h, expr)                                                     \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) { (h)->set_cuda_error(_e, #expr); return IBVH_ERR_CUDA; } \
    } while (0)

}  // namespace ibvh
