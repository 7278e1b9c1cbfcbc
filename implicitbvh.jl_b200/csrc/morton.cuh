// morton.cuh — kernels for scene bounds + Morton encoding (+ radix digit histograms).
// Replaces: _compute_extrema / bounding_volumes_extrema (src/morton/utils.jl:1-72), morton_encode!
// and morton_encode_single / morton_split3 (src/morton/default.jl:43-157) and, on the wrap path,
// wrap_bounding_volumes (src/build.jl:328-352).
//
// B200 design: two streaming passes over the leaf array (the encode pass cannot start before the
// global bounds are known). Pass 1 reduces the 3-D min / max of the centres in registers ->
// warp shuffles -> one atomicMin/atomicMax per block on order-preserving integer keys. Pass 2
// recomputes the padded bounds in-kernel (bit-exact with the host formula of the reference),
// encodes, writes the compact key array that the radix sort consumes, accumulates the per-digit
// histograms of ALL radix passes in shared memory (so the sort never re-reads the keys for them)
// and, on the in-place path, streams a copy of the leaves out for the later gather.
#pragma once
#include "common.cuh"

namespace ibvh {

// ---- order-preserving float <-> unsigned maps (for atomicMin / atomicMax) -----------------------
IBVH_HD uint32_t ord_encode(float f) {
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(f);
#else
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
IBVH_HD float ord_decode(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(b);
#else
    memcpy(&f, &b, 4);
#endif
    return f;
}
IBVH_HD unsigned long long ord_encode(double f) {
    unsigned long long b;
#ifdef __CUDA_ARCH__
    b = (unsigned long long)__double_as_longlong(f);
#else
    memcpy(&b, &f, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
IBVH_HD double ord_decode(unsigned long long u) {
    unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double f;
#ifdef __CUDA_ARCH__
    f = __longlong_as_double((long long)b);
#else
    memcpy(&f, &b, 8);
#endif
    return f;
}
template <class T> struct OrdOf;
template <> struct OrdOf<float> { using type = uint32_t; };
template <> struct OrdOf<double> { using type = unsigned long long; };

template <class T> struct FloatLimits;
template <> struct FloatLimits<float> {
    static IBVH_HD float fmax_() { return 3.402823466e+38f; }      // floatmax(Float32)
    static IBVH_HD float fmin_() { return 1.175494351e-38f; }      // floatmin(Float32): smallest positive normal
    static IBVH_HD float rel_prec() { return 1e-5f; }              // default.jl:180
};
template <> struct FloatLimits<double> {
    static IBVH_HD double fmax_() { return 1.7976931348623157e+308; }
    static IBVH_HD double fmin_() { return 2.2250738585072014e-308; }
    static IBVH_HD double rel_prec() { return 1e-14; }             // default.jl:181
};

// ---- morton_split3, default.jl:118-157 ---------------------------------------------------------------
IBVH_HD uint16_t morton_split3(uint16_t v) {
    uint32_t s = v & 0x001fu;
    s = (s | (s << 8)) & 0x100fu;
    s = (s | (s << 4)) & 0x10c3u;
    s = (s | (s << 2)) & 0x1249u;
    return (uint16_t)s;
}
IBVH_HD uint32_t morton_split3(uint32_t v) {
    uint32_t s = v & 0x000003ffu;
    s = (s | (s << 16)) & 0x030000ffu;
    s = (s | (s << 8)) & 0x0300f00fu;
    s = (s | (s << 4)) & 0x030c30c3u;
    s = (s | (s << 2)) & 0x09249249u;
    return s;
}
IBVH_HD uint64_t morton_split3(uint64_t v) {
    uint64_t s = v & 0x00000000001fffffull;
    s = (s | (s << 32)) & 0x001f00000000ffffull;
    s = (s | (s << 16)) & 0x001f0000ff0000ffull;
    s = (s | (s << 8)) & 0x100f00f00f00f00full;
    s = (s | (s << 4)) & 0x10c30c30c30c30c3ull;
    s = (s | (s << 2)) & 0x1249249249249249ull;
    return s;
}
template <class M> struct MortonTraits;
template <> struct MortonTraits<uint16_t> { static constexpr int scaling = 1 << 5;  static constexpr int key_bits = 15; };
template <> struct MortonTraits<uint32_t> { static constexpr int scaling = 1 << 10; static constexpr int key_bits = 30; };
template <> struct MortonTraits<uint64_t> { static constexpr int scaling = 1 << 21; static constexpr int key_bits = 63; };

// morton_encode_single, default.jl:91-108. unsafe_trunc == cvt.rzi for in-range values.
template <class M, class T> IBVH_HD M morton_encode_single(const T c[3], const T mins[3], const T maxs[3]) {
    const T scaling = T(MortonTraits<M>::scaling);
    T s1 = (c[0] - mins[0]) / (maxs[0] - mins[0]);
    T s2 = (c[1] - mins[1]) / (maxs[1] - mins[1]);
    T s3 = (c[2] - mins[2]) / (maxs[2] - mins[2]);
    M i1 = (M)(s1 * scaling), i2 = (M)(s2 * scaling), i3 = (M)(s3 * scaling);
    return (M)((M)(morton_split3(i1) << 2) | (M)(morton_split3(i2) << 1) | morton_split3(i3));
}

// bounding_volumes_extrema padding, morton/utils.jl:63-69: (m - rp*abs(m)) - floatmin, left to right.
template <class T> IBVH_HD void pad_extrema(T mins[3], T maxs[3]) {
    const T rp = FloatLimits<T>::rel_prec();
    const T fm = FloatLimits<T>::fmin_();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mins[k] = (mins[k] - rp * ibvh_abs(mins[k])) - fm;
        maxs[k] = (maxs[k] + rp * ibvh_abs(maxs[k])) + fm;
    }
}

// Radix digit width per key type (radix_sort.cuh; the histograms of encode_kernel follow it). 8 bits for every code
// width: 30-bit codes in three 10-bit passes and 63-bit codes in seven 9-bit passes were measured slower (radix_sort.cuh,
// tile-shape notes) and stay available as build-time variants for tools/sort_bench.cu.
#ifndef IBVH_SORT_RB32
#define IBVH_SORT_RB32 8
#endif
#ifndef IBVH_SORT_RB64
#define IBVH_SORT_RB64 8
#endif
constexpr int kMaxRadixPasses = 8;
template <class M> constexpr int radix_bits() { return sizeof(M) == 4 ? IBVH_SORT_RB32 : (sizeof(M) == 8 ? IBVH_SORT_RB64 : 8); }
template <class M> constexpr int radix_bins() { return 1 << radix_bits<M>(); }
template <class M> constexpr int radix_passes() { return (MortonTraits<M>::key_bits + radix_bits<M>() - 1) / radix_bits<M>(); }
constexpr int kMaxRadixBins = 1024;

// ---- init: bounds seeds (floatmax / floatmin — morton/utils.jl:28-29,39-40), zero histograms ------------
template <class T>
__global__ void init_build_kernel(typename OrdOf<T>::type* bounds, uint32_t* hist, int hist_words, uint32_t* tickets, int ntickets) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3) bounds[t] = ord_encode(FloatLimits<T>::fmax_());
    else if (t < 6) bounds[t] = ord_encode(FloatLimits<T>::fmin_());
    for (int i = t; i < hist_words; i += gridDim.x * blockDim.x) hist[i] = 0;
    if (t < ntickets) tickets[t] = 0;
}

// ---- pass 1: min / max of the centres --------------------------------------------------------------------
// SRC is either a raw volume (wrap path) or a wrapped leaf. Records are fetched with the widest loads the array's
// alignment allows (`vec` = 16 / 8 / 4 bytes, chosen by the host from the pointer: a struct of floats only promises
// 4-byte alignment) and kStreamUnroll records per thread are in flight at once — with one scalar-load record per
// thread the two streaming passes ran at 72 % / 53 % of the copy bandwidth (round 1).
constexpr int kStreamUnroll = 4;

template <class V> IBVH_D V fetch_volume(const V* p, int vec) { return load_volume(p, vec); }
template <class V, class I, class M> IBVH_D V fetch_volume(const Leaf<V, I, M>* p, int) {
    Words<Leaf<V, I, M>> wv = load_words(p);
    return words_volume<Leaf<V, I, M>>(wv);
}

template <class SRC, class T>
__global__ void __launch_bounds__(256) bounds_kernel(const SRC* __restrict__ src, int64_t n, typename OrdOf<T>::type* bounds, int vec) {
    T mn[3] = {FloatLimits<T>::fmax_(), FloatLimits<T>::fmax_(), FloatLimits<T>::fmax_()};
    T mx[3] = {FloatLimits<T>::fmin_(), FloatLimits<T>::fmin_(), FloatLimits<T>::fmin_()};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kStreamUnroll - 1) * stride < n; i += kStreamUnroll * stride) {
        decltype(fetch_volume(src, vec)) v[kStreamUnroll];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) v[u] = fetch_volume(src + i + u * stride, vec);
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            T c[3];
            center(v[u], c);
#pragma unroll
            for (int k = 0; k < 3; ++k) { mn[k] = mn[k] < c[k] ? mn[k] : c[k]; mx[k] = mx[k] > c[k] ? mx[k] : c[k]; }
        }
    }
    for (; i < n; i += stride) {
        T c[3];
        center(fetch_volume(src + i, vec), c);
#pragma unroll
        for (int k = 0; k < 3; ++k) { mn[k] = mn[k] < c[k] ? mn[k] : c[k]; mx[k] = mx[k] > c[k] ? mx[k] : c[k]; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T o = __shfl_xor_sync(0xffffffffu, mn[k], off); mn[k] = mn[k] < o ? mn[k] : o;
            T p = __shfl_xor_sync(0xffffffffu, mx[k], off); mx[k] = mx[k] > p ? mx[k] : p;
        }
    }
    __shared__ T smn[8][3], smx[8][3];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { smn[w][k] = mn[k]; smx[w][k] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int k = threadIdx.x;
        T a = smn[0][k], b = smx[0][k];
        for (int j = 1; j < (int)(blockDim.x >> 5); ++j) { a = a < smn[j][k] ? a : smn[j][k]; b = b > smx[j][k] ? b : smx[j][k]; }
        atomicMin(&bounds[k], ord_encode(a));
        atomicMax(&bounds[3 + k], ord_encode(b));
    }
}

// ---- pass 2: encode (+ histograms, + leaf copy) ----------------------------------------------------------
// bounds_in: ordered keys of the raw extrema (compute_extrema) or nullptr when user bounds are given
// in `user_bounds` (6 values, unpadded — SURVEY.md §8c quirk 2).
// used_bounds: 6 T's written by block 0 for read-back.
// Digit histograms of all radix passes: one shared-memory atomic per key and pass into a histogram PRIVATE to the
// warp (IBVH_ENCODE_HIST == 2; 8 warps never collide on a counter) or shared by the block (== 1).
#ifndef IBVH_ENCODE_HIST
#define IBVH_ENCODE_HIST 1
#endif
template <class SRC, class L, bool COPY>
__global__ void __launch_bounds__(256) encode_kernel(const SRC* __restrict__ src, int64_t n,
                                                    const typename OrdOf<typename L::value_type>::type* __restrict__ bounds_in,
                                                    const typename L::value_type* __restrict__ user_bounds,
                                                    typename L::value_type* used_bounds,
                                                    typename L::mor_t* __restrict__ keys,
                                                    L* __restrict__ copy_out, uint32_t* __restrict__ hist, int vec) {
    using T = typename L::value_type;
    using M = typename L::mor_t;
    constexpr int P = radix_passes<M>();
    constexpr int RB = radix_bits<M>(), BINS = radix_bins<M>();
    constexpr int HW = IBVH_ENCODE_HIST == 2 ? 8 : 1;                      // private histograms per block
    __shared__ uint32_t sh[HW][P][BINS];
    for (int i = threadIdx.x; i < HW * P * BINS; i += blockDim.x) (&sh[0][0][0])[i] = 0;
    uint32_t (*myh)[BINS] = sh[HW == 1 ? 0 : (threadIdx.x >> 5)];
    T mins[3], maxs[3];
    if (bounds_in) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { mins[k] = ord_decode(bounds_in[k]); maxs[k] = ord_decode(bounds_in[3 + k]); }
        pad_extrema(mins, maxs);
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { mins[k] = user_bounds[k]; maxs[k] = user_bounds[3 + k]; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { used_bounds[k] = mins[k]; used_bounds[3 + k] = maxs[k]; }
    }
    __syncthreads();
    auto one = [&](int64_t i, const T (&c)[3]) {
        M m = morton_encode_single<M>(c, mins, maxs);
        keys[i] = m;
#if IBVH_ENCODE_HIST
#pragma unroll
        for (int p = 0; p < P; ++p) atomicAdd(&myh[p][(uint32_t)(m >> (p * RB)) & (BINS - 1)], 1u);
#endif
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (COPY) {
        // move the leaf as raw words so every byte (padding included) survives the copy
        for (; i < n; i += stride) {
            T c[3];
            Words<L> wv = load_words(reinterpret_cast<const L*>(src) + i);
            store_words(copy_out + i, wv);
            center(words_volume<L>(wv), c);
            one(i, c);
        }
    } else {
        for (; i + (kStreamUnroll - 1) * stride < n; i += kStreamUnroll * stride) {
            decltype(fetch_volume(src, vec)) v[kStreamUnroll];
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) v[u] = fetch_volume(src + i + u * stride, vec);
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                T c[3];
                center(v[u], c);
                one(i + u * stride, c);
            }
        }
        for (; i < n; i += stride) {
            T c[3];
            center(fetch_volume(src + i, vec), c);
            one(i, c);
        }
    }
    __syncthreads();
    for (int i2 = threadIdx.x; i2 < P * BINS; i2 += blockDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (int hw = 0; hw < HW; ++hw) v += (&sh[hw][0][0])[i2];
        if (v) atomicAdd(&hist[i2], v);
    }
}

// Stand-alone morton_encode! (stage-level entry point): rewrite .morton in place from a key array.
template <class L>
__global__ void __launch_bounds__(256) scatter_morton_kernel(L* leaves, int64_t n, const typename L::mor_t* __restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) leaves[i].morton = keys[i];     // touches only the morton field
}

// BSphere{T}(p1, p2, p3) / BBox{T}(p1, p2, p3) for every triangle (bsphere.jl:43-112, bbox.jl:59-70):
// tri = T[n][3][3]. Nine coalesced-ish loads per thread, one volume store.
template <class V>
__global__ void __launch_bounds__(256) triangles_kernel(const typename V::value_type* __restrict__ tri, int64_t n, V* __restrict__ out) {
    using T = typename V::value_type;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T p[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) p[k] = tri[9 * i + k];
    if constexpr (V::kind == IBVH_BSPHERE) out[i] = sphere_from_triangle<T>(p, p + 3, p + 6);
    else out[i] = box_from_triangle<T>(p, p + 3, p + 6);
}

// wrap_bounding_volumes, build.jl:340-350
template <class L>
__global__ void __launch_bounds__(256) wrap_kernel(const typename L::vol_t* __restrict__ vols, int64_t n, L* __restrict__ leaves) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        Words<L> wv = zero_words<L>();            // padding bytes are zero, deterministically
        words_set_volume<L>(wv, load_volume(vols + i, 4));
        words_set_index<L>(wv, (typename L::idx_t)(i + 1));
        store_words(leaves + i, wv);
    }
}

}  // namespace ibvh
