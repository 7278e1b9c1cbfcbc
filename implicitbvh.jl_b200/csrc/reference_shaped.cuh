// reference_shaped.cuh — a deliberately naive, "reference-shaped" GPU build: the proxy for the reference's own
// CUDA.jl backend, which cannot run in this environment (no Julia). It launches what BVH(...) launches through
// KernelAbstractions / AcceleratedKernels, one thread per item, with the reference's data movement:
//   wrap_bounding_volumes (build.jl:328-352, one launch)                                   -> wrap_kernel (morton.cuh)
//   _compute_extrema: TWO mapreduce passes over the leaf structs + two scalar read-backs (morton/utils.jl:24-44)
//   _morton_encode!: every leaf struct rewritten with its code (morton/default.jl:63-82)
//   AK.sort!(leaves, by = morton): comparison MERGE sort that moves the whole 24-byte structs (build.jl:248-253):
//       block sort in shared memory, then log2(n / block) global merge passes (merge-path split per thread)
//   aggregate_oibvh!: one launch per tree level (build.jl:381-523)
// The matching traversal is lvt_thread_kernel (traverse.cuh; IBVH_TRAVERSE_REFERENCE_SHAPED): one thread per
// query leaf, private 32-entry stack, count pass -> scan -> write pass (traverse_single.jl:52-75).
// Results are bit-identical to the product path (stable sort, same merges), so the proxy doubles as a second
// opinion in the parity tests. It is NOT the product path and it is NOT ImplicitBVH.jl: bench.py reports it as
// `reference_shaped` with that label. Default type set only (BSphere{Float32} / Int32 / UInt32 / BBox{Float32}).
#pragma once
#include "common.cuh"
#include "morton.cuh"

namespace ibvh {
namespace refshaped {

using RLeaf = Leaf<BSphere<float>, int32_t, uint32_t>;
using RNode = BBox<float>;
constexpr int kBlockSort = 512;          // leaves per block-level sort
constexpr int kMergePerThread = 8;       // outputs per thread in a global merge pass

// mapreduce(min) / mapreduce(max) over the centres: one pass each, block partials -> one block -> 3 floats
template <bool IS_MAX>
static __global__ void __launch_bounds__(256) extrema_partial_kernel(const RLeaf* __restrict__ leaves, int64_t n, float* __restrict__ partial) {
    float acc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = IS_MAX ? FloatLimits<float>::fmin_() : FloatLimits<float>::fmax_();     // morton/utils.jl:28-29,39-40
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const RLeaf l = leaves[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = IS_MAX ? (acc[k] > l.volume.x[k] ? acc[k] : l.volume.x[k]) : (acc[k] < l.volume.x[k] ? acc[k] : l.volume.x[k]);
    }
    __shared__ float sh[256][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sh[threadIdx.x][k] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float a = sh[threadIdx.x][k], b = sh[threadIdx.x + s][k];
                sh[threadIdx.x][k] = IS_MAX ? (a > b ? a : b) : (a < b ? a : b);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) partial[blockIdx.x * 3 + threadIdx.x] = sh[0][threadIdx.x];
}
template <bool IS_MAX>
__global__ void extrema_final_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ out) {
    if (threadIdx.x < 3) {
        float a = partial[threadIdx.x];
        for (int b = 1; b < blocks; ++b) { const float v = partial[b * 3 + threadIdx.x]; a = IS_MAX ? (a > v ? a : v) : (a < v ? a : v); }
        out[threadIdx.x] = a;
    }
}

// _morton_encode!: bounds (already padded on the host, as the reference does) arrive by value
struct Bounds6 { float mins[3], maxs[3]; };
static __global__ void __launch_bounds__(256) encode_kernel(RLeaf* leaves, int64_t n, Bounds6 b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RLeaf l = leaves[i];
    float c[3];
    center(l.volume, c);
    l.morton = morton_encode_single<uint32_t>(c, b.mins, b.maxs);
    leaves[i] = l;                                   // the whole struct is rewritten (default.jl:76-80)
}

// block-level stable sort of kBlockSort structs by key: bitonic network on (key << 32 | slot), then the structs move
static __global__ void __launch_bounds__(kBlockSort) block_sort_kernel(const RLeaf* __restrict__ in, RLeaf* __restrict__ out, int64_t n) {
    __shared__ unsigned long long comp[kBlockSort];
    __shared__ RLeaf sl[kBlockSort];
    const int64_t base = (int64_t)blockIdx.x * kBlockSort;
    const int t = threadIdx.x;
    const bool valid = base + t < n;
    if (valid) sl[t] = in[base + t];
    comp[t] = valid ? (((unsigned long long)sl[t].morton << 32) | (unsigned)t) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= kBlockSort; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int p = t ^ j;
            if (p > t) {
                const unsigned long long a = comp[t], b = comp[p];
                const bool up = (t & k) == 0;
                if ((a > b) == up) { comp[t] = b; comp[p] = a; }
            }
            __syncthreads();
        }
    }
    if (valid) out[base + t] = sl[(unsigned)(comp[t] & 0xffffffffu)];
}

// one global merge pass: runs of `width` sorted structs are merged pairwise; each thread produces kMergePerThread
// consecutive outputs after a merge-path binary search (ties take the left run first: stable)
static __global__ void __launch_bounds__(256) merge_pass_kernel(const RLeaf* __restrict__ in, RLeaf* __restrict__ out, int64_t n, int64_t width) {
    const int64_t o0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kMergePerThread;
    if (o0 >= n) return;
    const int64_t pair = o0 / (2 * width);
    const int64_t a0 = pair * 2 * width;
    const int64_t a1 = min(a0 + width, n), b1 = min(a0 + 2 * width, n);
    const int64_t na = a1 - a0, nb = b1 - a1;
    const int64_t d = o0 - a0;                                   // diagonal inside this pair's merge
    // merge path: smallest i with A[i] > B[d - i - 1] (A first on ties)
    int64_t lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
    while (lo < hi) {
        const int64_t i = (lo + hi) >> 1;
        if (in[a0 + i].morton <= in[a1 + (d - i - 1)].morton) lo = i + 1; else hi = i;
    }
    int64_t i = lo, j = d - lo;
    for (int k = 0; k < kMergePerThread; ++k) {
        const int64_t o = o0 + k;
        if (o >= b1) break;
        bool take_a;
        if (i >= na) take_a = false;
        else if (j >= nb) take_a = true;
        else take_a = in[a0 + i].morton <= in[a1 + j].morton;
        out[o] = take_a ? in[a0 + i] : in[a1 + j];                // whole 24-byte structs move
        if (take_a) ++i; else ++j;
    }
}

// aggregate_last_level!: parent i of leaves (2i, 2i+1), one thread per node (build.jl:427-457)
static __global__ void __launch_bounds__(256) aggregate_last_level_kernel(const RLeaf* __restrict__ leaves, RNode* __restrict__ nodes, int64_t n, int64_t start_pos, int64_t num_nodes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    const int64_t l = 2 * i, r = 2 * i + 1;
    nodes[start_pos + i] = r >= n ? NodeOps<RNode>::convert(leaves[l].volume) : NodeOps<RNode>::merge_leaves(leaves[l].volume, leaves[r].volume);
}
// aggregate_level!: parent i of nodes (2i, 2i+1) of the level below (build.jl:503-523)
static __global__ void __launch_bounds__(256) aggregate_level_kernel(RNode* nodes, int64_t start_pos, int64_t num_nodes, int64_t start_next, int64_t num_next) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    const int64_t l = 2 * i, r = 2 * i + 1;
    nodes[start_pos + i] = r >= num_next ? nodes[start_next + l] : merge(nodes[start_next + l], nodes[start_next + r]);
}

}  // namespace refshaped
}  // namespace ibvh
