// sort_bench.cu — development harness (not the product): times the onesweep radix sort of radix_sort.cuh in several
// tile shapes against cub::DeviceRadixSort on the same keys, and checks every variant against CUB's (stable) result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -o sort_bench tools/sort_bench.cu
//   ./sort_bench [n] [dist]     dist: 0 uniform 30-bit, 1 clustered (few distinct top digits), 2 many ties (15 distinct bits)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cub/device/device_radix_sort.cuh>

#include "../implicitbvh.jl_b200/csrc/radix_sort.cuh"

using namespace ibvh;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __host__ inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
template <class K> __global__ void gen_keys(K* keys, int64_t n, int dist) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r = mix64((uint64_t)i * 2654435761ull + 12345);
    constexpr int kb = MortonTraits<K>::key_bits;
    uint64_t k = r & ((1ull << kb) - 1);
    if (dist == 1) k = (k & ((1ull << (kb - 8)) - 1)) | ((uint64_t)(r >> 60) << (kb - 4));      // 16 distinct top nibbles
    if (dist == 2) k &= 0x7FFFull;                                                          // many ties
    keys[i] = (K)k;
}
template <class K> __global__ void hist_kernel(const K* keys, int64_t n, uint32_t* hist) {
    constexpr int P = radix_passes<K>(), RB = radix_bits<K>(), BINS = radix_bins<K>();
    __shared__ uint32_t sh[P][BINS];
    for (int i = threadIdx.x; i < P * BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        K m = keys[i];
#pragma unroll
        for (int p = 0; p < P; ++p) atomicAdd(&sh[p][(uint32_t)(m >> (RB * p)) & (BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * BINS; i += blockDim.x) if ((&sh[0][0])[i]) atomicAdd(&hist[i], (&sh[0][0])[i]);
}

struct NoScope { int operator()(const char*) const { return 0; } };
static const char* g_filter = nullptr;     // run only the variants whose name contains this
static int g_reps = 10;

template <class K, class LB, int THREADS, int ITEMS, int MINB>
void run_variant(const char* name, const K* d_keys, int64_t n, const K* ref_keys, const uint32_t* ref_vals, int reps) {
    if (g_filter && !strstr(name, g_filter)) return;
    constexpr int P = radix_passes<K>(), RB = radix_bits<K>(), BINS = radix_bins<K>();
    const int64_t tiles = (n + THREADS * ITEMS - 1) / (THREADS * ITEMS);
    K *kA, *kB; uint32_t *vA, *vB, *hist, *tickets; LB* lb;
    CK(cudaMalloc(&kA, n * sizeof(K))); CK(cudaMalloc(&kB, n * sizeof(K)));
    CK(cudaMalloc(&vA, n * 4)); CK(cudaMalloc(&vB, n * 4));
    CK(cudaMalloc(&hist, P * BINS * 4)); CK(cudaMalloc(&tickets, 64));
    CK(cudaMalloc(&lb, (size_t)P * tiles * BINS * sizeof(LB)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f, sum = 0;
    K* ko = nullptr; uint32_t* vo = nullptr;
    for (int it = 0; it < reps + 2; ++it) {
        CK(cudaMemcpy(kA, d_keys, n * sizeof(K), cudaMemcpyDeviceToDevice));
        CK(cudaMemset(hist, 0, P * BINS * 4)); CK(cudaMemset(tickets, 0, 64));
        hist_kernel<K><<<148 * 8, 256>>>(kA, n, hist);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        CK(cudaMemsetAsync(lb, 0, (size_t)P * tiles * BINS * sizeof(LB)));
        cudaError_t e = sort_pairs_impl<K, LB, THREADS, ITEMS, MINB>(kA, kB, vA, vB, n, hist, lb, tickets, 0, &ko, &vo, NoScope{});
        CK(e);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    // check against the reference
    std::vector<K> hk(n); std::vector<uint32_t> hv(n);
    CK(cudaMemcpy(hk.data(), ko, n * sizeof(K), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hv.data(), vo, n * 4, cudaMemcpyDeviceToHost));
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) if (hk[i] != ref_keys[i] || hv[i] != ref_vals[i]) { if (bad < 3) printf("   mismatch at %lld: key %llx/%llx val %u/%u\n", (long long)i, (unsigned long long)hk[i], (unsigned long long)ref_keys[i], hv[i], ref_vals[i]); ++bad; }
    const double bytes = (double)n * (sizeof(K) + (double)P * 2 * (sizeof(K) + 4));      // SURVEY.md §8d S2 formula (histogram read + P passes of key + 4-byte value, read and written)
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, onesweep_kernel<K, LB, THREADS, ITEMS, RB, RB, MINB>, THREADS, onesweep_smem_bytes<K, THREADS, ITEMS, BINS>());
    printf("%-28s n=%lld: best %.4f ms avg %.4f ms  %.1f Gpairs/s  %.0f GB/s (S2 bytes)  CTAs/SM %d  %s\n", name, (long long)n, best, sum / reps, n / best / 1e6,
           bytes / best / 1e6, nb, bad ? "MISMATCH" : "ok");
    fflush(stdout);
    cudaFree(kA); cudaFree(kB); cudaFree(vA); cudaFree(vB); cudaFree(hist); cudaFree(tickets); cudaFree(lb);
}

template <class K> void run_all(int64_t n, int dist, int reps) {
    printf("== key bytes %d, n %lld, dist %d, radix bits %d (%d passes)\n", (int)sizeof(K), (long long)n, dist, radix_bits<K>(), radix_passes<K>());
    K* d_keys; CK(cudaMalloc(&d_keys, n * sizeof(K)));
    gen_keys<K><<<(unsigned)((n + 255) / 256), 256>>>(d_keys, n, dist);
    // CUB reference (stable LSD), iota values
    K* ck; uint32_t *cv_in, *cv; CK(cudaMalloc(&ck, n * sizeof(K))); CK(cudaMalloc(&cv_in, n * 4)); CK(cudaMalloc(&cv, n * 4));
    std::vector<uint32_t> iota(n); for (int64_t i = 0; i < n; ++i) iota[i] = (uint32_t)i;
    CK(cudaMemcpy(cv_in, iota.data(), n * 4, cudaMemcpyHostToDevice));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, ck, cv_in, cv, (int)n, 0, MortonTraits<K>::key_bits);
    void* tmp; CK(cudaMalloc(&tmp, tmp_bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < reps + 2; ++it) {
        CK(cudaEventRecord(e0));
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_keys, ck, cv_in, cv, (int)n, 0, MortonTraits<K>::key_bits);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2 && ms < best) best = ms;
    }
    printf("%-28s n=%lld: best %.4f ms  %.1f Gpairs/s (includes its own histogram pass)\n", "cub::DeviceRadixSort", (long long)n, best, n / best / 1e6);
    std::vector<K> rk(n); std::vector<uint32_t> rv(n);
    CK(cudaMemcpy(rk.data(), ck, n * sizeof(K), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rv.data(), cv, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(ck); cudaFree(cv_in); cudaFree(cv); cudaFree(tmp);
#define V(T, I, M) run_variant<K, uint32_t, T, I, M>(#T "x" #I " minb" #M, d_keys, n, rk.data(), rv.data(), reps)
    V(512, 8, 3); V(512, 8, 2); V(512, 8, 4); V(256, 16, 4); V(256, 16, 3); V(256, 8, 6); V(256, 8, 4); V(384, 8, 4); V(384, 12, 3); V(512, 16, 2); V(512, 12, 2); V(1024, 8, 1); V(1024, 4, 2);
    V(256, 12, 4); V(512, 4, 4);
#undef V
    run_variant<K, unsigned long long, 512, 8, 3>("512x8 minb3 wide-lookback", d_keys, n, rk.data(), rv.data(), reps);
    cudaFree(d_keys);
}

int main(int argc, char** argv) {
    int64_t n = argc > 1 ? atoll(argv[1]) : 10000000;
    int dist = argc > 2 ? atoi(argv[2]) : 0;
    int kb = argc > 3 ? atoi(argv[3]) : 4;
    if (argc > 4) g_filter = argv[4];
    if (argc > 5) g_reps = atoi(argv[5]);
    int reps = g_reps;
    if (kb == 4) run_all<uint32_t>(n, dist, reps);
    else if (kb == 8) run_all<uint64_t>(n, dist, reps);
    else run_all<uint16_t>(n, dist, reps);
    return 0;
}
