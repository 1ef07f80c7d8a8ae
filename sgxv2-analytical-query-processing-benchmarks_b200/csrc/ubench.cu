// ubench.cu — on-chip micro-benchmarks that decide the partition-kernel design on B200:
// how fast are shared-memory atomics (with / without return value, narrow / wide histograms),
// warp match (__match_any_sync) and plain streaming reads. Standalone: nvcc ubench.cu -o ubench.
// GPU analogue of Scan-Micro-Benchmarks/microbenchmarks/RadixPartitioning (histogram / scatter loop
// variants, Shared/histogram_algorithms.hpp:10-100).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <bool RET>
__global__ void k_atoms(uint32_t *sink, int iters, uint32_t mask, size_t smem_words) {
    extern __shared__ uint32_t sh[];
    for (uint32_t i = threadIdx.x; i < smem_words; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t s = blockIdx.x * 9781u + threadIdx.x * 6151u + 1, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t d = lcg(s) & mask;
            if (RET) acc += atomicAdd(&sh[d], 1u); else atomicAdd(&sh[d], 1u);
        }
    }
    __syncthreads();
    if (acc == 0xdeadbeef || sh[threadIdx.x & mask] == 0xdeadbeef) sink[0] = acc;
}

__global__ void k_match(uint32_t *sink, int iters, uint32_t mask) {
    uint32_t s = blockIdx.x * 9781u + threadIdx.x * 6151u + 1, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t d = lcg(s) & mask;
            acc += __popc(__match_any_sync(0xffffffffu, d));
        }
    }
    if (acc == 0xdeadbeef) sink[0] = acc;
}

// ballot-based peer mask for a `bits`-bit digit
__global__ void k_ballot(uint32_t *sink, int iters, uint32_t mask, int bits) {
    uint32_t s = blockIdx.x * 9781u + threadIdx.x * 6151u + 1, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t d = lcg(s) & mask;
            uint32_t peers = 0xffffffffu;
            for (int b = 0; b < bits; ++b) {
                uint32_t v = __ballot_sync(0xffffffffu, (d >> b) & 1);
                peers &= ((d >> b) & 1) ? v : ~v;
            }
            acc += __popc(peers);
        }
    }
    if (acc == 0xdeadbeef) sink[0] = acc;
}

__global__ void k_lcg_only(uint32_t *sink, int iters, uint32_t mask) {
    uint32_t s = blockIdx.x * 9781u + threadIdx.x * 6151u + 1, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += lcg(s) & mask;
    }
    if (acc == 0xdeadbeef) sink[0] = acc;
}

__global__ void k_read(const uint4 *in, size_t n, uint32_t *sink) {
    uint32_t acc = 0;
    size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    if (acc == 0xdeadbeef) sink[0] = acc;
}

__global__ void k_copy(const uint4 *in, uint4 *out, size_t n) {
    size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
    }
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    uint32_t *sink; CK(cudaMalloc(&sink, 64));
    const int iters = 2000, threads = 256;
    const double ops_per_block = (double) iters * 8 * threads;
    for (int bps : {2, 4, 8}) {
        int grid = p.multiProcessorCount * bps;
        double total = ops_per_block * grid;
        float base = time_ms([&] { k_lcg_only<<<grid, threads>>>(sink, iters, 127); });
        printf("[blocks/SM=%d] lcg only: %.1f Gop/s\n", bps, total / base * 1e-6);
        for (int bits : {7, 8, 11, 14}) {
            uint32_t mask = (1u << bits) - 1;
            size_t words = (size_t) 1 << bits;
            if (words * 4 * bps > 200 * 1024) continue;
            CK(cudaFuncSetAttribute(k_atoms<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            CK(cudaFuncSetAttribute(k_atoms<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            float t1 = time_ms([&] { k_atoms<true><<<grid, threads, words * 4>>>(sink, iters, mask, words); });
            float t0 = time_ms([&] { k_atoms<false><<<grid, threads, words * 4>>>(sink, iters, mask, words); });
            printf("[blocks/SM=%d] smem atomicAdd bins=2^%d: with return %.1f Gop/s, no return %.1f Gop/s\n", bps, bits,
                   total / t1 * 1e-6, total / t0 * 1e-6);
        }
        float tm = time_ms([&] { k_match<<<grid, threads>>>(sink, iters, 127); });
        float tb = time_ms([&] { k_ballot<<<grid, threads>>>(sink, iters, 127, 7); });
        printf("[blocks/SM=%d] match_any 7-bit: %.1f Gop/s ; 7x ballot: %.1f Gop/s\n", bps, total / tm * 1e-6, total / tb * 1e-6);
    }
    size_t n = (size_t) 1 << 26;   // 1 GiB of uint4
    uint4 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
    CK(cudaMemset(a, 1, n * 16)); CK(cudaMemset(b, 2, n * 16));
    for (int bps : {4, 8, 16}) {
        int grid = p.multiProcessorCount * bps;
        float tr = time_ms([&] { k_read<<<grid, 256>>>(a, n, sink); });
        float tc = time_ms([&] { k_copy<<<grid, 256>>>(a, b, n); });
        printf("[blocks/SM=%d] read 1GiB: %.1f GB/s ; copy 1GiB->1GiB: %.1f GB/s (r+w)\n", bps, n * 16.0 / tr * 1e-6,
               2.0 * n * 16.0 / tc * 1e-6);
    }
    return 0;
}
