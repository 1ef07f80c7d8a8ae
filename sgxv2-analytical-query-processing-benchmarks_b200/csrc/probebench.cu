// probebench.cu — micro-benchmark behind one design decision (not part of the product): how fast can a probe stream
// look up a hash table that lives in L2 instead of shared memory? VERDICT r01 task 9 proposes to skip pass 2 for S:
// probe S's pass-1 partitions (4 M tuples each at 2^27 x 2^29) against a direct-mapped table of the matching R
// partition (1 M tuples = 8 MiB) in global memory, saving S's pass-2 read + write (16 of 48 B/tuple). That only pays
// if 2^29 random 8-byte lookups out of L2 cost less than the 1.5 ms the S half of pass 2 takes today.
// Every thread streams 4 S tuples (coalesced 8-byte loads) and reads table[key & mask] for each; tables from 1 MiB
// to 1 GiB show the L2 -> HBM transition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/probebench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__global__ void k_fill_s(uint2 *s, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        s[i] = make_uint2(mix((uint32_t) i), (uint32_t) i);
}
__global__ void k_fill_t(uint2 *t, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        t[i] = make_uint2((uint32_t) i, (uint32_t) i * 3u);
}
template <int kUnroll>
__global__ void __launch_bounds__(512) k_probe(const uint2 *__restrict__ s, size_t n, const uint2 *__restrict__ table, uint32_t mask,
                                               unsigned long long *out) {
    unsigned long long matches = 0, sum = 0;
    const size_t stride = (size_t) gridDim.x * blockDim.x * kUnroll;
    for (size_t base = (size_t) blockIdx.x * blockDim.x * kUnroll + threadIdx.x; base < n; base += stride) {
        uint2 v[kUnroll], r[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            size_t i = base + (size_t) j * blockDim.x;
            v[j] = i < n ? __ldcs(s + i) : make_uint2(0, 0);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) r[j] = __ldg(table + (v[j].x & mask));
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            if (r[j].x == (v[j].x & mask)) {
                ++matches;
                sum += (unsigned long long) r[j].y + v[j].y;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        matches += __shfl_down_sync(0xffffffffu, matches, o);
        sum += __shfl_down_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, matches);
        atomicAdd(out + 1, sum);
    }
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs=%d L2=%d MiB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
    const size_t nS = (size_t) 1 << 28;   // 2 GiB of S tuples: larger than L2, like a pass-1 output
    uint2 *s, *t;
    unsigned long long *out;
    CK(cudaMalloc(&s, nS * 8));
    CK(cudaMalloc(&t, (size_t) 1 << 30));
    CK(cudaMalloc(&out, 16));
    k_fill_s<<<p.multiProcessorCount * 8, 256>>>(s, nS);
    k_fill_t<<<p.multiProcessorCount * 8, 256>>>(t, (size_t) 1 << 27);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("%-12s %10s %12s %12s\n", "table", "ms", "Glookups/s", "S GB/s");
    for (int lg = 17; lg <= 27; lg += (lg < 23 ? 3 : 1)) {   // entries: 2^17 (1 MiB) .. 2^27 (1 GiB)
        const uint32_t mask = (1u << lg) - 1;
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemset(out, 0, 16));
            CK(cudaEventRecord(e0));
            k_probe<4><<<p.multiProcessorCount * 4, 512>>>(s, nS, t, mask, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        unsigned long long h[2];
        CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
        char name[32];
        snprintf(name, sizeof name, "%d MiB", (8 << lg) >> 20);
        printf("%-12s %10.3f %12.1f %12.1f   matches=%llu\n", name, best, nS / best * 1e-6, nS * 8.0 / best * 1e-6, h[0]);
    }
    // reference point: the same stream without any lookup
    {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemset(out, 0, 16));
            CK(cudaEventRecord(e0));
            k_probe<4><<<p.multiProcessorCount * 4, 512>>>(s, nS, t, 0u, out);   // mask 0: every lookup hits one line
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        printf("%-12s %10.3f %12.1f %12.1f\n", "one line", best, nS / best * 1e-6, nS * 8.0 / best * 1e-6);
    }
    return 0;
}
