// bulkbench.cu — micro-benchmark behind the scatter kernel's write-out design: a CTA holds a 32 KiB staging
// buffer of partition runs in shared memory and has to put each run (R bytes, R = 128 ... 2048) at its own place
// in HBM. Variant A: the CTA's threads store the runs with coalesced 8-byte SM stores (what radix_scatter_kernel
// does). Variant B: one thread per run issues a TMA bulk store (cp.async.bulk.global.shared::cta) and the CTA
// waits for the bulk group to have read shared memory. Destinations are 16-byte aligned and pseudo-random inside
// a 4 GiB buffer, so nothing is cache-resident. Standalone: nvcc -gencode arch=compute_100a,code=sm_100a bulkbench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int kThreads = 256;
constexpr int kStageBytes = 32768;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <bool kBulk>
__global__ void __launch_bounds__(kThreads, 2) k_runs(unsigned char *out, size_t out_bytes, int run_bytes, int iters,
                                                      int misalign) {
    extern __shared__ __align__(128) unsigned char stage[];
    for (int i = threadIdx.x; i < kStageBytes / 8; i += kThreads) reinterpret_cast<uint2 *>(stage)[i] = make_uint2(i, blockIdx.x);
    __syncthreads();
    const int nruns = kStageBytes / run_bytes;
    const size_t nslots = out_bytes / run_bytes - 1;
    // misalign: every run starts 16..112 bytes into a 128-byte line (what the scatter's 8-byte granular runs look like)
    auto shift = [&](uint32_t r) { return misalign ? 16u * (1u + (mix(r * 2246822519u) % 7u)) : 0u; };
    for (int it = 0; it < iters; ++it) {
        if (kBulk) {
            for (int r = threadIdx.x; r < nruns; r += kThreads) {
                size_t slot = mix((uint32_t) (it * 1315423911u) ^ (blockIdx.x * 2654435761u) ^ (uint32_t) r * 40503u) % nslots;
                unsigned char *dst = out + slot * (size_t) run_bytes + shift((uint32_t) r + it);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                             "r"(smem_u32(stage + (size_t) r * run_bytes)), "r"(run_bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();   // the staging buffer may be refilled now
        } else {
            const int per_run = run_bytes / 8;
            for (int s = threadIdx.x; s < kStageBytes / 8; s += kThreads) {
                int r = s / per_run, o = s % per_run;
                size_t slot = mix((uint32_t) (it * 1315423911u) ^ (blockIdx.x * 2654435761u) ^ (uint32_t) r * 40503u) % nslots;
                reinterpret_cast<uint2 *>(out + slot * (size_t) run_bytes + shift((uint32_t) r + it))[o] = reinterpret_cast<uint2 *>(stage)[s];
            }
            __syncthreads();
        }
    }
}

int main(int argc, char **argv) {
    int dev = 0, peer = argc > 1 ? atoi(argv[1]) : -1;
    CK(cudaSetDevice(dev));
    const size_t out_bytes = (size_t) 4 << 30;
    unsigned char *out;
    if (peer >= 0) {
        CK(cudaSetDevice(peer));
        CK(cudaMalloc(&out, out_bytes));
        CK(cudaSetDevice(dev));
        CK(cudaDeviceEnablePeerAccess(peer, 0));
        printf("destination: device %d (peer over NVLink)\n", peer);
    } else {
        CK(cudaMalloc(&out, out_bytes));
        printf("destination: local HBM\n");
    }
    CK(cudaFuncSetAttribute(k_runs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
    CK(cudaFuncSetAttribute(k_runs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = 148 * 2, iters = peer >= 0 ? 200 : 1000;
    for (int misalign = 0; misalign < 2; ++misalign)
    for (int run_bytes : {128, 256, 512, 1024, 2048}) {
        float ms[2];
        for (int v = 0; v < 2; ++v) {
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                if (v) k_runs<true><<<grid, kThreads, kStageBytes>>>(out, out_bytes, run_bytes, iters, misalign);
                else k_runs<false><<<grid, kThreads, kStageBytes>>>(out, out_bytes, run_bytes, iters, misalign);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                CK(cudaEventElapsedTime(&ms[v], e0, e1));
            }
        }
        double bytes = (double) grid * iters * kStageBytes;
        printf("%s run %4d B: SM stores %7.1f GB/s   TMA bulk stores %7.1f GB/s\n", misalign ? "16B-aligned " : "line-aligned", run_bytes, bytes / ms[0] / 1e6, bytes / ms[1] / 1e6);
    }
    return 0;
}
