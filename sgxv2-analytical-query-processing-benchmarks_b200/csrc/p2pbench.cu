// p2pbench.cu — NVLink peer-store micro-benchmark for the fused scatter+exchange design (needs 2 GPUs).
// How fast can SM-issued stores push data into a peer's HBM, as a function of run length and alignment?
//   (a) cudaMemcpyPeerAsync (copy engines) as the reference
//   (b) contiguous 16-byte stores from a kernel
//   (c) runs of L tuples (8 B each) at arbitrary 8-byte-aligned offsets, 8-byte stores per lane —
//       the access pattern of radix_scatter_kernel's write-out
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_copy16(const uint4 *in, uint4 *out, size_t n) {
    size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}
// each warp writes runs of `len` tuples; run r goes to offset r*(len+gap) (gap = 1 tuple breaks alignment)
__global__ void k_runs(const uint2 *in, uint2 *out, size_t ntuples, int len, int gap) {
    size_t warp = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    size_t nruns = ntuples / len;
    for (size_t r = warp; r < nruns; r += nwarps) {
        size_t src = r * len, dst = r * (size_t) (len + gap);
        for (int k = lane; k < len; k += 32) out[dst + k] = in[src + k];
    }
}
template <typename F> float time_ms(F f, int dev) {
    cudaEvent_t a, b; CK(cudaSetDevice(dev)); cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    size_t bytes = (size_t) 1 << 30;
    void *src, *dst_peer, *dst_local;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&dst_peer, bytes * 2));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst_local, bytes * 2));
    CK(cudaMemset(src, 1, bytes));
    float t = time_ms([&] { cudaMemcpyPeerAsync(dst_peer, 1, src, 0, bytes, 0); }, 0);
    printf("cudaMemcpyPeer 1 GiB: %.1f GB/s\n", bytes / t * 1e-6);
    for (int bps : {2, 4, 8}) {
        t = time_ms([&] { k_copy16<<<148 * bps, 256>>>((const uint4 *) src, (uint4 *) dst_peer, bytes / 16); }, 0);
        float tl = time_ms([&] { k_copy16<<<148 * bps, 256>>>((const uint4 *) src, (uint4 *) dst_local, bytes / 16); }, 0);
        printf("[blocks/SM=%d] kernel contiguous 16B stores: peer %.1f GB/s, local %.1f GB/s (write side)\n", bps, bytes / t * 1e-6, bytes / tl * 1e-6);
    }
    for (int len : {8, 16, 32, 64, 128, 512, 2048}) {
        for (int gap : {0, 1}) {
            t = time_ms([&] { k_runs<<<148 * 8, 256>>>((const uint2 *) src, (uint2 *) dst_peer, bytes / 8, len, gap); }, 0);
            float tl = time_ms([&] { k_runs<<<148 * 8, 256>>>((const uint2 *) src, (uint2 *) dst_local, bytes / 8, len, gap); }, 0);
            printf("runs of %4d tuples (%5d B), gap %d: peer %.1f GB/s, local %.1f GB/s\n", len, len * 8, gap, bytes / t * 1e-6, bytes / tl * 1e-6);
        }
    }
    return 0;
}
