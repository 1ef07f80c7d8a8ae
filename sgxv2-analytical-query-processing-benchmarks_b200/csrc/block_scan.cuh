// block_scan.cuh — single-CTA exclusive scan used by the small planning kernels.
#pragma once

#include "common.cuh"

namespace aqp {

constexpr int kScanBlock = 1024;

// Exclusive scan of get(i), i in [0,n), by one 1024-thread block; calls put(i, exclusive_prefix)
// and returns the grand total to every thread.
template <typename Get, typename Put>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t n, Get get, Put put) {
    __shared__ uint32_t wsum[kScanBlock / 32];
    __shared__ uint32_t s_total;
    const uint32_t per = (n + kScanBlock - 1) / kScanBlock;
    const uint32_t b = threadIdx.x * per, e = min(n, b + per);
    uint32_t local = 0;
    for (uint32_t i = b; i < e; ++i) local += get(i);
    uint32_t incl = warp_incl_scan(local);
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = wsum[threadIdx.x];
        uint32_t wi = warp_incl_scan(w);
        wsum[threadIdx.x] = wi - w;
        if (threadIdx.x == 31) s_total = wi;
    }
    __syncthreads();
    uint32_t run = wsum[threadIdx.x >> 5] + incl - local;
    for (uint32_t i = b; i < e; ++i) {
        uint32_t v = get(i);
        put(i, run);
        run += v;
    }
    uint32_t total = s_total;
    __syncthreads();
    return total;
}

}  // namespace aqp
