// block_scan.cuh — single-CTA exclusive scan used by the small planning kernels.
#pragma once

#include "common.cuh"

namespace aqp {

constexpr int kScanBlock = 1024;
constexpr int kScanRounds = 16;   // values per thread held in registers at once

// Exclusive scan of get(i), i in [0,n), by one 1024-thread block; calls put(i, exclusive_prefix)
// and returns the grand total to every thread.
// The planning kernels sit on the critical path between two bandwidth kernels, so latency is what
// counts: thread t takes elements t, t + 1024, ... (coalesced), ALL of a 16384-element span's loads are
// issued before the first one is used, and each 1024-element round then costs one warp scan, one
// barrier and a redundant 32-value scan of the warp totals in every warp (double-buffered, so one
// barrier per round suffices). The first version gave each thread 16 consecutive elements and read
// them twice, thread-serially: 35-93 us per call under ncu, now a few us.
template <typename Get, typename Put>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t n, Get get, Put put) {
    __shared__ uint32_t wsum[2][kScanBlock / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t carry = 0, buf = 0;
    for (uint32_t base = 0; base < n; base += kScanBlock * kScanRounds) {
        const uint32_t span = min(n - base, (uint32_t) (kScanBlock * kScanRounds));
        const uint32_t rounds = (span + kScanBlock - 1) / kScanBlock;
        uint32_t v[kScanRounds];
#pragma unroll
        for (int j = 0; j < kScanRounds; ++j) {
            const uint32_t i = base + j * kScanBlock + threadIdx.x;
            v[j] = i < n ? get(i) : 0u;
        }
#pragma unroll
        for (int j = 0; j < kScanRounds; ++j) {
            if (j < (int) rounds) {   // block-uniform
                const uint32_t i = base + j * kScanBlock + threadIdx.x;
                const uint32_t incl = warp_incl_scan(v[j]);
                if (lane == 31) wsum[buf][warp] = incl;
                __syncthreads();
                const uint32_t w = wsum[buf][lane];
                const uint32_t wi = warp_incl_scan(w);
                const uint32_t wprefix = __shfl_sync(0xffffffffu, wi - w, warp);
                const uint32_t round_total = __shfl_sync(0xffffffffu, wi, 31);
                if (i < n) put(i, carry + wprefix + incl - v[j]);
                carry += round_total;
                buf ^= 1u;
            }
        }
    }
    __syncthreads();   // callers may read what put() wrote
    return carry;
}

}  // namespace aqp
