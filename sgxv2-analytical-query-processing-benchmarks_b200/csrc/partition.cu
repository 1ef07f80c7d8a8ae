// partition.cu — radix partitioning of 8-byte {key,payload} tuples (sm_100a).
//
// Replaces the partitioning loops of Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp:
//   partition_hist[_unrolled]      :617-654  -> radix_hist_kernel   (ONE read of each relation yields
//                                               the histogram over ALL radix bits of both passes)
//   prefix sums                    :886-915, -> plan_offsets_kernel (partition boundaries of both
//                                   :739-746     passes from that one histogram)
//   partition_copy[_unrolled]      :659-697  -> radix_scatter_kernel (pass 1: whole relation, one
//   radix_cluster/serial_radix_..  :715-841     segment; pass 2: every pass-1 partition is a segment)
//
// Digit function is the reference's: (key & MASK) >> R on raw key bits, pass 1 on bits [0,b1),
// pass 2 on bits [b1,b1+b2) (radix_join.cpp:47,:1118-1119,:1262).
//
// Scatter design: a CTA takes a tile of the input, ranks every tuple inside its partition with a
// shared-memory atomic counter, reorders the tile in shared memory so that tuples of one
// partition are adjacent, reserves the tile's run in every partition's output range with ONE
// global atomic per (tile, partition), and writes the runs out as contiguous coalesced stores
// (software write-combining; the analogue of the reference's dormant SWWC path :1013-1053).
// Order inside a partition is therefore unspecified — exactly like the reference, where it
// depends on thread interleaving — and never affects join results.
#include "common.cuh"
#include "join_internal.cuh"
#include "block_scan.cuh"

namespace aqp {

// ---------------------------------------------------------------------------------------------
// histogram over `bits` key bits starting at `shift`
// ---------------------------------------------------------------------------------------------
constexpr int kHistThreads = 512;
constexpr int kHistUnroll = 8;   // 8-byte loads in flight per thread

// shared-memory privatised histogram (bits <= kMaxSmemHistBits), flushed with global REDs
__global__ void __launch_bounds__(kHistThreads)
radix_hist_smem_kernel(const uint2 *__restrict__ in, uint64_t n, uint32_t shift, uint32_t bits,
                       uint32_t *__restrict__ ghist) {
    extern __shared__ uint32_t sh[];
    const uint32_t fan = 1u << bits, mask = fan - 1;
    for (uint32_t i = threadIdx.x; i < fan; i += kHistThreads) sh[i] = 0;
    __syncthreads();
    const uint64_t tile = (uint64_t) kHistThreads * kHistUnroll;
    for (uint64_t base = (uint64_t) blockIdx.x * tile; base < n; base += (uint64_t) gridDim.x * tile) {
        uint2 v[kHistUnroll];
#pragma unroll
        for (int j = 0; j < kHistUnroll; ++j) {
            uint64_t i = base + (uint64_t) j * kHistThreads + threadIdx.x;
            if (i < n) v[j] = ld_stream_v2(in + i);
        }
#pragma unroll
        for (int j = 0; j < kHistUnroll; ++j) {
            uint64_t i = base + (uint64_t) j * kHistThreads + threadIdx.x;
            if (i < n) atomicAdd(&sh[(v[j].x >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < fan; i += kHistThreads) {
        uint32_t c = sh[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// fallback for very wide histograms: global REDs only
__global__ void __launch_bounds__(kHistThreads)
radix_hist_global_kernel(const uint2 *__restrict__ in, uint64_t n, uint32_t shift, uint32_t bits,
                         uint32_t *__restrict__ ghist) {
    const uint32_t mask = (1u << bits) - 1;
    uint64_t stride = (uint64_t) gridDim.x * kHistThreads;
    for (uint64_t i = (uint64_t) blockIdx.x * kHistThreads + threadIdx.x; i < n; i += stride)
        atomicAdd(&ghist[(ld_stream_v2(in + i).x >> shift) & mask], 1u);
}

int radix_hist_device(const row_t *d_in, uint64_t n, uint32_t shift, uint32_t bits, uint32_t *d_hist,
                      cudaStream_t st) {
    if (n == 0) return 0;
    if (bits > 24) {
        set_error("radix_hist: bits > 24 not supported");
        return -1;
    }
    const uint2 *in = reinterpret_cast<const uint2 *>(d_in);
    if (bits <= kMaxSmemHistBits) {
        size_t smem = sizeof(uint32_t) << bits;
        static bool attr_set = false;
        if (!attr_set) {
            AQP_CUDA_OK(cudaFuncSetAttribute(radix_hist_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) (sizeof(uint32_t) << kMaxSmemHistBits)));
            attr_set = true;
        }
        uint64_t tiles = (n + (uint64_t) kHistThreads * kHistUnroll - 1) / ((uint64_t) kHistThreads * kHistUnroll);
        int per_sm = bits <= 13 ? 3 : (bits == 14 ? 2 : 1);
        int grid = (int) (tiles < (uint64_t) kNumSMs * per_sm ? tiles : (uint64_t) kNumSMs * per_sm);
        radix_hist_smem_kernel<<<grid, kHistThreads, smem, st>>>(in, n, shift, bits, d_hist);
    } else {
        radix_hist_global_kernel<<<kNumSMs * 4, kHistThreads, 0, st>>>(in, n, shift, bits, d_hist);
    }
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// single-block exclusive scan helpers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanBlock)
exclusive_scan_u32_kernel(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out) {
    uint32_t total = block_exclusive_scan(n, [&](uint32_t i) { return in[i]; }, [&](uint32_t i, uint32_t v) { out[i] = v; });
    if (threadIdx.x == 0) out[n] = total;
}

int exclusive_scan_u32_device(const uint32_t *d_in, uint32_t n, uint32_t *d_out, cudaStream_t st) {
    exclusive_scan_u32_kernel<<<1, kScanBlock, 0, st>>>(d_in, n, d_out);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// plan: all partition boundaries of both passes from the full-width histogram.
//   hist is indexed by the raw digit D = key & (2^B-1) = p1 | (p2 << b1)
//   final partitions are ordered by f = p1 * F2 + p2 (pass-1 partition major)
// blockIdx.x selects the relation (0 = R, 1 = S).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanBlock) plan_offsets_kernel(PlanArgs a) {
    const RelPlan &r = a.rel[blockIdx.x];
    const uint32_t b1 = a.bits1, F2 = 1u << a.bits2, P = 1u << (a.bits1 + a.bits2), F1 = 1u << a.bits1;
    auto digit_of = [&](uint32_t f) { return (f / F2) | ((f % F2) << b1); };
    uint32_t total = block_exclusive_scan(
        P, [&](uint32_t f) { return r.hist[digit_of(f)]; },
        [&](uint32_t f, uint32_t v) {
            r.part_off[f] = v;
            r.cursor2[f] = v;
            if (f % F2 == 0) {
                r.cursor1[f / F2] = v;
                r.seg_off[f / F2] = v;
            }
        });
    if (threadIdx.x == 0) {
        r.part_off[P] = total;
        r.seg_off[F1] = total;
        r.seg1[0] = 0;
        r.seg1[1] = total;
        r.seg1[2] = 0;
        r.seg1[3] = (total + kScatterTile - 1) / kScatterTile;
    }
    __syncthreads();
    // tiles of the pass-2 scatter: every pass-1 partition is cut into kScatterTile-tuple tiles
    uint32_t tiles = block_exclusive_scan(
        F1,
        [&](uint32_t s) {
            uint32_t len = r.seg_off[s + 1] - r.seg_off[s];
            return (len + kScatterTile - 1) / kScatterTile;
        },
        [&](uint32_t s, uint32_t v) { r.seg_tile_start[s] = v; });
    if (threadIdx.x == 0) r.seg_tile_start[F1] = tiles;
}

int plan_offsets_device(const PlanArgs &a, cudaStream_t st) {
    plan_offsets_kernel<<<2, kScanBlock, 0, st>>>(a);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// scatter
// ---------------------------------------------------------------------------------------------
constexpr int kScatterThreads = 256;
constexpr int kScatterItems = kScatterTile / kScatterThreads;   // 16 tuples per thread
static_assert(kScatterItems * kScatterThreads == kScatterTile, "tile shape");
static_assert(kScatterThreads >= kMaxFanout, "one thread per bin in the bookkeeping steps");

__global__ void __launch_bounds__(kScatterThreads, 3)
radix_scatter_kernel(const uint2 *__restrict__ in, uint2 *__restrict__ out,
                     const uint32_t *__restrict__ seg_off, const uint32_t *__restrict__ seg_tile_start,
                     uint32_t nseg, uint32_t shift, uint32_t bits, uint32_t *__restrict__ cursors) {
    __shared__ uint2 stage[kScatterTile];
    __shared__ uint32_t cnt[kMaxFanout];     // tuples of this tile per partition
    __shared__ uint32_t lbase[kMaxFanout];   // start of the partition's run inside `stage`
    __shared__ uint32_t gdst[kMaxFanout];    // global index of stage slot s of partition d is gdst[d] + s
    __shared__ uint32_t s_tstart[kMaxFanout + 1];
    __shared__ uint32_t s_soff[kMaxFanout + 1];

    const uint32_t fan = 1u << bits, mask = fan - 1;
    for (uint32_t i = threadIdx.x; i <= nseg; i += kScatterThreads) {
        s_tstart[i] = seg_tile_start[i];
        s_soff[i] = seg_off[i];
    }
    __syncthreads();
    const uint32_t ntiles = s_tstart[nseg];

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // segment of this tile: last s with s_tstart[s] <= tile
        uint32_t lo = 0, hi = nseg;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_tstart[mid] <= tile) lo = mid; else hi = mid;
        }
        const uint32_t seg = lo;
        const uint32_t begin = s_soff[seg] + (tile - s_tstart[seg]) * kScatterTile;
        const uint32_t end = min(begin + (uint32_t) kScatterTile, s_soff[seg + 1]);
        const uint32_t ntile = end - begin;

        uint2 v[kScatterItems];
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t i = begin + j * kScatterThreads + threadIdx.x;
            if (i < end) v[j] = ld_stream_v2(in + i);
        }
        if (threadIdx.x < fan) cnt[threadIdx.x] = 0;
        __syncthreads();   // (A) previous tile fully written out, counters cleared

        uint32_t rank[kScatterItems];
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t i = begin + j * kScatterThreads + threadIdx.x;
            if (i < end) rank[j] = atomicAdd(&cnt[(v[j].x >> shift) & mask], 1u);
        }
        __syncthreads();   // (B) tile histogram complete

        // warp 0: exclusive scan of the tile histogram, then reserve the runs globally
        uint32_t my_g[kMaxFanout / 32], my_b[kMaxFanout / 32];
        if (threadIdx.x < 32) {
            const uint32_t per = (fan + 31) / 32;   // consecutive bins per lane (<= 8)
            uint32_t c[kMaxFanout / 32], sum = 0;
#pragma unroll
            for (int k = 0; k < kMaxFanout / 32; ++k) {
                uint32_t d = threadIdx.x * per + k;
                c[k] = (k < (int) per && d < fan) ? cnt[d] : 0;
                sum += c[k];
            }
            uint32_t run = warp_incl_scan(sum) - sum;
#pragma unroll
            for (int k = 0; k < kMaxFanout / 32; ++k) {
                uint32_t d = threadIdx.x * per + k;
                my_b[k] = run;
                if (k < (int) per && d < fan) {
                    lbase[d] = run;
                    my_g[k] = c[k] ? atomicAdd(&cursors[(seg << bits) + d], c[k]) : 0u;
                }
                run += c[k];
            }
        }
        __syncthreads();   // (C) lbase ready

#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t i = begin + j * kScatterThreads + threadIdx.x;
            if (i < end) stage[lbase[(v[j].x >> shift) & mask] + rank[j]] = v[j];
        }
        if (threadIdx.x < 32) {
            const uint32_t per = (fan + 31) / 32;
#pragma unroll
            for (int k = 0; k < kMaxFanout / 32; ++k) {
                uint32_t d = threadIdx.x * per + k;
                if (k < (int) per && d < fan) gdst[d] = my_g[k] - my_b[k];
            }
        }
        __syncthreads();   // (D) tile reordered, destinations known

        for (uint32_t s = threadIdx.x; s < ntile; s += kScatterThreads) {
            uint2 t = stage[s];
            out[gdst[(t.x >> shift) & mask] + s] = t;
        }
    }
}

int radix_scatter_launch(const row_t *d_in, row_t *d_out, const uint32_t *d_seg_off,
                         const uint32_t *d_seg_tile_start, uint32_t nseg, uint64_t n_total, uint32_t shift,
                         uint32_t bits, uint32_t *d_cursors, cudaStream_t st) {
    if (bits > kMaxFanoutBits || nseg > kMaxFanout) {
        set_error("radix_scatter: fan-out too large");
        return -1;
    }
    if (n_total == 0) return 0;
    uint64_t max_tiles = n_total / kScatterTile + nseg;
    int grid = (int) (max_tiles < (uint64_t) kNumSMs * 3 ? max_tiles : (uint64_t) kNumSMs * 3);
    radix_scatter_kernel<<<grid, kScatterThreads, 0, st>>>(reinterpret_cast<const uint2 *>(d_in),
                                                          reinterpret_cast<uint2 *>(d_out), d_seg_off,
                                                          d_seg_tile_start, nseg, shift, bits, d_cursors);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// single-segment helpers for the stage-level C API: builds seg tables {0,n} / {0,ceil(n/tile)} and
// copies offsets into the cursor array.
__global__ void single_segment_setup_kernel(uint32_t n, const uint32_t *offsets, uint32_t fan, uint32_t *cursors,
                                            uint32_t *seg_tables /* [4] */) {
    for (uint32_t i = threadIdx.x; i < fan; i += blockDim.x) cursors[i] = offsets[i];
    if (threadIdx.x == 0) {
        seg_tables[0] = 0;
        seg_tables[1] = n;
        seg_tables[2] = 0;
        seg_tables[3] = (n + kScatterTile - 1) / kScatterTile;
    }
}

int single_segment_setup(uint32_t n, const uint32_t *d_offsets, uint32_t fan, uint32_t *d_cursors,
                         uint32_t *d_seg_tables, cudaStream_t st) {
    single_segment_setup_kernel<<<1, 256, 0, st>>>(n, d_offsets, fan, d_cursors, d_seg_tables);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace aqp
