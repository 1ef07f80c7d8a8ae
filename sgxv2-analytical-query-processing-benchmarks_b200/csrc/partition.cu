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

// shared-memory privatised histogram (bits <= kMaxSmemHistBits), flushed with global REDs.
// Block b reads the contiguous chunk [b*chunk, (b+1)*chunk) — the same chunk the pass-1 scatter
// assigns to block b — and, besides adding into the global full-width histogram, emits its own
// pass-1 histogram row block_hist[b][p1] (p1 = low bits1 of the digit). Those rows give every
// scatter block private, pre-reserved output ranges (parallel_radix_partition's per-thread
// histogram + prefix scheme, radix_join.cpp:881-915, with CTAs in the role of threads).
__global__ void __launch_bounds__(kHistThreads)
radix_hist_smem_kernel(const uint2 *__restrict__ in, uint64_t n, uint32_t shift, uint32_t bits,
                       uint32_t *__restrict__ ghist, uint64_t chunk, uint32_t bits1,
                       uint32_t *__restrict__ block_hist) {
    extern __shared__ uint32_t sh[];
    const uint32_t fan = 1u << bits, mask = fan - 1;
    for (uint32_t i = threadIdx.x; i < fan; i += kHistThreads) sh[i] = 0;
    __syncthreads();
    const uint64_t cbeg = (uint64_t) blockIdx.x * chunk;
    const uint64_t cend = cbeg + chunk < n ? cbeg + chunk : n;
    const uint64_t tile = (uint64_t) kHistThreads * kHistUnroll;
    for (uint64_t base = cbeg; base < cend; base += tile) {
        uint2 v[kHistUnroll];
#pragma unroll
        for (int j = 0; j < kHistUnroll; ++j) {
            uint64_t i = base + (uint64_t) j * kHistThreads + threadIdx.x;
            if (i < cend) v[j] = ld_stream_v2(in + i);
        }
#pragma unroll
        for (int j = 0; j < kHistUnroll; ++j) {
            uint64_t i = base + (uint64_t) j * kHistThreads + threadIdx.x;
            if (i < cend) atomicAdd(&sh[(v[j].x >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < fan; i += kHistThreads) {
        uint32_t c = sh[i];
        if (c) atomicAdd(&ghist[i], c);
    }
    if (block_hist) {
        const uint32_t F1 = 1u << bits1, F2 = fan >> bits1;
        for (uint32_t p1 = threadIdx.x; p1 < F1; p1 += kHistThreads) {
            uint32_t c = 0;
            for (uint32_t p2 = 0; p2 < F2; ++p2) c += sh[p1 + (p2 << bits1)];
            block_hist[(size_t) blockIdx.x * F1 + p1] = c;
        }
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

// nblocks == 0: plain histogram with a grid chosen here. nblocks > 0: exactly nblocks CTAs, CTA b
// covering tuples [b*chunk, (b+1)*chunk), and block_hist[nblocks][2^bits1] rows are written.
int radix_hist_device(const row_t *d_in, uint64_t n, uint32_t shift, uint32_t bits, uint32_t *d_hist,
                      uint32_t nblocks, uint64_t chunk, uint32_t bits1, uint32_t *d_block_hist, cudaStream_t st) {
    if (bits > 24) {
        set_error("radix_hist: bits > 24 not supported");
        return -1;
    }
    if (nblocks && bits > (uint32_t) kMaxSmemHistBits) {
        set_error("radix_hist: per-block histograms need bits <= 15");
        return -1;
    }
    if (n == 0 && !nblocks) return 0;
    const uint2 *in = reinterpret_cast<const uint2 *>(d_in);
    if (bits <= (uint32_t) kMaxSmemHistBits) {
        size_t smem = sizeof(uint32_t) << bits;
        static bool attr_set = false;
        if (!attr_set) {
            AQP_CUDA_OK(cudaFuncSetAttribute(radix_hist_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) (sizeof(uint32_t) << kMaxSmemHistBits)));
            attr_set = true;
        }
        uint32_t grid = nblocks;
        if (!grid) {
            const uint64_t tile = (uint64_t) kHistThreads * kHistUnroll;
            uint64_t tiles = (n + tile - 1) / tile;
            int per_sm = bits <= 13 ? 3 : (bits == 14 ? 2 : 1);
            grid = (uint32_t) (tiles < (uint64_t) kNumSMs * per_sm ? tiles : (uint64_t) kNumSMs * per_sm);
            chunk = (tiles + grid - 1) / grid * tile;
        }
        radix_hist_smem_kernel<<<grid, kHistThreads, smem, st>>>(in, n, shift, bits, d_hist, chunk, bits1,
                                                                nblocks ? d_block_hist : nullptr);
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
    // private pass-1 cursors: block b starts writing partition p1 at
    //   part_start[p1] + sum over earlier blocks of their pass-1 histogram rows (radix_join.cpp:901-915)
    if (r.block_hist) {
        __syncthreads();
        for (uint32_t p1 = threadIdx.x; p1 < F1; p1 += kScanBlock) {
            uint32_t run = r.cursor1[p1];
            for (uint32_t b = 0; b < a.nblocks1; ++b) {
                uint32_t c = r.block_hist[(size_t) b * F1 + p1];
                r.block_base[(size_t) b * F1 + p1] = run;
                run += c;
            }
        }
    }
}

int plan_offsets_device(const PlanArgs &a, cudaStream_t st) {
    plan_offsets_kernel<<<2, kScanBlock, 0, st>>>(a);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// scatter
//
// A CTA walks its tiles with the NEXT tile's tuples already in flight into registers (software
// pipelining: the loads are issued before the current tile is reordered and written out), so HBM
// reads overlap everything else. Per tile:
//   rank      every tuple takes a slot in its partition with a shared-memory atomicAdd
//   scan      every warp scans the tile histogram for itself (redundant, but it saves a block barrier)
//   reserve   warp 0 reserves the tile's run in every partition: pass 1 from CTA-private cursors
//             (pre-computed from per-block histograms, no atomics), pass 2 with one global atomicAdd
//             per (tile, partition) whose latency hides behind the staging step
//   stage     tuples are reordered in shared memory so that each partition's run is contiguous
//   write-out the runs leave the SM as coalesced stores
// Two block barriers per tile. ncu on the first (TMA-ring) version of this kernel showed the
// shared-memory pipe, not HBM, as the limiter (l1tex 71-82 %), so this version keeps tuples in
// registers between load and staging instead of bouncing them through a shared-memory ring.
// ---------------------------------------------------------------------------------------------
constexpr int kScatterThreads = AQP_SCATTER_THREADS;
constexpr int kScatterWarps = kScatterThreads / 32;
constexpr int kScatterItems = kScatterTile / kScatterThreads;
static_assert(kScatterItems * kScatterThreads == kScatterTile, "tile shape");
constexpr int kBinsPerLane = kMaxFanout / 32;
// 228 KiB of shared memory per SM: staging buffer + per-warp scan copies + bookkeeping per CTA
constexpr size_t kScatterSmemApprox = (size_t) kScatterTile * 8 + (size_t) kScatterWarps * kMaxFanout * 4 + 6 * 1024;
constexpr int kScatterBlocksBySmem = (int) (228 * 1024 / kScatterSmemApprox);
constexpr int kScatterBlocksByThreads = 2048 / kScatterThreads;
constexpr int kScatterBlocksByRegs = 65536 / (kScatterThreads * 64);   // kernel is held to 64 registers
constexpr int kScatterBlocksPerSM0 = kScatterBlocksBySmem < kScatterBlocksByThreads ? kScatterBlocksBySmem : kScatterBlocksByThreads;
constexpr int kScatterBlocksPerSM1 = kScatterBlocksPerSM0 < kScatterBlocksByRegs ? kScatterBlocksPerSM0 : kScatterBlocksByRegs;
constexpr int kScatterBlocksPerSM = kScatterBlocksPerSM1 < 1 ? 1 : kScatterBlocksPerSM1;

uint32_t pass1_blocks() { return (uint32_t) kNumSMs * kScatterBlocksPerSM; }

__global__ void __launch_bounds__(kScatterThreads, kScatterBlocksPerSM)
radix_scatter_kernel(const uint2 *__restrict__ in, uint2 *__restrict__ out,
                     const uint32_t *__restrict__ seg_off, const uint32_t *__restrict__ seg_tile_start,
                     uint32_t nseg, uint32_t shift, uint32_t bits, uint32_t *__restrict__ cursors,
                     const uint32_t *__restrict__ block_base, uint32_t tiles_per_block) {
    extern __shared__ __align__(16) uint2 stage[];         // kScatterTile tuples
    __shared__ uint32_t cnt[2][kMaxFanout];                // tile histograms, double-buffered
    __shared__ uint32_t wbase[kScatterWarps][kMaxFanout];  // per-warp copy of the exclusive scan
    __shared__ uint32_t gdst[kMaxFanout];    // global index of stage slot s of partition d is gdst[d] + s
    __shared__ uint32_t scur[kMaxFanout];    // CTA-private write cursors (pass 1)
    __shared__ uint32_t s_tstart[kMaxFanout + 1];
    __shared__ uint32_t s_soff[kMaxFanout + 1];

    const uint32_t fan = 1u << bits, mask = fan - 1;
    const uint32_t per = (fan + 31) / 32;   // consecutive bins per lane in the warp scan (<= 8)
    const bool priv = block_base != nullptr;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i <= nseg; i += kScatterThreads) {
        s_tstart[i] = seg_tile_start[i];
        s_soff[i] = seg_off[i];
    }
    for (uint32_t i = threadIdx.x; i < fan; i += kScatterThreads) {
        cnt[0][i] = 0;
        cnt[1][i] = 0;
        if (priv) scur[i] = block_base[(size_t) blockIdx.x * fan + i];
    }
    __syncthreads();
    const uint32_t ntiles = s_tstart[nseg];
    // tiles of this CTA: a contiguous range (private cursors) or a grid-strided sequence (atomic reservation)
    uint32_t first, step, n_my;
    if (priv) {
        first = blockIdx.x * tiles_per_block;
        step = 1;
        n_my = first < ntiles ? min(tiles_per_block, ntiles - first) : 0;
    } else {
        first = blockIdx.x;
        step = gridDim.x;
        n_my = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    }
    // tile -> (segment, [begin, end))
    auto tile_range = [&](uint32_t tile, uint32_t &seg, uint32_t &begin, uint32_t &end) {
        uint32_t lo = 0, hi = nseg;   // last s with s_tstart[s] <= tile
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_tstart[mid] <= tile) lo = mid; else hi = mid;
        }
        seg = lo;
        begin = s_soff[lo] + (tile - s_tstart[lo]) * kScatterTile;
        end = min(begin + (uint32_t) kScatterTile, s_soff[lo + 1]);
    };

    uint2 vn[kScatterItems];   // next tile, in flight
    uint32_t seg_n = 0, begin_n = 0, end_n = 0;
    if (n_my > 0) {
        tile_range(first, seg_n, begin_n, end_n);
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = begin_n + j * kScatterThreads + threadIdx.x;
            if (k < end_n) vn[j] = ld_stream_v2(in + k);
        }
    }

    for (uint32_t i = 0; i < n_my; ++i) {
        const uint32_t seg = seg_n, ntile = end_n - begin_n;
        uint32_t *const cn = cnt[i & 1];
        uint2 v[kScatterItems];
        uint32_t rank[kScatterItems];
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) v[j] = vn[j];
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            if (k < ntile) rank[j] = atomicAdd(&cn[(v[j].x >> shift) & mask], 1u);
        }
        if (i + 1 < n_my) {   // put the next tile in flight before anything else happens
            tile_range(first + (i + 1) * step, seg_n, begin_n, end_n);
#pragma unroll
            for (int j = 0; j < kScatterItems; ++j) {
                uint32_t k = begin_n + j * kScatterThreads + threadIdx.x;
                if (k < end_n) vn[j] = ld_stream_v2(in + k);
            }
        }
        __syncthreads();   // (1) tile histogram complete; previous tile fully written out

        // every warp: exclusive scan of the tile histogram into its own copy
        uint32_t c[kBinsPerLane], my_g[kBinsPerLane], sum = 0;
#pragma unroll
        for (int k = 0; k < kBinsPerLane; ++k) {
            uint32_t d = lane * per + k;
            c[k] = (k < (int) per && d < fan) ? cn[d] : 0;
            sum += c[k];
        }
        uint32_t run = warp_incl_scan(sum) - sum;
#pragma unroll
        for (int k = 0; k < kBinsPerLane; ++k) {
            uint32_t d = lane * per + k;
            if (k < (int) per && d < fan) wbase[warp][d] = run;
            run += c[k];
        }
        if (warp == 0) {   // reserve the runs
#pragma unroll
            for (int k = 0; k < kBinsPerLane; ++k) {
                uint32_t d = lane * per + k;
                if (k < (int) per && d < fan) {
                    if (priv) {
                        my_g[k] = scur[d];
                        scur[d] = my_g[k] + c[k];
                    } else {
                        my_g[k] = c[k] ? atomicAdd(&cursors[(seg << bits) + d], c[k]) : 0u;
                    }
                }
            }
        }
        __syncwarp();

#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            if (k < ntile) stage[wbase[warp][(v[j].x >> shift) & mask] + rank[j]] = v[j];
        }
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < kBinsPerLane; ++k) {
                uint32_t d = lane * per + k;
                if (k < (int) per && d < fan) gdst[d] = my_g[k] - wbase[0][d];
            }
        }
        __syncthreads();   // (2) tile reordered, destinations known

        if (threadIdx.x < fan) cn[threadIdx.x] = 0;   // this buffer is used again two tiles from now
        for (uint32_t s = threadIdx.x; s < ntile; s += kScatterThreads) {
            uint2 t = stage[s];
            out[gdst[(t.x >> shift) & mask] + s] = t;
        }
    }
}

int radix_scatter_launch(const row_t *d_in, row_t *d_out, const uint32_t *d_seg_off,
                         const uint32_t *d_seg_tile_start, uint32_t nseg, uint64_t n_total, uint32_t shift,
                         uint32_t bits, uint32_t *d_cursors, const uint32_t *d_block_base, uint32_t nblocks,
                         uint32_t tiles_per_block, cudaStream_t st) {
    if (bits > (uint32_t) kMaxFanoutBits || nseg > (uint32_t) kMaxFanout) {
        set_error("radix_scatter: fan-out too large");
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(d_in) & 7u) != 0) {
        set_error("radix_scatter: relation must be 8-byte aligned");
        return -1;
    }
    if (n_total == 0) return 0;
    uint32_t grid;
    if (d_block_base) {
        grid = nblocks;
    } else {
        uint64_t max_tiles = n_total / kScatterTile + nseg;
        uint64_t g = (uint64_t) kNumSMs * kScatterBlocksPerSM;
        grid = (uint32_t) (max_tiles < g ? max_tiles : g);
    }
    static bool attr_set = false;
    if (!attr_set) {
        AQP_CUDA_OK(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) (kScatterTile * sizeof(uint2))));
        attr_set = true;
    }
    radix_scatter_kernel<<<grid, kScatterThreads, kScatterTile * sizeof(uint2), st>>>(
        reinterpret_cast<const uint2 *>(d_in), reinterpret_cast<uint2 *>(d_out), d_seg_off, d_seg_tile_start, nseg,
        shift, bits, d_cursors, d_block_base, tiles_per_block);
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
