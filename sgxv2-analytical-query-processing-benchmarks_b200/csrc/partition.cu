// partition.cu — radix partitioning of 8-byte {key,payload} tuples (sm_100a).
//
// Replaces the partitioning loops of Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp:
//   partition_hist[_unrolled]      :617-654  -> radix_hist_kernel   (ONE read of each relation yields
//                                               the histogram over ALL radix bits of both passes)
//   prefix sums                    :886-915, -> plan_offsets_kernel (partition boundaries of both
//                                   :739-746     passes from that one histogram)
//   partition_copy[_unrolled]      :659-697  -> radix_scatter_bins_kernel (pass 1: whole relation, one
//   radix_cluster/serial_radix_..  :715-841     segment; pass 2: every pass-1 partition is a segment)
//
// Digit function is the reference's: (key & MASK) >> R on raw key bits, pass 1 on bits [0,b1),
// pass 2 on bits [b1,b1+b2) (radix_join.cpp:47,:1118-1119,:1262).
//
// Scatter design (radix_scatter_bins_kernel, the default): a CTA bulk-loads a tile of 4096 tuples with TMA, ranks every
// tuple inside its partition with a shared-memory atomic counter, stages it in that partition's FIXED bin of the
// staging buffer (slot = (digit << log2 cap) + rank: no per-tile scan, no offset look-ups) and drains every bin with
// one TMA bulk store per (tile, partition); the destination comes from CTA-private cursors in pass 1 (pre-computed
// from per-CTA histogram rows, no atomics) and from one global atomic per (tile, partition) in pass 2. This is
// software write-combining - the analogue of the reference's dormant SWWC path (:1013-1053). A tile in which a bin
// would overflow (skew) takes a compacting path in the same kernel. radix_scatter_kernel is the older staged kernel
// (scan + look-ups, SM-store or bulk write-out) kept for unaligned destinations and A/B runs;
// radix_scatter_peer_kernel is the multi-GPU variant whose bins are rings across tiles so that only whole 128-byte
// lines cross NVLink. Order inside a partition is unspecified - exactly like the reference, where it depends on
// thread interleaving - and never affects join results.
#include "common.cuh"
#include "join_internal.cuh"
#include "block_scan.cuh"
#include <cstdlib>
#include <cstring>

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
template <bool kRot>
__global__ void __launch_bounds__(kHistThreads)
radix_hist_smem_kernel(const uint2 *__restrict__ in, uint64_t n, DigitFn digit, uint32_t bits,
                       uint32_t *__restrict__ ghist, uint64_t chunk, uint32_t bits1,
                       uint32_t *__restrict__ block_hist) {
    extern __shared__ uint32_t sh[];
    const uint32_t fan = 1u << bits;
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
            if (i < cend) atomicAdd(&sh[digit.template get<kRot>(v[j].x)], 1u);
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
radix_hist_global_kernel(const uint2 *__restrict__ in, uint64_t n, DigitFn digit, uint32_t *__restrict__ ghist) {
    uint64_t stride = (uint64_t) gridDim.x * kHistThreads;
    for (uint64_t i = (uint64_t) blockIdx.x * kHistThreads + threadIdx.x; i < n; i += stride)
        atomicAdd(&ghist[digit(ld_stream_v2(in + i).x)], 1u);
}

// nblocks == 0: plain histogram with a grid chosen here. nblocks > 0: exactly nblocks CTAs, CTA b
// covering tuples [b*chunk, (b+1)*chunk), and block_hist[nblocks][2^bits1] rows are written.
int radix_hist_device(const row_t *d_in, uint64_t n, DigitFn digit, uint32_t bits, uint32_t *d_hist,
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
        static unsigned attr_set = ~0u;   // device epoch the opt-in was made for (b200_shutdown + b200_init(other device) re-arms it)
        if (attr_set != g_device_epoch) {
            AQP_CUDA_OK(cudaFuncSetAttribute(radix_hist_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) (sizeof(uint32_t) << kMaxSmemHistBits)));
            AQP_CUDA_OK(cudaFuncSetAttribute(radix_hist_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) (sizeof(uint32_t) << kMaxSmemHistBits)));
            attr_set = g_device_epoch;
        }
        uint32_t grid = nblocks;
        if (!grid) {
            const uint64_t tile = (uint64_t) kHistThreads * kHistUnroll;
            uint64_t tiles = (n + tile - 1) / tile;
            int per_sm = bits <= 13 ? 3 : (bits == 14 ? 2 : 1);
            grid = (uint32_t) (tiles < (uint64_t) kNumSMs * per_sm ? tiles : (uint64_t) kNumSMs * per_sm);
            chunk = (tiles + grid - 1) / grid * tile;
        }
        if (digit.rot)
            radix_hist_smem_kernel<true><<<grid, kHistThreads, smem, st>>>(in, n, digit, bits, d_hist, chunk, bits1,
                                                                          nblocks ? d_block_hist : nullptr);
        else
            radix_hist_smem_kernel<false><<<grid, kHistThreads, smem, st>>>(in, n, digit, bits, d_hist, chunk, bits1,
                                                                           nblocks ? d_block_hist : nullptr);
    } else {
        radix_hist_global_kernel<<<kNumSMs * 4, kHistThreads, 0, st>>>(in, n, digit, d_hist);
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

// CTA-private pass-1 cursors: block_base[b][p] = start[p] + sum of block_hist[b'][p] over b' < b. One WARP per
// partition column p: the lanes take blocks lane, lane + 32, ... (128 loads of a column in flight at once) and a
// warp scan with a carry turns them into prefixes; columns are spread over many small CTAs. The first version
// walked each column with one thread and nblocks dependent iterations inside a single CTA: 57-93 us under ncu
// on the critical path of every join (and four times per multi-GPU join).
// Returns the column total to every lane.
__device__ __forceinline__ uint32_t block_base_column(const uint32_t *__restrict__ block_hist, uint32_t start,
                                                      uint32_t p, uint32_t fan, uint32_t nblocks,
                                                      uint32_t *__restrict__ block_base) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t run = start;
    for (uint32_t b0 = 0; b0 < nblocks; b0 += 128) {
        uint32_t c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = b0 + k * 32 + lane;
            c[k] = b < nblocks ? block_hist[(size_t) b * fan + p] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = b0 + k * 32 + lane;
            const uint32_t incl = warp_incl_scan(c[k]);
            if (block_base && b < nblocks) block_base[(size_t) b * fan + p] = run + incl - c[k];
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    return run - start;
}

constexpr int kBlockBaseWarps = 4;   // columns per CTA of block_base_kernel

// CTA-private cursors of `nrel` relations in one launch (blockIdx.y = relation). part_start[p] is where the
// relation's partition p begins in its destination: the local pass-1 output (single GPU), the send buffer, or —
// fused exchange — the owner's receive buffer (computed by the host from the all-gathered counts); null = 0.
// counts[p] (optional) receives the relation's size of partition p, seg1 (optional) the single-segment tile table.
struct BlockBaseRel {
    const uint32_t *block_hist;
    const uint32_t *part_start;
    uint32_t *block_base;
    uint32_t *counts;
    uint32_t *seg1;
    uint32_t n;
};
struct BlockBaseArgs {
    BlockBaseRel rel[2];
};
__global__ void __launch_bounds__(kBlockBaseWarps * 32)
block_base_kernel(BlockBaseArgs a, uint32_t fan, uint32_t nblocks) {
    const BlockBaseRel &r = a.rel[blockIdx.y];
    const uint32_t p = blockIdx.x * kBlockBaseWarps + (threadIdx.x >> 5);
    if (p < fan) {
        const uint32_t start = r.part_start ? r.part_start[p] : 0u;
        const uint32_t tot = block_base_column(r.block_hist, start, p, fan, nblocks, r.block_base);
        if (r.counts && (threadIdx.x & 31u) == 0) r.counts[p] = tot;
    }
    if (r.seg1 && blockIdx.x == 0 && threadIdx.x == 0) {
        r.seg1[0] = 0;
        r.seg1[1] = r.n;
        r.seg1[2] = 0;
        r.seg1[3] = (r.n + kScatterTile - 1) / kScatterTile;
    }
}

static int block_base_launch(const BlockBaseArgs &a, uint32_t nrel, uint32_t fan, uint32_t nblocks, cudaStream_t st) {
    dim3 grid((fan + kBlockBaseWarps - 1) / kBlockBaseWarps, nrel);
    block_base_kernel<<<grid, kBlockBaseWarps * 32, 0, st>>>(a, fan, nblocks);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int block_base_device(const uint32_t *d_block_hist, const uint32_t *d_part_start, uint32_t fan, uint32_t nblocks,
                      uint32_t *d_block_base, uint32_t *d_counts, uint32_t *d_seg1, uint32_t n, cudaStream_t st) {
    BlockBaseArgs a{};
    a.rel[0] = BlockBaseRel{d_block_hist, d_part_start, d_block_base, d_counts, d_seg1, n};
    return block_base_launch(a, 1, fan, nblocks, st);
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
}

int plan_offsets_device(const PlanArgs &a, cudaStream_t st) {
    plan_offsets_kernel<<<2, kScanBlock, 0, st>>>(a);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    if (a.rel[0].block_hist) {   // CTA-private pass-1 cursors from the per-CTA histogram rows
        BlockBaseArgs b{};
        for (int i = 0; i < 2; ++i)
            b.rel[i] = BlockBaseRel{a.rel[i].block_hist, a.rel[i].cursor1, a.rel[i].block_base, nullptr, nullptr, 0};
        return block_base_launch(b, 2, 1u << a.bits1, a.nblocks1, st);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// plans for the sharded (multi-GPU) join
// ---------------------------------------------------------------------------------------------
// pass 1 on one rank: partition sizes of the pass-1 digit from the full-width histogram, their
// exclusive prefix (= layout of the send buffer) and the CTA-private cursors
__global__ void __launch_bounds__(kScanBlock)
plan_pass1_kernel(const uint32_t *__restrict__ hist, uint32_t bits1, uint32_t bits2, uint32_t *__restrict__ part1_off,
                  uint32_t *__restrict__ seg1, const uint32_t *__restrict__ block_hist,
                  uint32_t *__restrict__ block_base, uint32_t nblocks) {
    const uint32_t F1 = 1u << bits1, F2 = 1u << bits2;
    uint32_t total = block_exclusive_scan(
        F1,
        [&](uint32_t p1) {
            uint32_t c = 0;
            for (uint32_t p2 = 0; p2 < F2; ++p2) c += hist[p1 + (p2 << bits1)];
            return c;
        },
        [&](uint32_t p1, uint32_t v) { part1_off[p1] = v; });
    if (threadIdx.x == 0) {
        part1_off[F1] = total;
        seg1[0] = 0;
        seg1[1] = total;
        seg1[2] = 0;
        seg1[3] = (total + kScatterTile - 1) / kScatterTile;
    }
}

int plan_pass1_device(const uint32_t *d_hist, uint32_t bits1, uint32_t bits2, uint32_t *d_part1_off, uint32_t *d_seg1,
                      const uint32_t *d_block_hist, uint32_t *d_block_base, uint32_t nblocks, cudaStream_t st) {
    plan_pass1_kernel<<<1, kScanBlock, 0, st>>>(d_hist, bits1, bits2, d_part1_off, d_seg1, d_block_hist, d_block_base,
                                                nblocks);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return block_base_device(d_block_hist, d_part1_off, 1u << bits1, nblocks, d_block_base, nullptr, nullptr, 0, st);
}

// Everything a rank derives from the two sizing collectives of the fused exchange, in ONE launch (the host side
// used ~45 small tensor kernels for this, 0.2 ms of launch latency per join):
//   counts_all[s][rel][p]  tuples source rank s holds of routed pass-1 partition p of relation rel (all-gathered)
//   hist_global[rel][D]    full-width histograms summed over ranks, D = p1' | (p2 << bits1)
// ->
//   seg_off[rel][world*per + 1]   starts of the segments this rank receives, ordered (source, local partition)
//   dest_off[rel][F1]             where THIS rank's segment of partition p starts inside its owner's buffer
//   hist_slice[rel][per << bits2] this rank's final partition sizes, order (local partition, p2)
//   host_vals[6]                  largest receive size of any rank for R, for S; this rank's receive sizes R, S;
//                                 tuples this rank keeps (R, S)
__global__ void __launch_bounds__(256)
exchange_plan_kernel(const uint32_t *__restrict__ counts_all, uint32_t world, uint32_t rank, uint32_t bits1,
                     uint32_t bits2, const uint32_t *__restrict__ hist_global, uint32_t *__restrict__ seg_off,
                     uint32_t *__restrict__ dest_off, uint32_t *__restrict__ hist_slice,
                     unsigned long long *__restrict__ host_vals) {
    __shared__ uint32_t tot[2][8][8];   // [rel][source][owner]
    const uint32_t F1 = 1u << bits1, F2 = 1u << bits2, P = F1 << bits2, per = F1 / world, nseg = world * per;
    auto cnt = [&](uint32_t s, uint32_t rel, uint32_t p) { return counts_all[((size_t) s * 2 + rel) * F1 + p]; };
    for (uint32_t t = threadIdx.x; t < 2 * world * world; t += blockDim.x) {
        const uint32_t rel = t / (world * world), s = (t / world) % world, g = t % world;
        uint32_t a = 0;
        for (uint32_t j = 0; j < per; ++j) a += cnt(s, rel, g * per + j);
        tot[rel][s][g] = a;
    }
    __syncthreads();
    if (threadIdx.x < 2) {   // sizes for the host
        const uint32_t rel = threadIdx.x;
        unsigned long long worst = 0, mine = 0;
        for (uint32_t g = 0; g < world; ++g) {
            unsigned long long r = 0;
            for (uint32_t s = 0; s < world; ++s) r += tot[rel][s][g];
            worst = r > worst ? r : worst;
            if (g == rank) mine = r;
        }
        host_vals[rel] = worst;
        host_vals[2 + rel] = mine;
        host_vals[4 + rel] = tot[rel][rank][rank];
    }
    for (uint32_t t = threadIdx.x; t < 2 * world; t += blockDim.x) {   // dest_off: one thread per (relation, owner)
        const uint32_t rel = t / world, g = t % world;
        uint32_t run = 0;
        for (uint32_t s = 0; s < rank; ++s) run += tot[rel][s][g];
        for (uint32_t j = 0; j < per; ++j) {
            dest_off[rel * F1 + g * per + j] = run;
            run += cnt(rank, rel, g * per + j);
        }
    }
    // seg_off: exclusive prefix over (source, local partition); one warp per relation
    if (threadIdx.x < 64) {
        const uint32_t rel = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        uint32_t run = 0;
        for (uint32_t i0 = 0; i0 < nseg; i0 += 32) {
            const uint32_t i = i0 + lane;
            const uint32_t v = i < nseg ? cnt(i / per, rel, rank * per + i % per) : 0u;
            const uint32_t incl = warp_incl_scan(v);
            if (i < nseg) seg_off[rel * (nseg + 1) + i] = run + incl - v;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) seg_off[rel * (nseg + 1) + nseg] = run;
    }
    for (uint32_t t = threadIdx.x; t < 2 * (per << bits2); t += blockDim.x) {
        const uint32_t rel = t / (per << bits2), f = t % (per << bits2), j = f >> bits2, p2 = f & (F2 - 1);
        hist_slice[t] = hist_global[(size_t) rel * P + ((size_t) p2 << bits1) + rank * per + j];
    }
}

int exchange_plan_device(const uint32_t *d_counts_all, uint32_t world, uint32_t rank, uint32_t bits1, uint32_t bits2,
                         const uint32_t *d_hist_global, uint32_t *d_seg_off, uint32_t *d_dest_off, uint32_t *d_hist_slice,
                         unsigned long long *d_host_vals, cudaStream_t st) {
    exchange_plan_kernel<<<1, 256, 0, st>>>(d_counts_all, world, rank, bits1, bits2, d_hist_global, d_seg_off, d_dest_off,
                                            d_hist_slice, d_host_vals);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// after the exchange: final partition boundaries from the (already globally reduced) histogram slice
// of this rank in final order, and the tile table of the received segments. blockIdx.x = relation.
__global__ void __launch_bounds__(kScanBlock) plan_shard_kernel(ShardPlanArgs a) {
    const ShardRelPlan &r = a.rel[blockIdx.x];
    uint32_t total = block_exclusive_scan(
        a.nparts, [&](uint32_t f) { return r.hist[f]; },
        [&](uint32_t f, uint32_t v) {
            r.part_off[f] = v;
            r.cursor2[f] = v;
        });
    if (threadIdx.x == 0) r.part_off[a.nparts] = total;
    __syncthreads();
    uint32_t tiles = block_exclusive_scan(
        a.nseg,
        [&](uint32_t s) {
            if (a.seg_group && a.seg_group[s] == kGapSegment) return 0u;   // unused tail of a source's region
            uint32_t len = r.seg_off[s + 1] - r.seg_off[s];
            return (len + kScatterTile - 1) / kScatterTile;
        },
        [&](uint32_t s, uint32_t v) { r.seg_tile_start[s] = v; });
    if (threadIdx.x == 0) r.seg_tile_start[a.nseg] = tiles;
}

// ---- region layout of the fused exchange (multi-GPU host, csrc/mg.cu) ------------------------------------------
// Every owner's receive buffer is cut into one fixed region per source rank (cap tuples, sized for the worst case:
// a source's whole shard). A source packs its partitions of that owner tightly at the start of its region, so it can
// compute every destination from its OWN counts - no collective in front of the scatter. The receiver learns the
// segment boundaries from the all-gathered counts, which travel while the scatter runs.
//   dest_off[rel][p] = rank * cap_rel + (tuples of this rank in the partitions of p's owner that precede p)
__global__ void __launch_bounds__(256)
region_dest_kernel(const uint32_t *__restrict__ counts1, uint32_t world, uint32_t rank, uint32_t bits1, uint32_t cap_r,
                   uint32_t cap_s, uint32_t *__restrict__ dest_off, unsigned long long *__restrict__ kept) {
    const uint32_t F1 = 1u << bits1, per = F1 / world;
    for (uint32_t t = threadIdx.x; t < 2 * world; t += blockDim.x) {   // one thread per (relation, owner)
        const uint32_t rel = t / world, g = t % world;
        uint32_t run = rank * (rel ? cap_s : cap_r);
        for (uint32_t j = 0; j < per; ++j) {
            dest_off[rel * F1 + g * per + j] = run;
            run += counts1[rel * F1 + g * per + j];
        }
        if (g == rank && kept) kept[rel] = run - rank * (rel ? cap_s : cap_r);   // tuples that stay on this GPU
    }
}
int region_dest_device(const uint32_t *d_counts1, uint32_t world, uint32_t rank, uint32_t bits1, uint64_t cap_r, uint64_t cap_s,
                       uint32_t *d_dest_off, unsigned long long *d_kept, cudaStream_t st) {
    region_dest_kernel<<<1, 256, 0, st>>>(d_counts1, world, rank, bits1, (uint32_t) cap_r, (uint32_t) cap_s, d_dest_off, d_kept);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// receiver side: counts_all[s][rel][p] (all-gathered), hist_global[rel][D] (all-reduced) ->
//   seg_off[rel][nseg + 1], nseg = world * (per + 1): for every source its per segments, then the gap up to the next
//   region; seg_group[nseg] = local partition of a segment or kGapSegment; hist_slice[rel][per << bits2]
__global__ void __launch_bounds__(256)
region_plan_kernel(const uint32_t *__restrict__ counts_all, uint32_t world, uint32_t rank, uint32_t bits1, uint32_t bits2,
                   const uint32_t *__restrict__ hist_global, uint32_t cap_r, uint32_t cap_s, uint32_t *__restrict__ seg_off,
                   uint32_t *__restrict__ seg_group, uint32_t *__restrict__ hist_slice) {
    const uint32_t F1 = 1u << bits1, F2 = 1u << bits2, P = F1 << bits2, per = F1 / world, nseg = world * (per + 1);
    for (uint32_t t = threadIdx.x; t < 2 * world; t += blockDim.x) {   // one thread per (relation, source)
        const uint32_t rel = t / world, s = t % world, cap = rel ? cap_s : cap_r;
        uint32_t run = s * cap;
        uint32_t *so = seg_off + rel * (nseg + 1) + s * (per + 1);
        for (uint32_t j = 0; j < per; ++j) {
            so[j] = run;
            run += counts_all[((size_t) s * 2 + rel) * F1 + rank * per + j];
        }
        so[per] = run;                                   // the gap: [end of the source's data, next region)
        if (s == world - 1) so[per + 1] = world * cap;   // = seg_off[rel][nseg]
    }
    for (uint32_t i = threadIdx.x; i < nseg; i += blockDim.x) seg_group[i] = i % (per + 1) == per ? kGapSegment : i % (per + 1);
    for (uint32_t t = threadIdx.x; t < 2 * (per << bits2); t += blockDim.x) {
        const uint32_t rel = t / (per << bits2), f = t % (per << bits2), j = f >> bits2, p2 = f & (F2 - 1);
        hist_slice[t] = hist_global[(size_t) rel * P + ((size_t) p2 << bits1) + rank * per + j];
    }
}
int region_plan_device(const uint32_t *d_counts_all, uint32_t world, uint32_t rank, uint32_t bits1, uint32_t bits2,
                       const uint32_t *d_hist_global, uint64_t cap_r, uint64_t cap_s, uint32_t *d_seg_off,
                       uint32_t *d_seg_group, uint32_t *d_hist_slice, cudaStream_t st) {
    region_plan_kernel<<<1, 256, 0, st>>>(d_counts_all, world, rank, bits1, bits2, d_hist_global, (uint32_t) cap_r,
                                          (uint32_t) cap_s, d_seg_off, d_seg_group, d_hist_slice);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- histogram-free plan: regions of fixed capacity (RegionArgs, join_internal.cuh). blockIdx.x = relation. ----
__global__ void __launch_bounds__(kScanBlock) region_init_kernel(RegionArgs a) {
    const RegionRel &r = a.rel[blockIdx.x];
    const uint32_t F1 = 1u << a.bits1, P = F1 << a.bits2;
    for (uint32_t p = threadIdx.x; p < F1; p += kScanBlock) {
        r.cur1[p] = p * r.cap1;
        if (blockIdx.x == 0) {
            a.seg_group[2 * p] = p;
            a.seg_group[2 * p + 1] = kGapSegment;
        }
    }
    for (uint32_t f = threadIdx.x; f < P; f += kScanBlock) {
        r.cur2[f] = f * r.cap2;
    }
    if (threadIdx.x == 0) {
        r.seg1[0] = 0;
        r.seg1[1] = r.n;
        r.seg1[2] = 0;
        r.seg1[3] = (r.n + kScatterTile - 1) / kScatterTile;
    }
}
// pass-2 input: partition p of pass 1 is the segment [p cap1, p cap1 + what pass 1 put there), followed by a gap
__global__ void __launch_bounds__(kScanBlock) region_plan2_kernel(RegionArgs a) {
    const RegionRel &r = a.rel[blockIdx.x];
    const uint32_t F1 = 1u << a.bits1;
    auto filled = [&](uint32_t p) { return min(r.cur1[p], (p + 1) * r.cap1) - p * r.cap1; };
    for (uint32_t p = threadIdx.x; p < F1; p += kScanBlock) {
        r.seg_off[2 * p] = p * r.cap1;
        r.seg_off[2 * p + 1] = p * r.cap1 + filled(p);
    }
    if (threadIdx.x == 0) r.seg_off[2 * F1] = F1 * r.cap1;
    uint32_t tiles = block_exclusive_scan(
        2 * F1, [&](uint32_t s) { return (s & 1u) ? 0u : (filled(s >> 1) + kScatterTile - 1) / kScatterTile; },
        [&](uint32_t s, uint32_t v) { r.seg_tile[s] = v; });
    if (threadIdx.x == 0) r.seg_tile[2 * F1] = tiles;
}
__global__ void __launch_bounds__(kScanBlock) region_plan3_kernel(RegionArgs a) {
    const RegionRel &r = a.rel[blockIdx.x];
    const uint32_t P = 1u << (a.bits1 + a.bits2);
    for (uint32_t f = threadIdx.x; f < P; f += kScanBlock) {
        r.beg[f] = f * r.cap2;
        r.end[f] = min(r.cur2[f], (f + 1) * r.cap2);
    }
}
// one thread per tuple of the sampled lines (16 tuples per line) of BOTH relations in one launch, four independent
// loads in flight per thread; fire-and-forget global atomics - a few million per join
struct SampleRel {
    const uint2 *in;
    uint32_t n, line_stride, nthreads;   // nthreads = sampled lines * 16
    uint32_t *hist;
};
__global__ void __launch_bounds__(256) region_sample_kernel(SampleRel a, SampleRel b, uint32_t mask) {
    const uint32_t total = a.nthreads + b.nthreads, step = gridDim.x * blockDim.x;
    for (uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += 4 * step) {
        uint32_t key[4];
        uint32_t *dst[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t t = t0 + k * step;
            dst[k] = nullptr;
            if (t < total) {
                const SampleRel &r = t < a.nthreads ? a : b;
                const uint32_t u = t < a.nthreads ? t : t - a.nthreads;
                const uint32_t i = (u >> 4) * r.line_stride * 16 + (u & 15u);
                if (i < r.n) {
                    key[k] = r.in[i].x;
                    dst[k] = r.hist;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (dst[k]) atomicAdd(&dst[k][key[k] & mask], 1u);
    }
}
// blockIdx.x = relation. Full digit d = key & (P - 1): pass-1 partition d & (F1 - 1), final partition = the digit itself.
__global__ void __launch_bounds__(kScanBlock)
region_verdict_kernel(const uint32_t *__restrict__ hist_R, const uint32_t *__restrict__ hist_S, uint32_t bits1, uint32_t bits2,
                      uint32_t *__restrict__ out) {
    __shared__ uint32_t part1[kMaxFanout];
    __shared__ uint32_t s_total, s_max1, s_max2;
    const uint32_t *hist = blockIdx.x ? hist_S : hist_R;
    const uint32_t F1 = 1u << bits1, P = F1 << bits2;
    for (uint32_t p = threadIdx.x; p < F1; p += kScanBlock) part1[p] = 0;
    if (threadIdx.x == 0) s_total = s_max1 = s_max2 = 0;
    __syncthreads();
    uint32_t total = 0, mx = 0;
    for (uint32_t d = threadIdx.x; d < P; d += kScanBlock) {
        const uint32_t c = hist[d];
        total += c;
        mx = max(mx, c);
        if (c) atomicAdd(&part1[d & (F1 - 1)], c);
    }
    atomicAdd(&s_total, total);
    atomicMax(&s_max2, mx);
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < F1; p += kScanBlock) atomicMax(&s_max1, part1[p]);
    __syncthreads();
    if (threadIdx.x == 0) {
        out[blockIdx.x * 3 + 0] = s_total;
        out[blockIdx.x * 3 + 1] = s_max1;
        out[blockIdx.x * 3 + 2] = s_max2;
    }
}
int region_sample_device(const row_t *d_R, uint64_t nR, uint32_t stride_R, uint32_t *d_hist_R, const row_t *d_S, uint64_t nS,
                         uint32_t stride_S, uint32_t *d_hist_S, uint32_t bits, cudaStream_t st) {
    auto rel = [](const row_t *d, uint64_t n, uint32_t stride, uint32_t *hist) {
        const uint64_t lines = (n + 15) / 16;
        return SampleRel{reinterpret_cast<const uint2 *>(d), (uint32_t) n, stride, (uint32_t) ((lines + stride - 1) / stride * 16), hist};
    };
    const SampleRel a = rel(d_R, nR, stride_R, d_hist_R), b = rel(d_S, nS, stride_S, d_hist_S);
    const uint64_t threads = (uint64_t) a.nthreads + b.nthreads;
    if (threads == 0) return 0;
    const uint64_t want = (threads + 4 * 256 - 1) / (4 * 256), cap = (uint64_t) kNumSMs * 8;
    region_sample_kernel<<<(unsigned) (want < cap ? want : cap), 256, 0, st>>>(a, b, (1u << bits) - 1);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}
int region_verdict_device(const uint32_t *d_hist_R, const uint32_t *d_hist_S, uint32_t bits1, uint32_t bits2, uint32_t *d_out,
                          cudaStream_t st) {
    region_verdict_kernel<<<2, kScanBlock, 0, st>>>(d_hist_R, d_hist_S, bits1, bits2, d_out);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}
int region_init_device(const RegionArgs &a, cudaStream_t st) {
    region_init_kernel<<<2, kScanBlock, 0, st>>>(a);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}
int region_plan2_device(const RegionArgs &a, cudaStream_t st) {
    region_plan2_kernel<<<2, kScanBlock, 0, st>>>(a);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}
int region_plan3_device(const RegionArgs &a, cudaStream_t st) {
    region_plan3_kernel<<<2, kScanBlock, 0, st>>>(a);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int plan_shard_device(const ShardPlanArgs &a, cudaStream_t st) {
    plan_shard_kernel<<<2, kScanBlock, 0, st>>>(a);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// scatter
//
// One CTA streams its tiles through a two-deep TMA ring: an elected thread issues 1-D bulk copies
// (cp.async.bulk, completion on an mbarrier) for tile i+2 while the CTA works on tile i, so HBM reads
// never wait for compute. Per tile:
//   rank      every tuple takes a slot in its partition with a shared-memory atomicAdd
//   reserve   warp 0 scans the tile histogram and reserves the tile's run in every partition:
//             pass 1 from CTA-private cursors (pre-computed from per-block histograms, no atomics),
//             pass 2 with one global atomicAdd per (tile, partition)
//   stage     tuples are reordered in shared memory so that each partition's run is contiguous
//   write-out the runs leave the SM as coalesced stores
// ncu (profiles/) shows this kernel limited by the shared-memory/L1 pipe (l1tex 71-82 %), not by HBM:
// DRAM traffic equals the algorithmic 16 B/tuple. A variant that kept tuples in registers between an
// LDG prefetch and the staging step (no shared-memory ring) was 25 % slower: the LDG data crosses the
// same L1 data stage the ring reads do, while TMA writes into shared memory do not.
// ---------------------------------------------------------------------------------------------
constexpr int kScatterThreads = AQP_SCATTER_THREADS;
constexpr int kScatterItems = kScatterTile / kScatterThreads;
static_assert(kScatterItems * kScatterThreads == kScatterTile, "tile shape");
constexpr int kInBufTuples = kScatterTile + 2;   // +1 leading tuple when the tile starts on an odd index, +1 to round up
constexpr int kStageTuples = kScatterTile + 3 * kMaxFanout;   // bulk write-out: carried tuple + parity pad + round-up per run
constexpr size_t kScatterSmemBytes = (size_t) (2 * kInBufTuples + kStageTuples) * sizeof(uint2);
// 228 KiB of shared memory per SM; every CTA also pays ~6.2 KiB of static arrays + 1 KiB reserved
constexpr int kScatterBlocksPerSMRaw = (int) (228 * 1024 / (kScatterSmemBytes + 7424));
constexpr int kScatterBlocksPerSM = kScatterBlocksPerSMRaw < 1 ? 1 : (kScatterBlocksPerSMRaw * kScatterThreads > 2048 ? 2048 / kScatterThreads : kScatterBlocksPerSMRaw);

uint32_t pass1_blocks() { return (uint32_t) kNumSMs * kScatterBlocksPerSM; }

// 1-D bulk copy shared -> global through the TMA unit (src/dst 16-byte aligned, bytes % 16 == 0), tracked by the
// thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}

// kBulk (CTA-private cursors only): every partition's run of a tile leaves shared memory as ONE TMA bulk store
// issued by the partition's own thread, instead of the CTA's threads copying the staging buffer with 8-byte stores.
// The write-out then costs no LSU instructions and — more important — runs asynchronously under the next tile's
// rank/scan phases (the CTA only waits, before it overwrites the staging buffer, until the bulk group has READ it).
// Bulk stores need 16-byte aligned source, destination and size; runs are 8-byte granular, so
//   * a run is staged at a slot whose parity equals the parity of its destination index,
//   * an odd destination start is fixed once per (CTA, partition) with one scalar store of the first tuple,
//   * an odd run length keeps its last tuple in shared memory (carry[d]) as the first tuple of the partition's
//     next run; the carries are flushed with scalar stores when the CTA is done.
// bulkbench (profiles/r01_bulkbench.txt): 256-byte bulk stores sustain 5.9 TB/s into local HBM and 714 GB/s to an
// NVLink peer (the link rate), far above what the scatter needs.
template <bool kRot, bool kPeer, bool kBulk>
__global__ void __launch_bounds__(kScatterThreads, kScatterBlocksPerSM)
radix_scatter_kernel(const uint2 *__restrict__ in, uint2 *__restrict__ out,
                     const uint32_t *__restrict__ seg_off, const uint32_t *__restrict__ seg_tile_start,
                     const uint32_t *__restrict__ seg_group, uint32_t nseg, DigitFn digit, uint32_t bits,
                     uint32_t *__restrict__ cursors, const uint32_t *__restrict__ block_base,
                     uint32_t tiles_per_block, PeerTable peers) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2 *inbuf0 = reinterpret_cast<uint2 *>(smem_raw);
    uint2 *inbuf1 = inbuf0 + kInBufTuples;
    uint2 *stage = inbuf1 + kInBufTuples;
    __shared__ uint32_t cnt[kMaxFanout];     // tuples of this tile per partition
    __shared__ uint32_t lbase[kMaxFanout];   // start of the partition's run inside `stage`
    __shared__ uint32_t gdst[kMaxFanout];    // global index of stage slot s of partition d is gdst[d] + s
    __shared__ uint32_t scur[kMaxFanout];    // CTA-private write cursors (pass 1)
    __shared__ uint2 *s_peer[8];             // receive buffers of the owners (fused exchange)
    __shared__ uint2 carry[kBulk ? kMaxFanout : 1];        // kBulk: the odd tuple a run left behind
    __shared__ uint32_t runlen[kBulk ? kMaxFanout : 1];    // kBulk: tuples in the run incl. the carried one | carried << 31
    __shared__ uint32_t s_tstart[kMaxSegs + 1];
    __shared__ uint32_t s_soff[kMaxSegs + 1];
    __shared__ __align__(8) uint64_t mbar[2];

    const uint32_t fan = 1u << bits;
    const bool priv = block_base != nullptr;
    for (uint32_t i = threadIdx.x; i <= nseg; i += kScatterThreads) {
        s_tstart[i] = seg_tile_start[i];
        s_soff[i] = seg_off[i];
    }
    for (uint32_t i = threadIdx.x; i < fan; i += kScatterThreads) {
        cnt[i] = 0;
        if (priv) scur[i] = block_base[(size_t) blockIdx.x * fan + i];
        if (kBulk) runlen[i] = 0;
    }
    if (kPeer && threadIdx.x < 8) s_peer[threadIdx.x] = peers.base[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    auto issue = [&](uint32_t i) {   // one thread: bulk-load tile #i of this CTA into ring slot i & 1
        uint32_t seg, begin, end;
        tile_range(first + i * step, seg, begin, end);
        const uint2 *src = in + begin;
        uint32_t skew = (uint32_t) ((reinterpret_cast<uintptr_t>(src) >> 3) & 1u);   // 16-byte alignment of the source
        uint32_t bytes = ((end - begin + skew + 1) & ~1u) * (uint32_t) sizeof(uint2);
        uint64_t *bar = &mbar[i & 1];
        mbar_expect_tx(bar, bytes);
        tma_load_1d((i & 1) ? inbuf1 : inbuf0, src - skew, bytes, bar);
    };
    if (threadIdx.x == 0) {
        if (n_my > 0) issue(0);
        if (n_my > 1) issue(1);
    }

    for (uint32_t i = 0; i < n_my; ++i) {
        uint32_t seg, begin, end;
        tile_range(first + i * step, seg, begin, end);
        const uint32_t ntile = end - begin;
        const uint32_t group = seg_group ? seg_group[seg] : seg;
        const uint32_t skew = (uint32_t) ((reinterpret_cast<uintptr_t>(in + begin) >> 3) & 1u);
        const uint2 *buf = ((i & 1) ? inbuf1 : inbuf0) + skew;

        mbar_wait(&mbar[i & 1], (i >> 1) & 1);
        uint2 v[kScatterItems];
        uint32_t rank[kScatterItems];
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            if (k < ntile) v[j] = buf[k];
        }
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            if (k < ntile) rank[j] = atomicAdd(&cnt[digit.template get<kRot>(v[j].x)], 1u);
        }
        __syncthreads();   // (1) tile histogram complete; ring slot consumed; previous write-out finished

        if (threadIdx.x == 0 && i + 2 < n_my) issue(i + 2);

        // warp 0: exclusive scan of the tile histogram, then reserve the runs
        uint32_t my_g[kMaxFanout / 32], my_b[kMaxFanout / 32];
        if (threadIdx.x < 32) {
            const uint32_t per = (fan + 31) / 32;   // consecutive bins per lane (<= 8)
            uint32_t c[kMaxFanout / 32], sum = 0;
            if (kBulk) {
                // run d occupies an even-sized region of the staging buffer that starts on an even slot; inside it the
                // run starts at the parity of its destination index: [pad?][carried tuple?][new tuples by rank]
                uint32_t hc[kMaxFanout / 32], par[kMaxFanout / 32];
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    bool ok = k < (int) per && d < fan;
                    c[k] = ok ? cnt[d] : 0;
                    hc[k] = ok ? runlen[d] >> 31 : 0;
                    par[k] = ok ? (scur[d] & 1u) : 0;
                    if (ok) cnt[d] = 0;
                    sum += (c[k] + hc[k] + par[k] + 1u) & ~1u;
                }
                uint32_t run = warp_incl_scan(sum) - sum;
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    if (k < (int) per && d < fan) {
                        const uint32_t n = c[k] + hc[k];          // tuples of the run, carried one first
                        lbase[d] = run + par[k] + hc[k];           // where the new tuples go (rank 0)
                        gdst[d] = run + par[k];                    // where the run starts in the staging buffer
                        runlen[d] = n | (hc[k] << 31);
                    }
                    run += (c[k] + hc[k] + par[k] + 1u) & ~1u;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    bool ok = k < (int) per && d < fan;
                    c[k] = ok ? cnt[d] : 0;
                    if (ok) cnt[d] = 0;   // ready for the next tile
                    sum += c[k];
                }
                uint32_t run = warp_incl_scan(sum) - sum;
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    my_b[k] = run;
                    if (k < (int) per && d < fan) {
                        lbase[d] = run;
                        if (priv) {
                            my_g[k] = scur[d];
                            scur[d] = my_g[k] + c[k];
                        } else {
                            my_g[k] = c[k] ? atomicAdd(&cursors[(group << bits) + d], c[k]) : 0u;
                        }
                    }
                    run += c[k];
                }
            }
        }
        if (kBulk && threadIdx.x < fan) {
            // the previous tile's bulk stores must have read the staging buffer before it is overwritten
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();   // (2) lbase ready

#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            if (k < ntile) stage[lbase[digit.template get<kRot>(v[j].x)] + rank[j]] = v[j];
        }
        if (kBulk) {
            // the partition's own thread: carried tuple to the front of the run, then (after the barrier) head fix,
            // bulk store of the even body, new carry
            const uint32_t d = threadIdx.x;
            uint32_t n = 0, sb = 0;
            if (d < fan) {
                const uint32_t rl = runlen[d];
                n = rl & 0x7FFFFFFFu;
                sb = gdst[d];
                if (rl >> 31) stage[sb] = carry[d];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged tuples visible to the TMA unit
            __syncthreads();   // (3) tile reordered
            if (d < fan) {
                uint32_t left = 0;
                if (n) {
                    uint2 *dst = kPeer ? s_peer[d >> peers.per_shift] : out;
                    uint32_t g = scur[d];
                    if (g & 1u) {   // odd destination start: once per (CTA, partition)
                        dst[g] = stage[sb];
                        ++g;
                        ++sb;
                        --n;
                    }
                    const uint32_t body = n & ~1u;
                    if (body) tma_store_1d(dst + g, stage + sb, body * (uint32_t) sizeof(uint2));
                    left = n & 1u;
                    if (left) carry[d] = stage[sb + body];
                    scur[d] = g + body;
                }
                runlen[d] = left << 31;
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            if (threadIdx.x < 32) {
                const uint32_t per = (fan + 31) / 32;
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    if (k < (int) per && d < fan) gdst[d] = my_g[k] - my_b[k];
                }
            }
            __syncthreads();   // (3) tile reordered, destinations known

            for (uint32_t s = threadIdx.x; s < ntile; s += kScatterThreads) {
                uint2 t = stage[s];
                const uint32_t d = digit.template get<kRot>(t.x);
                if (kPeer)   // fused exchange: the run goes straight into its owner's receive buffer over NVLink
                    s_peer[d >> peers.per_shift][gdst[d] + s] = t;
                else
                    out[gdst[d] + s] = t;
            }
        }
    }
    if (kBulk) {
        // flush the carried tuples, and do not leave before the TMA unit has finished reading shared memory
        const uint32_t d = threadIdx.x;
        if (d < fan) {
            if (runlen[d] >> 31) {
                uint2 *dst = kPeer ? s_peer[d >> peers.per_shift] : out;
                dst[scur[d]] = carry[d];
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------------
// scatter, fixed-bin variant (the default)
//
// Same tile pipeline as radix_scatter_kernel, but the staging buffer is 2^bits fixed bins of cap = 8192 >> bits
// slots (64 KiB): a tuple's staging slot is (digit << log2 cap) + rank — pure arithmetic, so
//   * no per-tile exclusive scan (a serial warp-0 section between two barriers in the staged kernel),
//   * no lbase[] look-up per tuple and no gdst[] look-up + digit recomputation per written tuple
//     (two random shared-memory loads out of ~24 wavefronts per 32 tuples on the pipe that bounds the kernel),
//   * every partition's own thread reserves its run (pass 2: 2^bits global atomics in parallel instead of a
//     warp-0 loop) and issues ONE TMA bulk store for it, in BOTH passes: an odd destination start is fixed by
//     storing the run's LAST tuple there (order inside a partition is free), so the bulk body always starts on an
//     even slot of an even-based bin and on a 16-byte aligned destination, however late the destination is known.
// Pass 1 (private cursors) keeps an odd tuple per partition for the next tile (carry[]), so it issues no scalar
// stores beyond the first run; pass 2 finishes an odd body with one scalar store.
// One TMA input buffer instead of two (the 64 KiB of bins need the room): tile i+1 is requested right after the
// rank phase of tile i has moved tile i into registers, and lands under its staging and write-out phases.
// A tile in which some partition would overflow its bin (skew) takes the staged kernel's compacting path inside
// this kernel (scan + lbase/gdst look-ups, SM-store write-out), so any distribution stays correct and fast.
// ---------------------------------------------------------------------------------------------
constexpr int kBinSlotsLog = 13;
#ifndef AQP_SCATTER_EARLY_REFILL
#define AQP_SCATTER_EARLY_REFILL 1
#endif
#ifndef AQP_SCATTER_LATE_WRITEOUT
#define AQP_SCATTER_LATE_WRITEOUT 0
#endif
constexpr size_t kBinsSmemBytes = (size_t) (kInBufTuples + (1 << kBinSlotsLog)) * sizeof(uint2);
static_assert(kScatterTile + kMaxFanout <= (1 << kBinSlotsLog), "the compacting path stages a whole tile plus carries");

template <bool kRot, bool kPeer, bool kPriv>
__global__ void __launch_bounds__(kScatterThreads, kScatterBlocksPerSM)
radix_scatter_bins_kernel(const uint2 *__restrict__ in, uint2 *__restrict__ out,
                          const uint32_t *__restrict__ seg_off, const uint32_t *__restrict__ seg_tile_start,
                          const uint32_t *__restrict__ seg_group, uint32_t nseg, DigitFn digit, uint32_t bits,
                          uint32_t *__restrict__ cursors, const uint32_t *__restrict__ block_base,
                          uint32_t tiles_per_block, PeerTable peers, uint32_t region_cap,
                          uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2 *inbuf = reinterpret_cast<uint2 *>(smem_raw);
    uint2 *bins = inbuf + kInBufTuples;
    __shared__ uint32_t cnt[kMaxFanout];      // tuples of this tile per partition
    __shared__ uint32_t scur[kMaxFanout];     // pass 1: CTA-private destination cursors
    __shared__ uint32_t hcarry[kMaxFanout];   // pass 1: 1 if carry[d] holds the odd tuple of the previous run
    __shared__ uint2 carry[kMaxFanout];
    __shared__ uint32_t lbase[kMaxFanout];    // compacting path only
    __shared__ uint32_t gdst[kMaxFanout];     // compacting path only
    __shared__ uint2 *s_peer[8];
    __shared__ uint32_t s_tstart[kMaxSegs + 1];
    __shared__ uint32_t s_soff[kMaxSegs + 1];
    __shared__ uint32_t s_group[kMaxSegs + 1];   // cursor group of a segment: read per tile by the thread that requests the input
    __shared__ uint32_t s_ovf[2];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_skip;         // histogram-free plan: a run of this tile found its region full (see region_cap)
    __shared__ uint32_t s_tile[2][4];   // {begin, end, cursor group} of the tile in flight / being processed
    __shared__ __align__(8) uint64_t mbar;
    __shared__ __align__(8) uint64_t mbar_free;   // input buffer read into registers by every warp

    const uint32_t fan = 1u << bits;
    const uint32_t lgcap = kBinSlotsLog - bits, cap = 1u << lgcap;
    for (uint32_t i = threadIdx.x; i <= nseg; i += kScatterThreads) {
        s_tstart[i] = seg_tile_start[i];
        s_soff[i] = seg_off[i];
        s_group[i] = (seg_group && i < nseg) ? seg_group[i] : i;
    }
    for (uint32_t i = threadIdx.x; i < fan; i += kScatterThreads) {
        cnt[i] = 0;
        hcarry[i] = 0;
        scur[i] = kPriv ? block_base[(size_t) blockIdx.x * fan + i] : 0u;
    }
    if (kPeer && threadIdx.x < 8) s_peer[threadIdx.x] = peers.base[threadIdx.x];
    if (threadIdx.x == 0) {
        s_ovf[0] = s_ovf[1] = 0;
        mbar_init(&mbar, 1);
        mbar_init(&mbar_free, kScatterThreads / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ntiles = s_tstart[nseg];
    uint32_t first, step, n_my;
    if (kPriv) {
        first = blockIdx.x * tiles_per_block;
        step = 1;
        n_my = first < ntiles ? min(tiles_per_block, ntiles - first) : 0;
    } else {
        first = blockIdx.x;
        step = gridDim.x;
        n_my = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    }
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
    // one thread: bulk-load tile #i of this CTA into the input buffer and publish its bounds (the other threads
    // read them after the next barrier instead of repeating the binary search)
    auto issue = [&](uint32_t i) {
        uint32_t seg, begin, end;
        tile_range(first + i * step, seg, begin, end);
        s_tile[i & 1][0] = begin;
        s_tile[i & 1][1] = end;
        s_tile[i & 1][2] = s_group[seg];   // (from shared memory: a global load here sat in front of every input request)
        const uint2 *src = in + begin;
        uint32_t skew = (uint32_t) ((reinterpret_cast<uintptr_t>(src) >> 3) & 1u);
        uint32_t bytes = ((end - begin + skew + 1) & ~1u) * (uint32_t) sizeof(uint2);
        mbar_expect_tx(&mbar, bytes);
        tma_load_1d(inbuf, src - skew, bytes, &mbar);
    };
    if (threadIdx.x == 0 && n_my > 0) issue(0);
    __syncthreads();   // s_tile[0] published
    // The input buffer is free again as soon as every warp holds its part of tile i in registers — before the rank
    // phase, not after it: each warp arrives on mbar_free, thread 0 waits for all of them and requests tile i+1,
    // which then has the rank, staging and write-out phases of tile i to land under.
    // (Pass 2 gains 5 % from it; pass 1, whose CTAs walk private contiguous tile ranges, loses 2.5 % and keeps the
    // request behind barrier (1): profiles/r02_sweep_join_early_refill.txt.)
    constexpr bool kEarlyRefill = AQP_SCATTER_EARLY_REFILL && !kPriv;
    auto refill = [&](uint32_t i) {
        if (!kEarlyRefill) return;
        __syncwarp();
        if ((threadIdx.x & 31u) == 0) mbar_arrive(&mbar_free);
        if (threadIdx.x == 0 && i + 1 < n_my) {
            mbar_wait(&mbar_free, i & 1);
            issue(i + 1);
        }
    };
    // The partition this thread reserves and writes out (if < fan). A bulk store takes its operands from uniform
    // registers, so a warp issues the stores of its lanes one after another: the partitions are dealt out to ALL
    // warps (the first fan/nwarps lanes of each) to keep that serial section short.
    constexpr uint32_t kWarps = kScatterThreads / 32;
    const uint32_t own_per = (fan + kWarps - 1) / kWarps;
    const uint32_t d_own = (threadIdx.x & 31u) < own_per ? (threadIdx.x >> 5) * own_per + (threadIdx.x & 31u) : fan;

    // tile #i of this CTA: wait for it, move it into registers, let the input buffer be refilled
    uint2 v[kScatterItems];
    auto load_tile = [&](uint32_t i) {
        const uint32_t begin = s_tile[i & 1][0], ntile = s_tile[i & 1][1] - begin;
        const uint2 *buf = inbuf + (uint32_t) ((reinterpret_cast<uintptr_t>(in + begin) >> 3) & 1u);
        mbar_wait(&mbar, i & 1);
        if (ntile == (uint32_t) kScatterTile) {   // full tile: no per-item bounds checks
#pragma unroll
            for (int j = 0; j < kScatterItems; ++j) v[j] = buf[j * kScatterThreads + threadIdx.x];
        } else {
#pragma unroll
            for (int j = 0; j < kScatterItems; ++j) {
                uint32_t k = j * kScatterThreads + threadIdx.x;
                if (k < ntile) v[j] = buf[k];
            }
        }
        refill(i);
    };
    // Experiment kept behind AQP_SCATTER_LATE_WRITEOUT (off): with shared cursors a run's destination comes back from a
    // global atomicAdd issued after barrier (1), and the owner threads wait for it in front of their bulk stores (ncu
    // source view: 24 % of the kernel's stall samples, long scoreboard, in the store issue). Moving the write-out of
    // tile i behind the next tile's input wait and register loads gives the atomic one more phase to come home - but
    // takes the same time away from the bulk stores, which must have read the bins before the next staging phase:
    // pass 2 1.85 -> 1.96 ms, pass 1 1.86 -> 1.91 ms (profiles/r02_sweep_join_late_writeout.txt). Not used.
    constexpr bool kLateWriteOut = AQP_SCATTER_LATE_WRITEOUT && !kPriv;
    if (kLateWriteOut && n_my > 0) load_tile(0);

    for (uint32_t i = 0; i < n_my; ++i) {
        const uint32_t begin = s_tile[i & 1][0], end = s_tile[i & 1][1], group = s_tile[i & 1][2];
        const uint32_t ntile = end - begin;

        if (!kLateWriteOut) load_tile(i);
        uint32_t rank[kScatterItems];
        bool tight = false;
        if (ntile == (uint32_t) kScatterTile) {
#pragma unroll
            for (int j = 0; j < kScatterItems; ++j) {
                rank[j] = atomicAdd(&cnt[digit.template get<kRot>(v[j].x)], 1u);
                tight |= rank[j] + 2 > cap;   // bin full (one slot is kept for the carried tuple)
            }
        } else {
#pragma unroll
            for (int j = 0; j < kScatterItems; ++j) {
                uint32_t k = j * kScatterThreads + threadIdx.x;
                if (k < ntile) {
                    rank[j] = atomicAdd(&cnt[digit.template get<kRot>(v[j].x)], 1u);
                    tight |= rank[j] + 2 > cap;
                }
            }
        }
        if (tight) s_ovf[i & 1] = 1;
        // the previous tile's bulk stores must have read the bins before anybody stages into them again
        if (d_own < fan) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();   // (1) tile histogram complete, input buffer consumed, bins free

        if (threadIdx.x == 0) {
            if (!kEarlyRefill && i + 1 < n_my) issue(i + 1);
            s_ovf[(i + 1) & 1] = 0;
            s_skip = 0;
        }
        if (!s_ovf[i & 1]) {
            // ---------------- fixed bins: slot = (digit << lgcap) + rank ----------------
            if (ntile == (uint32_t) kScatterTile) {
#pragma unroll
                for (int j = 0; j < kScatterItems; ++j) bins[(digit.template get<kRot>(v[j].x) << lgcap) + rank[j]] = v[j];
            } else {
#pragma unroll
                for (int j = 0; j < kScatterItems; ++j) {
                    uint32_t k = j * kScatterThreads + threadIdx.x;
                    if (k < ntile) bins[(digit.template get<kRot>(v[j].x) << lgcap) + rank[j]] = v[j];
                }
            }
            uint32_t n = 0, g = 0;
            if (d_own < fan) {
                const uint32_t c = cnt[d_own];
                cnt[d_own] = 0;
                n = c;
                if (kPriv) {
                    if (hcarry[d_own]) {
                        bins[(d_own << lgcap) + c] = carry[d_own];
                        ++n;
                    }
                    g = scur[d_own];
                } else if (n) {
                    g = atomicAdd(&cursors[(group << bits) + d_own], n);   // its latency hides behind the barrier
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged tuples visible to the TMA unit
            __syncthreads();   // (2) tile staged
            if (kLateWriteOut && i + 1 < n_my) load_tile(i + 1);
            if (d_own < fan) {
                // histogram-free plan (api.cu): cursor c runs inside the region [c cap, (c + 1) cap); a run that does
                // not fit is dropped and reported - the caller then repeats the join with exact offsets. (Checked here,
                // where g is needed anyway: next to the atomicAdd the check waited for it in front of the barrier and
                // cost both passes 4-6 %.)
                if (!kPriv && region_cap && n && g + n > ((group << bits) + d_own + 1) * region_cap) {
                    *overflow = 1;
                    n = 0;
                }
                if (n) {
                    uint2 *dst = kPeer ? s_peer[d_own >> peers.per_shift] : out;
                    const uint2 *src = bins + (d_own << lgcap);
                    if (g & 1u) {   // odd destination start: the run's last tuple goes there
                        dst[g] = src[n - 1];
                        ++g;
                        --n;
                    }
                    const uint32_t body = n & ~1u;
                    if (body) tma_store_1d(dst + g, src, body * (uint32_t) sizeof(uint2));
                    if (kPriv) {
                        if (n & 1u) carry[d_own] = src[body];
                        hcarry[d_own] = n & 1u;
                        scur[d_own] = g + body;
                    } else if (n & 1u) {
                        dst[g + body] = src[body];
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // ---------------- a bin would overflow (skew): compact the tile like radix_scatter_kernel ----------------
            uint32_t my_g[kMaxFanout / 32], my_b[kMaxFanout / 32];
            if (threadIdx.x < 32) {
                const uint32_t per = (fan + 31) / 32;
                uint32_t c[kMaxFanout / 32], hc[kMaxFanout / 32], sum = 0;
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    bool ok = k < (int) per && d < fan;
                    c[k] = ok ? cnt[d] : 0;
                    hc[k] = (ok && kPriv) ? hcarry[d] : 0;
                    if (ok && kPriv) cnt[d] = 0;   // (shared cursors: the partition's owner thread clears it below)
                    sum += c[k] + hc[k];
                }
                uint32_t run = warp_incl_scan(sum) - sum;
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    my_b[k] = run;
                    const uint32_t n = c[k] + hc[k];
                    if (k < (int) per && d < fan) {
                        lbase[d] = run;
                        if (hc[k]) {
                            bins[run + c[k]] = carry[d];   // the carried tuple joins the run
                            hcarry[d] = 0;
                        }
                        if (kPriv) {
                            my_g[k] = scur[d];
                            scur[d] = my_g[k] + n;
                        }
                    }
                    run += n;
                }
                if (threadIdx.x == 31) s_total = run;
            }
            __syncthreads();   // lbase ready
            uint32_t own_n = 0, own_g = 0;
            if (!kPriv && d_own < fan) {
                // shared cursors: every partition's owner thread reserves its run - 2^bits global atomics in flight at
                // once, their latency under the staging stores below (warp 0 doing them one after another, 8 per lane
                // at 2^8 partitions, cost 8 round trips per tile: TPC-H Q12's pass 1 ran at 1.8 TB/s that way)
                own_n = cnt[d_own];
                cnt[d_own] = 0;
                own_g = own_n ? atomicAdd(&cursors[(group << bits) + d_own], own_n) : 0u;
            }
#pragma unroll
            for (int j = 0; j < kScatterItems; ++j) {
                uint32_t k = j * kScatterThreads + threadIdx.x;
                if (k < ntile) bins[lbase[digit.template get<kRot>(v[j].x)] + rank[j]] = v[j];
            }
            if (!kPriv && d_own < fan) {   // the atomics have had the staging stores to come back
                if (region_cap && own_n && own_g + own_n > ((group << bits) + d_own + 1) * region_cap) {
                    *overflow = 1;   // region full (histogram-free plan): drop the tile, report
                    s_skip = 1;
                }
                gdst[d_own] = own_g - lbase[d_own];
            }
            if (kPriv && threadIdx.x < 32) {
                const uint32_t per = (fan + 31) / 32;
#pragma unroll
                for (int k = 0; k < kMaxFanout / 32; ++k) {
                    uint32_t d = threadIdx.x * per + k;
                    if (k < (int) per && d < fan) gdst[d] = my_g[k] - my_b[k];
                }
            }
            __syncthreads();   // tile compacted, destinations known
            if (kLateWriteOut && i + 1 < n_my) load_tile(i + 1);
            const uint32_t total = s_skip ? 0u : s_total;
            for (uint32_t s = threadIdx.x; s < total; s += kScatterThreads) {
                uint2 t = bins[s];
                const uint32_t d = digit.template get<kRot>(t.x);
                if (kPeer)
                    s_peer[d >> peers.per_shift][gdst[d] + s] = t;
                else
                    out[gdst[d] + s] = t;
            }
            // (the next tile's barrier (1) separates these reads from its staging)
        }
    }
    if (d_own < fan) {
        if (kPriv && hcarry[d_own]) {   // flush the carried tuples
            uint2 *dst = kPeer ? s_peer[d_own >> peers.per_shift] : out;
            dst[scur[d_own]] = carry[d_own];
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the TMA unit is done with shared memory
    }
}

// ---------------------------------------------------------------------------------------------
// scatter into PEER memory (the fused exchange), line-aligned variant
//
// Over NVLink a store — SM-issued or TMA — reaches the link rate (714 GB/s) only when it covers whole, aligned
// 128-byte lines; 256-byte runs that start at arbitrary 16-byte offsets reach 550 GB/s (profiles/r01_bulkbench.txt),
// because every partial line travels as its own packet and cannot be merged on the way. So here a partition's bin
// is a RING that lives across tiles: what leaves it per tile is the part that ends on a destination line boundary,
// the remainder (< 16 tuples) stays for the next tile. One packed shared-memory atomicAdd per tuple returns the
// tuple's position in its partition's stream (high half; kept congruent to the destination index mod 2^16, so ring
// slot = pos mod cap and "aligned in pos" = "aligned at the destination") and the bin's fill (low half). The owner
// thread of a partition flushes [oldest, last line boundary) as one or two TMA bulk stores (two when the range wraps
// around the ring) and subtracts it from the fill BEFORE the barrier that releases the next tile's rank phase, so
// fills are always exact. A tuple that finds its bin full (skew) is stored directly behind the flushed range from
// registers. Unaligned heads (the first run of a CTA in a partition, the run after an overflow) and the final
// remainders are scalar stores.
// ---------------------------------------------------------------------------------------------
template <bool kRot>
__global__ void __launch_bounds__(kScatterThreads, kScatterBlocksPerSM)
radix_scatter_peer_kernel(const uint2 *__restrict__ in, uint32_t n, DigitFn digit, uint32_t bits,
                          const uint32_t *__restrict__ block_base, uint32_t tiles_per_block, PeerTable peers) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint2 *inbuf = reinterpret_cast<uint2 *>(smem_raw);
    uint2 *bins = inbuf + kInBufTuples;
    __shared__ uint32_t w[kMaxFanout];      // (position of the next tuple mod 2^16) << 16 | tuples in the bin
    __shared__ uint32_t scur[kMaxFanout];   // destination index of the bin's oldest tuple
    __shared__ uint32_t gst[kMaxFanout];    // overflow path: destination index ...
    __shared__ uint32_t spos[kMaxFanout];   //   ... and position of the oldest tuple when the tile was flushed
    __shared__ uint2 *s_peer[8];
    __shared__ uint32_t s_ovf[2];
    __shared__ __align__(8) uint64_t mbar;

    const uint32_t fan = 1u << bits;
    const uint32_t lgcap = kBinSlotsLog - bits, cap = 1u << lgcap, cmask = cap - 1;
    for (uint32_t i = threadIdx.x; i < fan; i += kScatterThreads) {
        const uint32_t base = block_base[(size_t) blockIdx.x * fan + i];
        scur[i] = base;
        w[i] = base << 16;
    }
    if (threadIdx.x < 8) s_peer[threadIdx.x] = peers.base[threadIdx.x];
    if (threadIdx.x == 0) {
        s_ovf[0] = s_ovf[1] = 0;
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ntiles = (n + kScatterTile - 1) / kScatterTile;
    const uint32_t first = blockIdx.x * tiles_per_block;
    const uint32_t n_my = first < ntiles ? min(tiles_per_block, ntiles - first) : 0;
    auto issue = [&](uint32_t i) {   // one thread: bulk-load tile #i of this CTA (tiles start on even tuple indices)
        const uint32_t begin = (first + i) * kScatterTile, end = min(begin + (uint32_t) kScatterTile, n);
        const uint32_t bytes = ((end - begin + 1) & ~1u) * (uint32_t) sizeof(uint2);
        mbar_expect_tx(&mbar, bytes);
        tma_load_1d(inbuf, in + begin, bytes, &mbar);
    };
    if (threadIdx.x == 0 && n_my > 0) issue(0);
    constexpr uint32_t kWarps = kScatterThreads / 32;
    const uint32_t own_per = (fan + kWarps - 1) / kWarps;
    const uint32_t d_own = (threadIdx.x & 31u) < own_per ? (threadIdx.x >> 5) * own_per + (threadIdx.x & 31u) : fan;

    for (uint32_t i = 0; i < n_my; ++i) {
        const uint32_t begin = (first + i) * kScatterTile, end = min(begin + (uint32_t) kScatterTile, n);
        const uint32_t ntile = end - begin;
        const bool last = i + 1 == n_my;

        mbar_wait(&mbar, i & 1);
        uint2 v[kScatterItems];
        uint32_t old[kScatterItems];
        bool full = false;
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            if (k < ntile) v[j] = inbuf[k];
        }
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            old[j] = 0;
            if (k < ntile) {
                old[j] = atomicAdd(&w[digit.template get<kRot>(v[j].x)], 0x10001u);
                full |= (old[j] & 0xFFFFu) >= cap;
            }
        }
        if (full) s_ovf[i & 1] = 1;
        if (d_own < fan) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();   // (1) positions assigned, input buffer consumed, previous flush has left the bins

        if (threadIdx.x == 0) {
            if (i + 1 < n_my) issue(i + 1);
            s_ovf[(i + 1) & 1] = 0;
        }
#pragma unroll
        for (int j = 0; j < kScatterItems; ++j) {
            uint32_t k = j * kScatterThreads + threadIdx.x;
            if (k < ntile && (old[j] & 0xFFFFu) < cap)
                bins[(digit.template get<kRot>(v[j].x) << lgcap) + ((old[j] >> 16) & cmask)] = v[j];
        }
        uint32_t f_start = 0, f_n = 0, f_g = 0;
        if (d_own < fan) {
            const uint32_t wd = w[d_own];
            const uint32_t fill = wd & 0xFFFFu, pos_end = wd >> 16;
            const uint32_t start = (pos_end - fill) & 0xFFFFu;
            const bool ovf = fill > cap;
            const uint32_t inring = ovf ? cap : fill;
            uint32_t consumed;
            if (ovf || last) {
                f_n = inring;
                consumed = fill;
            } else {
                const uint32_t left = (start + inring) & 15u;
                f_n = inring > left ? inring - left : 0u;
                consumed = f_n;
            }
            f_start = start;
            f_g = scur[d_own];
            scur[d_own] = f_g + consumed;
            w[d_own] = wd - consumed;   // the fill is exact again before the next tile's rank phase starts
            if (ovf) {
                gst[d_own] = f_g;
                spos[d_own] = start;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();   // (2) tile staged, fills updated, overflow destinations published
        if (d_own < fan) {
            if (f_n) {
                uint2 *dst = s_peer[d_own >> peers.per_shift] + f_g;
                const uint2 *ring = bins + (d_own << lgcap);
                uint32_t off = 0;
                uint32_t head = (16u - (f_start & 15u)) & 15u;   // up to the first destination line boundary
                head = head < f_n ? head : f_n;
                for (; off < head; ++off) dst[off] = ring[(f_start + off) & cmask];
                const uint32_t body = (f_n - off) & ~15u;
                if (body) {
                    const uint32_t r0 = (f_start + off) & cmask;
                    const uint32_t len1 = body < cap - r0 ? body : cap - r0;
                    tma_store_1d(dst + off, ring + r0, len1 * (uint32_t) sizeof(uint2));
                    if (body > len1) tma_store_1d(dst + off + len1, ring, (body - len1) * (uint32_t) sizeof(uint2));
                    off += body;
                }
                for (; off < f_n; ++off) dst[off] = ring[(f_start + off) & cmask];   // only after an overflow / at the end
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (s_ovf[i & 1]) {   // tuples that found their bin full: straight from registers, behind the flushed range
#pragma unroll
            for (int j = 0; j < kScatterItems; ++j) {
                uint32_t k = j * kScatterThreads + threadIdx.x;
                if (k < ntile && (old[j] & 0xFFFFu) >= cap) {
                    const uint32_t d = digit.template get<kRot>(v[j].x);
                    s_peer[d >> peers.per_shift][gst[d] + (((old[j] >> 16) - spos[d]) & 0xFFFFu)] = v[j];
                }
            }
        }
    }
    if (d_own < fan) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int radix_scatter_launch(const row_t *d_in, row_t *d_out, const uint32_t *d_seg_off,
                         const uint32_t *d_seg_tile_start, const uint32_t *d_seg_group, uint32_t nseg, uint64_t n_total,
                         DigitFn digit, uint32_t bits, uint32_t *d_cursors, const uint32_t *d_block_base,
                         uint32_t nblocks, uint32_t tiles_per_block, cudaStream_t st, const PeerTable *peers,
                         uint32_t region_cap, uint32_t *d_overflow) {
    if (bits > (uint32_t) kMaxFanoutBits || nseg > (uint32_t) kMaxSegs) {
        set_error("radix_scatter: fan-out too large");
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(d_in) & 7u) != 0) {
        set_error("radix_scatter: relation must be 8-byte aligned");
        return -1;
    }
    if (n_total == 0) return 0;
    static unsigned attr_set = ~0u;   // device epoch the opt-in was made for (b200_shutdown + b200_init(other device) re-arms it)
    if (attr_set != g_device_epoch) {
#define AQP_SCATTER_ATTR(ROT, PEER, BULK)                                                                             \
    AQP_CUDA_OK(cudaFuncSetAttribute(radix_scatter_kernel<ROT, PEER, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int) kScatterSmemBytes))
        AQP_SCATTER_ATTR(false, false, false);
        AQP_SCATTER_ATTR(false, false, true);
        AQP_SCATTER_ATTR(true, false, false);
        AQP_SCATTER_ATTR(true, false, true);
        AQP_SCATTER_ATTR(true, true, false);
        AQP_SCATTER_ATTR(true, true, true);
#undef AQP_SCATTER_ATTR
        attr_set = g_device_epoch;
    }
    uint32_t grid;
    if (d_block_base) {
        grid = nblocks;
    } else {
        uint64_t max_tiles = n_total / kScatterTile + nseg;
        uint64_t g = (uint64_t) kNumSMs * kScatterBlocksPerSM;
        grid = (uint32_t) (max_tiles < g ? max_tiles : g);
    }
    PeerTable none{};
    const uint2 *in = reinterpret_cast<const uint2 *>(d_in);
    uint2 *out = reinterpret_cast<uint2 *>(d_out);
    const bool peer = peers && peers->n;
    static const bool staged = getenv("B200_AQP_SCATTER") && !strcmp(getenv("B200_AQP_SCATTER"), "staged");
    bool aligned16 = (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
    if (peer)
        for (uint32_t i = 0; i < peers->n; ++i) aligned16 = aligned16 && (reinterpret_cast<uintptr_t>(peers->base[i]) & 15u) == 0;
    if (!staged && aligned16 && (1u << bits) <= (uint32_t) kScatterThreads) {
        static unsigned bins_attr_set = ~0u;   // device epoch the opt-in was made for (b200_shutdown + b200_init(other device) re-arms it)
        if (bins_attr_set != g_device_epoch) {
#define AQP_BINS_ATTR(ROT, PEER, PRIV)                                                                                  \
    AQP_CUDA_OK(cudaFuncSetAttribute(radix_scatter_bins_kernel<ROT, PEER, PRIV>,                                        \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kBinsSmemBytes))
            AQP_BINS_ATTR(false, false, false);
            AQP_BINS_ATTR(false, false, true);
            AQP_BINS_ATTR(true, false, false);
            AQP_BINS_ATTR(true, false, true);
            AQP_BINS_ATTR(true, true, true);
#undef AQP_BINS_ATTR
            bins_attr_set = g_device_epoch;
        }
#define AQP_BINS_LAUNCH(ROT, PEER, PRIV)                                                                               \
    radix_scatter_bins_kernel<ROT, PEER, PRIV><<<grid, kScatterThreads, kBinsSmemBytes, st>>>(                          \
        in, out, d_seg_off, d_seg_tile_start, d_seg_group, nseg, digit, bits, d_cursors, d_block_base, tiles_per_block, \
        peer ? *peers : none, region_cap, d_overflow)
        static const bool peer_ring_off = getenv("B200_AQP_PEER_RING") && atoi(getenv("B200_AQP_PEER_RING")) == 0;
        bool aligned128 = true;
        if (peer)
            for (uint32_t i = 0; i < peers->n; ++i) aligned128 = aligned128 && (reinterpret_cast<uintptr_t>(peers->base[i]) & 127u) == 0;
        if (peer) {
            if (!d_block_base) {
                set_error("radix_scatter: the peer variant needs CTA-private cursors");
                return -1;
            }
            // line-aligned ring flush needs >= 64 slots per bin (a remainder of up to 15 tuples plus a tile's arrivals),
            // a single input segment that starts on an even tuple and 128-byte aligned receive buffers
            if (bits <= (uint32_t) kBinSlotsLog - 6 && nseg == 1 && aligned128 && !peer_ring_off &&
                (reinterpret_cast<uintptr_t>(d_in) & 15u) == 0 && n_total < 0xFFFFFFFFull) {
                static unsigned peer_attr_set = ~0u;   // device epoch the opt-in was made for (b200_shutdown + b200_init(other device) re-arms it)
                if (peer_attr_set != g_device_epoch) {
                    AQP_CUDA_OK(cudaFuncSetAttribute(radix_scatter_peer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int) kBinsSmemBytes));
                    peer_attr_set = g_device_epoch;
                }
                radix_scatter_peer_kernel<true><<<grid, kScatterThreads, kBinsSmemBytes, st>>>(
                    in, (uint32_t) n_total, digit, bits, d_block_base, tiles_per_block, *peers);
            } else {
                AQP_BINS_LAUNCH(true, true, true);
            }
        } else if (digit.rot) {
            if (d_block_base) AQP_BINS_LAUNCH(true, false, true); else AQP_BINS_LAUNCH(true, false, false);
        } else {
            if (d_block_base) AQP_BINS_LAUNCH(false, false, true); else AQP_BINS_LAUNCH(false, false, false);
        }
#undef AQP_BINS_LAUNCH
        AQP_LAUNCHED();
        AQP_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (region_cap) {
        set_error("radix_scatter: region limits need the fixed-bin kernel (16-byte aligned output, fan-out <= 256)");
        return -1;
    }
    // bulk write-out needs CTA-private cursors (the destination parity must be known when the tile is staged) and a
    // 16-byte aligned destination buffer; B200_AQP_SCATTER_BULK=0 keeps the SM-store write-out (A/B measurements)
    static const bool bulk_off = getenv("B200_AQP_SCATTER_BULK") && atoi(getenv("B200_AQP_SCATTER_BULK")) == 0;
    bool bulk = d_block_base && !bulk_off && kMaxFanout <= kScatterThreads;
    if (bulk && !peer && (reinterpret_cast<uintptr_t>(d_out) & 15u)) bulk = false;
    if (bulk && peer)
        for (uint32_t i = 0; i < peers->n; ++i)
            if (reinterpret_cast<uintptr_t>(peers->base[i]) & 15u) bulk = false;
#define AQP_SCATTER_LAUNCH(ROT, PEER)                                                                              \
    do {                                                                                                           \
        if (bulk)                                                                                                  \
            radix_scatter_kernel<ROT, PEER, true><<<grid, kScatterThreads, kScatterSmemBytes, st>>>(               \
                in, out, d_seg_off, d_seg_tile_start, d_seg_group, nseg, digit, bits, d_cursors, d_block_base,     \
                tiles_per_block, peer ? *peers : none);                                                            \
        else                                                                                                       \
            radix_scatter_kernel<ROT, PEER, false><<<grid, kScatterThreads, kScatterSmemBytes, st>>>(              \
                in, out, d_seg_off, d_seg_tile_start, d_seg_group, nseg, digit, bits, d_cursors, d_block_base,     \
                tiles_per_block, peer ? *peers : none);                                                            \
    } while (0)
    if (peer)
        AQP_SCATTER_LAUNCH(true, true);
    else if (digit.rot)
        AQP_SCATTER_LAUNCH(true, false);
    else
        AQP_SCATTER_LAUNCH(false, false);
#undef AQP_SCATTER_LAUNCH
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
