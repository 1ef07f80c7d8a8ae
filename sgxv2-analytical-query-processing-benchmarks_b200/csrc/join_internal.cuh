// join_internal.cuh — constants, plan structs and launcher declarations shared by the join's
// translation units (partition.cu, build_probe.cu, gen.cu, api.cu).
#pragma once

#include "common.cuh"

namespace aqp {

// grow-only device buffer
// Device buffer that grows on demand. Every instance (globals, function-local statics, members) is listed in a
// registry so that b200_shutdown can free them all: their memory belongs to the device context that is being left.
struct DevBuf;
void devbuf_register(DevBuf *b, bool add);
void devbuf_release_all();
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    DevBuf() { devbuf_register(this, true); }
    DevBuf(const DevBuf &o) : p(o.p), cap(o.cap) { devbuf_register(this, true); }
    DevBuf &operator=(const DevBuf &o) = default;
    ~DevBuf() { devbuf_register(this, false); }
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 4) + 256;   // a little slack so near-equal sizes do not re-allocate
        AQP_CUDA_OK(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// ---- partitioning geometry ---------------------------------------------------------------------
constexpr int kMaxFanoutBits = 8;                 // bits per pass
constexpr int kMaxFanout = 1 << kMaxFanoutBits;
constexpr int kMaxSmemHistBits = 15;              // widest shared-memory histogram (128 KiB)
#ifndef AQP_SCATTER_TILE
#define AQP_SCATTER_TILE 4096
#endif
#ifndef AQP_SCATTER_THREADS
#define AQP_SCATTER_THREADS 256
#endif
constexpr int kScatterTile = AQP_SCATTER_TILE;    // tuples per scatter tile (32 KiB staging + 2 x 32 KiB TMA ring)

// ---- build/probe geometry ------------------------------------------------------------------------
constexpr int kBuildCap = 8192;                   // R tuples per shared-memory hash table (64 KiB)
constexpr int kProbeChunk = 32768;                // S tuples per work item
constexpr int kJoinThreads = 512;

// Partition digit of a key: the reference's (key & MASK) >> R (radix_join.cpp:47) on `bits` key bits
// starting at `shift`, optionally with its low `rbits` bits rotated right by `rot`. The rotation is the
// multi-GPU routing order: with G = 2^rot GPUs the pass-1 partitions owned by one GPU (those whose low
// rot key bits equal the GPU's rank) become a contiguous range, so the pass-1 output doubles as the
// all-to-all send buffer. rot = 0 is the identity.
struct DigitFn {
    uint32_t shift, mask, m1, rot, lrot;
#ifdef __CUDACC__
    __device__ __forceinline__ uint32_t operator()(uint32_t key) const {
        uint32_t d = (key >> shift) & mask;
        uint32_t p = d & m1;
        return (d & ~m1) | (((p >> rot) | (p << lrot)) & m1);
    }
    // kRot = false: the plain single-GPU digit, two instructions on the hot loops
    template <bool kRot>
    __device__ __forceinline__ uint32_t get(uint32_t key) const {
        return kRot ? (*this)(key) : ((key >> shift) & mask);
    }
#endif
};
inline DigitFn make_digit(uint32_t shift, uint32_t bits, uint32_t rbits = 0, uint32_t rot = 0) {
    DigitFn f;
    f.shift = shift;
    f.mask = (1u << bits) - 1;
    f.m1 = rot ? (1u << rbits) - 1 : 0;
    f.rot = rot;
    f.lrot = rot ? rbits - rot : 0;
    return f;
}

// Destination buffers of the fused scatter+exchange: partition d of the routed pass-1 digit is written
// into base[d >> per_shift], the receive buffer of the GPU that owns it (a peer-mapped pointer, or this
// GPU's own buffer). n == 0 means "no peer table": write to the kernel's `out` argument.
struct PeerTable {
    uint2 *base[8];
    uint32_t per_shift;
    uint32_t n;
};

struct RelPlan {
    const uint32_t *hist;       // [2^B]   raw-digit histogram
    uint32_t *part_off;         // [2^B+1] final partition starts, order f = p1*F2 + p2
    uint32_t *cursor1;          // [F1]    pass-1 write cursors
    uint32_t *cursor2;          // [2^B]   pass-2 write cursors
    uint32_t *seg_off;          // [F1+1]  pass-1 partition starts (= pass-2 input segments)
    uint32_t *seg_tile_start;   // [F1+1]  first pass-2 tile of each segment
    uint32_t *seg1;             // [4]     pass-1 single-segment tables {0, n, 0, ceil(n/tile)}
    const uint32_t *block_hist; // [nblocks1][F1] pass-1 histogram of each pass-1 scatter block (or null)
    uint32_t *block_base;       // [nblocks1][F1] private pass-1 write cursors of each block
};
struct PlanArgs {
    RelPlan rel[2];
    uint32_t bits1, bits2;
    uint32_t nblocks1;          // CTAs of the pass-1 scatter (= rows of block_hist)
};

struct ShardRelPlan {
    const uint32_t *hist;       // [nparts]   this rank's slice of the global histogram, final order
    const uint32_t *seg_off;    // [nseg+1]   received segments
    uint32_t *part_off;         // [nparts+1]
    uint32_t *cursor2;          // [nparts]
    uint32_t *seg_tile_start;   // [nseg+1]
};
struct ShardPlanArgs {
    ShardRelPlan rel[2];
    uint32_t nparts, nseg;
    const uint32_t *seg_group;  // [nseg] or null; kGapSegment marks a hole between two sources' regions (no tiles)
};
constexpr uint32_t kGapSegment = 0xffffffffu;
constexpr int kMaxSegs = kMaxFanout + 8;   // received segments: (source, local partition) + one gap per source

struct JoinResult {              // device-side accumulators
    unsigned long long matches;
    unsigned long long checksum;
    unsigned long long keysum;
    unsigned long long out_count;   // triples reserved in the output buffer
};

// api.cu
int join_device_internal(const row_t *dR, uint64_t nR, const row_t *dS, uint64_t nS, output_triple_t *d_out,
                         uint64_t out_cap, b200_join_stats_t *stats, cudaStream_t st, uint32_t dead_bits = 0);
cudaStream_t library_stream();

// partition.cu
int radix_hist_device(const row_t *d_in, uint64_t n, DigitFn digit, uint32_t bits, uint32_t *d_hist, uint32_t nblocks,
                      uint64_t chunk, uint32_t bits1, uint32_t *d_block_hist, cudaStream_t st);
int exclusive_scan_u32_device(const uint32_t *d_in, uint32_t n, uint32_t *d_out, cudaStream_t st);
int plan_offsets_device(const PlanArgs &a, cudaStream_t st);
// d_block_base == nullptr: runs are reserved with global atomics on d_cursors (any grid).
// d_block_base != nullptr: exactly nblocks CTAs, CTA b owns tiles [b*tiles_per_block, ...) and writes at
// its private cursors d_block_base[b][2^bits] (single segment only).
// d_seg_group (or null = identity) maps an input segment to the cursor group it scatters into:
// cursor index = (group << bits) + digit. Several segments may share a group (multi-GPU: the same
// pass-1 partition received from different source GPUs).
int radix_scatter_launch(const row_t *d_in, row_t *d_out, const uint32_t *d_seg_off, const uint32_t *d_seg_tile_start,
                         const uint32_t *d_seg_group, uint32_t nseg, uint64_t n_total, DigitFn digit, uint32_t bits,
                         uint32_t *d_cursors, const uint32_t *d_block_base, uint32_t nblocks, uint32_t tiles_per_block,
                         cudaStream_t st, const PeerTable *peers = nullptr, uint32_t region_cap = 0,
                         uint32_t *d_overflow = nullptr);

// ---- histogram-free plan (api.cu: join_device_optimistic): every partition of both passes gets a region of fixed
// capacity instead of an exact offset, so no histogram pass is needed; cursors run inside the regions, a run that does
// not fit raises a flag and the join is repeated with exact offsets.
struct RegionRel {
    uint32_t n, cap1, cap2;             // tuples; region capacity (tuples) of a pass-1 / a final partition
    uint32_t *cur1;                     // [F1]   pass-1 cursors: cursor p runs inside [p cap1, (p + 1) cap1)
    uint32_t *seg1;                     // [4]    pass-1 input: {0, n} and its tile table {0, tiles}
    uint32_t *seg_off, *seg_tile;       // [2 F1 + 1] pass-2 input segments (data, gap, data, gap, ...) and tile table
    uint32_t *cur2;                     // [P]    pass-2 cursors, likewise with cap2
    uint32_t *beg, *end;                // [P]    final partitions for build/probe
};
struct RegionArgs {
    uint32_t bits1, bits2;
    uint32_t *seg_group;                // [2 F1] pass-1 partition of a pass-2 segment, kGapSegment for the gaps
    RegionRel rel[2];
};
// the test that comes before it: every `stride`-th 128-byte line of each relation is counted into a full-width
// histogram (d_hist_*[2^bits], zeroed by the caller), both relations in one launch; region_verdict_device reduces the two sampled histograms to
// d_out[rel * 3 + {0, 1, 2}] = {samples, largest pass-1 partition, largest final partition} (in samples)
int region_sample_device(const row_t *d_R, uint64_t nR, uint32_t stride_R, uint32_t *d_hist_R, const row_t *d_S, uint64_t nS,
                         uint32_t stride_S, uint32_t *d_hist_S, uint32_t bits, cudaStream_t st);
int region_verdict_device(const uint32_t *d_hist_R, const uint32_t *d_hist_S, uint32_t bits1, uint32_t bits2, uint32_t *d_out,
                          cudaStream_t st);
int region_init_device(const RegionArgs &a, cudaStream_t st);    // before pass 1
int region_plan2_device(const RegionArgs &a, cudaStream_t st);   // between the passes
int region_plan3_device(const RegionArgs &a, cudaStream_t st);   // before build/probe
int block_base_device(const uint32_t *d_block_hist, const uint32_t *d_part_start, uint32_t fan, uint32_t nblocks,
                      uint32_t *d_block_base, uint32_t *d_counts, uint32_t *d_seg1, uint32_t n, cudaStream_t st);
uint32_t pass1_blocks();
int plan_pass1_device(const uint32_t *d_hist, uint32_t bits1, uint32_t bits2, uint32_t *d_part1_off, uint32_t *d_seg1,
                      const uint32_t *d_block_hist, uint32_t *d_block_base, uint32_t nblocks, cudaStream_t st);
int plan_shard_device(const ShardPlanArgs &a, cudaStream_t st);
int region_dest_device(const uint32_t *d_counts1, uint32_t world, uint32_t rank, uint32_t bits1, uint64_t cap_r, uint64_t cap_s,
                       uint32_t *d_dest_off, unsigned long long *d_kept, cudaStream_t st);
int region_plan_device(const uint32_t *d_counts_all, uint32_t world, uint32_t rank, uint32_t bits1, uint32_t bits2,
                       const uint32_t *d_hist_global, uint64_t cap_r, uint64_t cap_s, uint32_t *d_seg_off,
                       uint32_t *d_seg_group, uint32_t *d_hist_slice, cudaStream_t st);
int exchange_plan_device(const uint32_t *d_counts_all, uint32_t world, uint32_t rank, uint32_t bits1, uint32_t bits2,
                         const uint32_t *d_hist_global, uint32_t *d_seg_off, uint32_t *d_dest_off, uint32_t *d_hist_slice,
                         unsigned long long *d_host_vals, cudaStream_t st);
int single_segment_setup(uint32_t n, const uint32_t *d_offsets, uint32_t fan, uint32_t *d_cursors,
                         uint32_t *d_seg_tables, cudaStream_t st);

// build_probe.cu
int join_items_device(const uint32_t *d_offR, const uint32_t *d_offS, uint32_t nparts, uint32_t *d_item_start,
                      uint2 *d_items, cudaStream_t st, const uint32_t *d_endR = nullptr, const uint32_t *d_endS = nullptr);
int build_probe_device(const row_t *d_R, const uint32_t *d_offR, const row_t *d_S, const uint32_t *d_offS,
                       const uint32_t *d_item_start, const uint2 *d_items, uint32_t nparts, uint64_t max_items,
                       uint32_t hash_shift, JoinResult *d_res, output_triple_t *d_out, uint64_t out_cap,
                       cudaStream_t st, const uint32_t *d_endR = nullptr, const uint32_t *d_endS = nullptr);

// gen.cu
int gen_pk_device(row_t *d_rel, uint64_t n_total, uint64_t row_begin, uint64_t n, uint64_t seed, cudaStream_t st);
int gen_fk_device(row_t *d_rel, uint64_t n_total, uint64_t maxid, uint64_t row_begin, uint64_t n, uint64_t seed,
                  cudaStream_t st);
int gen_zipf_device(row_t *d_rel, uint64_t maxid, double z, uint64_t row_begin, uint64_t n, uint64_t seed,
                    cudaStream_t st);
int set_rowid_payload_device(row_t *d_rel, uint64_t row_begin, uint64_t n, cudaStream_t st);

// scan.cu
int bitvector_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_out, cudaStream_t st);
int scan_count_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_count, cudaStream_t st);
size_t index_scan_scratch_bytes(size_t n);
void scan_release();
// the planner with the caller's knowledge of dead low key bits (api.cu plan_bits)
void join_plan_internal(uint64_t nR, uint32_t dead_bits, uint32_t *total, uint32_t *b1, uint32_t *b2);

// ---- materialising shard join for the multi-GPU host (api.cu) ------------------------------------
int shard_join_materialize_internal(const row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R, const row_t *d_S, uint64_t nS,
                                    const uint32_t *d_segoff_S, const uint32_t *d_seg_group, uint32_t nseg, uint32_t ngroups,
                                    uint32_t shift2, uint32_t bits2, const uint32_t *d_hist_R, const uint32_t *d_hist_S,
                                    uint32_t hash_shift, output_triple_t *d_out, uint64_t out_cap, uint64_t *d_result4,
                                    cudaStream_t st);
int shard_probe_again_internal(output_triple_t *d_out, uint64_t out_cap, uint64_t *d_result4, cudaStream_t st);

// ---- host <-> device copies for the host-buffer entry points (hostcopy.cpp) ----------------------
int copy_h2d_any(void *dev, const void *host, size_t bytes, cudaStream_t st);
int copy_d2h_any(void *host, const void *dev, size_t bytes, cudaStream_t st);
void hostcopy_release();
void *host_alloc_prefer_pinned(size_t bytes);
void host_free_any(void *p);
int index_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t id_base, uint64_t *d_out,
                      uint64_t cap, uint64_t *d_count, void *d_scratch, cudaStream_t st);
int value_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint32_t *d_out, uint64_t cap,
                      uint64_t *d_count, void *d_scratch, cudaStream_t st);
int dict_scan_device(uint8_t code_lo, uint8_t code_hi, const int64_t *d_dict, const uint8_t *d_data, size_t n,
                     int64_t *d_out, uint64_t cap, uint64_t *d_count, void *d_scratch, cudaStream_t st);
int explicit_index_scan_device(uint8_t lo, uint8_t hi, const uint64_t *d_index, const uint8_t *d_data, size_t n,
                               uint64_t *d_out, uint64_t cap, uint64_t *d_count, cudaStream_t st);
int scalar_index_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t id_base, uint64_t *d_out,
                             uint64_t cap, uint64_t *d_count, void *d_scratch, cudaStream_t st);
int dict_scan16_device(uint32_t code_lo, uint32_t code_hi, const int64_t *d_dict, const uint16_t *d_data, size_t n,
                       int64_t *d_out, uint64_t cap, uint64_t *d_count, cudaStream_t st);
int dict_scan32_device(uint32_t code_lo, uint32_t code_hi, const int64_t *d_dict, const uint32_t *d_data, size_t n,
                       int64_t *d_out, uint64_t cap, uint64_t *d_count, cudaStream_t st);
int scan_sum_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_sum, cudaStream_t st);
int fill_tiled_column_device(uint8_t *d, size_t n, uint64_t pos_begin, cudaStream_t st);
int fill_skewed_column_device(uint8_t *d, size_t n, uint64_t pos_begin, uint32_t ppm, uint64_t seed, cudaStream_t st);

}  // namespace aqp
