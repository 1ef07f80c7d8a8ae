// scan.cu — range-predicate scans over a packed uint8 column (sm_100a).
//
// Replaces the AVX-512 loops of Scan-Micro-Benchmarks/shared_libraries/SimdScan/src/SIMD512.cpp:
//   count                :7-32     -> scan_count_kernel
//   bitvector_scan       :210-222  -> bitvector_scan_kernel
//   implicit_index_scan  :225-287  -> bitvector_scan_kernel<true> + tile_offsets_kernel + expand_rowids_kernel
// Semantics kept: unsigned inclusive range lo <= v <= hi, only n/64 whole blocks are processed,
// bit k of word i <-> value 64*i+k, row ids ascending uint64 positions.
//
// All three are HBM-bound streaming kernels: 16-byte coalesced loads, 4 independent loads in
// flight per thread, the byte compare done 4 values at a time with carry-free SWAR arithmetic
// (there is no per-byte compare instruction on sm_100; __vcmpgeu4 expands to more ops).
#include <cstring>

#include "common.cuh"
#include "join_internal.cuh"
#include "block_scan.cuh"

namespace aqp {

#ifndef AQP_INDEX_CHUNK_LOG
#define AQP_INDEX_CHUNK_LOG 30
#endif
constexpr int kScanThreads = 256;
constexpr int kScanUnroll = 4;                                  // uint4 loads in flight per thread
constexpr int kScanTileVec = kScanThreads * kScanUnroll;        // 1024 uint4 = 16 KiB per tile
constexpr int kScanTileVals = kScanTileVec * 16;                // 16384 values per tile
constexpr int kExpandWordsPerThread = 4;                        // row-id expansion: 64-bit bitvector words per thread
constexpr int kExpandTileWords = kScanThreads * kExpandWordsPerThread;   // 1024 words
constexpr int kExpandTileVals = kExpandTileWords * 64;          // 65536 values per expansion tile
constexpr int kExpandScanTiles = kExpandTileVals / kScanTileVals;   // = 4 scan tiles per expansion tile

// Range predicate on four packed bytes: bit 7 of every byte of the result is set iff lo <= byte <= hi (unsigned),
// all other bits are zero. sm_100 has no per-byte compare (__vcmpgeu4 expands to more instructions), so this is
// carry-free SWAR arithmetic on the low 7 bits of every byte plus one logic op for the top bits:
//   b = x & 0x7f7f7f7f
//   t = b + (0x80 - lo7)        bit 7 of each byte: low7(x) >= low7(lo)     (no carry leaves a byte: <= 0x7f + 0x80)
//   u = (0x80 | hi7) - b        bit 7 of each byte: low7(x) <= low7(hi)     (no borrow enters a byte)
// and, with x7 the top bit of the byte, the range test depends on the top bits of lo and hi only through WHICH
// three-input function of (x7, t7, u7) it is - one LOP3 each:
//   lo < 128, hi < 128  : ~x7 & t7 & u7        lo < 128 <= hi : x7 ? u7 : t7
//   lo, hi >= 128       :  x7 & t7 & u7        lo >= 128 > hi : empty
// The variant is a launch constant; kernels switch on it outside their inner loops.
//
// Pipe balance (ncu, profiles/r01_ncu_scan_kernels.md, profiles/r02_ncu_scan_fused.md): LOP3/IADD/SHF issue on the
// half-rate ALU pipe, which is what limits these kernels once the loads are coalesced (the round-1 predicate,
// 6 LOP3 + 4 IMAD per word, kept the ALU pipe 62 % busy in the fused row-id kernel). The additions are therefore
// written as multiply-adds (a + b = a * one + b with `one` a runtime 1 the compiler cannot fold) which ptxas places on
// the FMA pipe, and the four flags of a word are gathered into a nibble by ONE dp4a (flag bytes are 0x80 or 0, the
// weights 1,2,4,8 resp. 16,32,64,128 for the odd word: two words accumulate into one 8-bit group at bit 7).
// Per word: 3 LOP3 + 2 IMAD + 1 IDP.4A (was 6 LOP3 + 4 IMAD).
struct Pred {
    uint32_t addK, hiH, one, neg_one, variant;
};
static Pred make_pred(uint8_t lo, uint8_t hi) {
    Pred p;
    p.addK = 0x01010101u * (0x80u - (lo & 0x7fu));
    p.hiH = 0x01010101u * (0x80u | (hi & 0x7fu));
    p.one = 1u;
    p.neg_one = 0xffffffffu;
    p.variant = (uint32_t) (lo >> 7) | ((uint32_t) (hi >> 7) << 1);   // 0: both low, 2: lo low / hi high, 3: both high, 1: empty
    return p;
}

__device__ __forceinline__ uint32_t mad_fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int kVariant>
__device__ __forceinline__ uint32_t inrange_msb(uint32_t x, const Pred &p) {
    const uint32_t H = 0x80808080u;
    const uint32_t b = x & ~H;
    const uint32_t t = mad_fma(b, p.one, p.addK);
    const uint32_t u = mad_fma(b, p.neg_one, p.hiH);
    uint32_t r;
    if (kVariant == 0) r = ~x & t & u;
    else if (kVariant == 2) r = (x & u) | (~x & t);
    else if (kVariant == 3) r = x & t & u;
    else r = 0;
    return r & H;
}

// 16-bit mask of the 16 bytes of v (bit i <-> byte i in memory order)
template <int kVariant>
__device__ __forceinline__ uint32_t range_mask16(uint4 v, const Pred &p) {
    uint32_t g0 = __dp4a(inrange_msb<kVariant>(v.x, p), 0x08040201u, 0u);
    g0 = __dp4a(inrange_msb<kVariant>(v.y, p), 0x80402010u, g0);
    uint32_t g1 = __dp4a(inrange_msb<kVariant>(v.z, p), 0x08040201u, 0u);
    g1 = __dp4a(inrange_msb<kVariant>(v.w, p), 0x80402010u, g1);
    return __umulhi(g1 * 256u + g0, 1u << 25);   // both groups sit at bit 7
}

// run `body(std::integral_constant<int, variant>)` for the launch's predicate variant
#define AQP_PRED_SWITCH(P, BODY)                                        \
    switch ((P).variant) {                                              \
        case 0: { constexpr int kVariant = 0; BODY } break;             \
        case 2: { constexpr int kVariant = 2; BODY } break;             \
        case 3: { constexpr int kVariant = 3; BODY } break;             \
        default: { constexpr int kVariant = 1; BODY } break;            \
    }

// ---------------------------------------------------------------------------------------------
// bitvector scan (kCount: also emit the number of matches of every 16384-value tile)
// ---------------------------------------------------------------------------------------------
// L2 policy for data that is touched once: do not let the streamed column push reusable lines
// (the row-id scan's bitvector scratch) out of the 126 MB L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ld_stream_v4_hint(const uint4 *p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

template <bool kCount, int kVariant>
__device__ __forceinline__ void bitvector_scan_body(const uint4 *__restrict__ in, size_t nvec, uint64_t *__restrict__ out,
                                                    const Pred &p, uint32_t *__restrict__ tile_counts) {
    const uint64_t pol = l2_evict_first_policy();
    const size_t ntiles = (nvec + kScanTileVec - 1) / kScanTileVec;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * kScanTileVec + threadIdx.x;
        uint4 v[kScanUnroll];
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            v[j] = q < nvec ? ld_stream_v4_hint(in + q, pol) : make_uint4(0, 0, 0, 0);
        }
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            uint32_t m16 = range_mask16<kVariant>(v[j], p);
            if (kCount) c += q < nvec ? __popc(m16) : 0;
            // four neighbouring lanes hold the four 16-bit quarters of one output word
            uint32_t m32 = m16 | (__shfl_down_sync(0xffffffffu, m16, 1) << 16);
            uint32_t hi = __shfl_down_sync(0xffffffffu, m32, 2);
            if ((threadIdx.x & 3) == 0 && q < nvec) out[q >> 2] = (uint64_t) m32 | ((uint64_t) hi << 32);
        }
        if (kCount) {   // one reduction-add per warp into the counter of the enclosing expansion tile
            c = __reduce_add_sync(0xffffffffu, c);   // one REDUX instead of a 5-step shuffle tree
            if (lane_id() == 0 && c) atomicAdd(&tile_counts[tile / kExpandScanTiles], c);
        }
    }
}
template <bool kCount>
__global__ void __launch_bounds__(kScanThreads)
bitvector_scan_kernel(const uint4 *__restrict__ in, size_t nvec, uint64_t *__restrict__ out, Pred p,
                      uint32_t *__restrict__ tile_counts) {
    AQP_PRED_SWITCH(p, (bitvector_scan_body<kCount, kVariant>(in, nvec, out, p, tile_counts));)
}

// TMA-fed variant (the north star's "stream the column through TMA bulk loads"): one producer warp keeps a ring of
// kTmaStages stages filled with cp.async.bulk (completion on an mbarrier), kTmaWarps consumer warps take 1 KiB of every
// stage each from shared memory (two conflict-free 128-bit loads per lane), evaluate the predicate and write the words.
// The loads cost the consumers no registers and no address arithmetic and the ring keeps ~60 KiB per CTA in flight
// regardless of what the consumers are doing. A/B against the LDG kernel: csrc/scanbench.cu (count only: 7.08 vs
// 6.66 TB/s) and profiles/r02_sweep_bitvector_tma.txt (this kernel).
constexpr int kTmaWarps = 15, kTmaStages = 4;
constexpr uint32_t kTmaStageBytes = kTmaWarps * 1024;
template <int kVariant>
__device__ __forceinline__ void bitvector_tma_consume(const unsigned char *ring, uint64_t *full, uint64_t *empty, size_t s_begin,
                                                      size_t s_end, uint64_t *__restrict__ out, const Pred &p) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t it = 0;
    for (size_t q = s_begin; q < s_end; ++q, ++it) {
        const uint32_t s = it % kTmaStages, ph = (it / kTmaStages) & 1u;
        mbar_wait(&full[s], ph);
        const uint4 *src = reinterpret_cast<const uint4 *>(ring + (size_t) s * kTmaStageBytes + warp * 1024);
        const uint4 a = src[lane], b = src[lane + 32];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);   // the stage's bytes are in registers
        uint64_t *o = out + (q * kTmaWarps + warp) * 16;   // 1 KiB of values = 16 words
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t m16 = range_mask16<kVariant>(h ? b : a, p);
            const uint32_t m32 = m16 | (__shfl_down_sync(0xffffffffu, m16, 1) << 16);
            const uint32_t hi = __shfl_down_sync(0xffffffffu, m32, 2);
            if ((lane & 3) == 0) o[h * 8 + (lane >> 2)] = (uint64_t) m32 | ((uint64_t) hi << 32);
        }
    }
}
__global__ void __launch_bounds__((kTmaWarps + 1) * 32, 2)
bitvector_scan_tma_kernel(const uint8_t *__restrict__ in, size_t nstages, uint64_t *__restrict__ out, Pred p) {
    extern __shared__ __align__(128) unsigned char tma_ring[];
    __shared__ uint64_t full[kTmaStages], empty[kTmaStages];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTmaWarps);
        }
        mbar_init_fence();
    }
    __syncthreads();
    // the column's stages are dealt to the CTAs in contiguous runs
    const size_t per = (nstages + gridDim.x - 1) / gridDim.x;
    const size_t s_begin = (size_t) blockIdx.x * per, s_end = s_begin + per < nstages ? s_begin + per : nstages;
    if ((threadIdx.x >> 5) == kTmaWarps) {   // producer
        if (lane_id() == 0) {
            uint32_t it = 0;
            for (size_t q = s_begin; q < s_end; ++q, ++it) {
                const uint32_t s = it % kTmaStages, ph = (it / kTmaStages) & 1u;
                if (it >= (uint32_t) kTmaStages) mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], kTmaStageBytes);
                tma_load_1d(tma_ring + (size_t) s * kTmaStageBytes, in + q * kTmaStageBytes, kTmaStageBytes, &full[s]);
            }
        }
        return;
    }
    AQP_PRED_SWITCH(p, (bitvector_tma_consume<kVariant>(tma_ring, full, empty, s_begin, s_end, out, p));)
}

// ---------------------------------------------------------------------------------------------
// count
// ---------------------------------------------------------------------------------------------
template <int kVariant>
__device__ __forceinline__ uint32_t scan_count_body(const uint4 *__restrict__ in, size_t nvec, const Pred &p) {
    const size_t ntiles = (nvec + kScanTileVec - 1) / kScanTileVec;
    uint32_t c = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * kScanTileVec + threadIdx.x;
        uint4 v[kScanUnroll];
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            v[j] = q < nvec ? ld_stream_v4(in + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            uint32_t m16 = range_mask16<kVariant>(v[j], p);
            c += q < nvec ? __popc(m16) : 0;
        }
    }
    return c;
}
__global__ void __launch_bounds__(kScanThreads)
scan_count_kernel(const uint4 *__restrict__ in, size_t nvec, unsigned long long *__restrict__ count, Pred p) {
    uint32_t c = 0;
    AQP_PRED_SWITCH(p, c = scan_count_body<kVariant>(in, nvec, p);)
    __shared__ uint32_t wsum[kScanThreads / 32];
    c = warp_sum(c);
    if (lane_id() == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += wsum[w];
        if (t) atomicAdd(count, t);
    }
}

// ---------------------------------------------------------------------------------------------
// sum of the values in range (SIMD512::sum, SIMD512.cpp:34-86): bytes out of range are masked to zero, the four
// bytes of a word are added with one dp4a
// ---------------------------------------------------------------------------------------------
template <int kVariant>
__device__ __forceinline__ unsigned long long scan_sum_body(const uint4 *__restrict__ in, size_t nvec, const Pred &p) {
    const size_t ntiles = (nvec + kScanTileVec - 1) / kScanTileVec;
    unsigned long long acc = 0;
    auto word_sum = [&](uint32_t x) {
        const uint32_t keep = (inrange_msb<kVariant>(x, p) >> 7) * 0xFFu;   // 0xFF in every byte that is in range
        return __dp4a(x & keep, 0x01010101u, 0u);
    };
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * kScanTileVec + threadIdx.x;
        uint4 v[kScanUnroll];
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            v[j] = q < nvec ? ld_stream_v4(in + q) : make_uint4(0, 0, 0, 0);
        }
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            if (q < nvec) c += word_sum(v[j].x) + word_sum(v[j].y) + word_sum(v[j].z) + word_sum(v[j].w);
        }
        acc += c;
    }
    return acc;
}
__global__ void __launch_bounds__(kScanThreads)
scan_sum_kernel(const uint4 *__restrict__ in, size_t nvec, unsigned long long *__restrict__ sum, Pred p) {
    unsigned long long acc = 0;
    AQP_PRED_SWITCH(p, acc = scan_sum_body<kVariant>(in, nvec, p);)
    __shared__ unsigned long long wsum[kScanThreads / 32];
    acc = warp_sum(acc);
    if (lane_id() == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += wsum[w];
        if (t) atomicAdd(sum, t);
    }
}

// ---------------------------------------------------------------------------------------------
// row-id list scan. The column is processed in chunks whose bitvector (1/8 of the chunk) fits the L2:
//   A  bitvector_scan_kernel<true>   predicate -> bitvector scratch + matches per 16384-value tile
//   B  tile_offsets_kernel           exclusive scan of the tile counts, carried across chunks
//   C  expand_rowids_kernel          bitvector -> ascending uint64 row ids, written warp-cooperatively so
//                                    that every store instruction covers one contiguous run of ids
// The scratch bitvector is rewritten by every chunk and read back while still L2-resident, so HBM
// traffic stays at the algorithmic 1 + 8*selectivity bytes per value and no kernel waits on another
// CTA (a single-pass decoupled look-back variant measured 3-5x slower here: its per-tile chain of
// ticket, load, publish, look back, write kept too few bytes in flight).
// ---------------------------------------------------------------------------------------------
constexpr size_t kIndexChunkVals = (size_t) 1 << AQP_INDEX_CHUNK_LOG;                    // 1 GiB of column -> 128 MiB bitvector
constexpr size_t kIndexChunkTiles = kIndexChunkVals / kExpandTileVals;  // 8192 expansion tiles per chunk

constexpr int kExpandDenseTile = kExpandTileVals / 4 * 3;   // tiles at least this full take the whole-warp path

// Exclusive scan of the per-tile match counts of one chunk (<= 16384 tiles) by ONE CTA, entirely in shared
// memory: coalesced load, 16 consecutive counts per thread (index skewed by i/32 against bank conflicts), warp and
// block scan, coalesced write of the 64-bit tile offsets. The same pass sorts the non-empty tiles into the
// "window" and "dense" lists the two expansion kernels walk (warp-aggregated appends).
// lists[0] = number of window tiles, lists[1] = number of dense tiles, then the two tile-id lists
// (window tiles from lists + 2, dense tiles from lists + 2 + ntiles_cap); empty tiles are in neither.
constexpr uint32_t kPlanMaxTiles = 16384;
constexpr size_t kPlanSmemBytes = (kPlanMaxTiles + kPlanMaxTiles / 32 + 32) * sizeof(uint32_t);
__global__ void __launch_bounds__(kScanBlock)
tile_offsets_kernel(const uint32_t *__restrict__ counts, uint32_t ntiles, uint64_t *__restrict__ offsets,
                    unsigned long long *__restrict__ running, uint32_t *__restrict__ lists, uint32_t ntiles_cap) {
    extern __shared__ uint32_t sc[];
    __shared__ uint32_t wsum[kScanBlock / 32];
    __shared__ uint32_t nclass[2];
    auto at = [](uint32_t i) { return i + (i >> 5); };
    if (threadIdx.x < 2) nclass[threadIdx.x] = 0;
    for (uint32_t i = threadIdx.x; i < ntiles; i += kScanBlock) sc[at(i)] = counts[i];
    __syncthreads();
    const uint32_t per = (ntiles + kScanBlock - 1) / kScanBlock;
    const uint32_t b = threadIdx.x * per, e = min(ntiles, b + per);
    uint32_t local = 0;
    for (uint32_t i = b; i < e; ++i) local += sc[at(i)];
    const uint32_t incl = warp_incl_scan(local);
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t w = wsum[threadIdx.x];
        const uint32_t wi = warp_incl_scan(w);
        wsum[threadIdx.x] = wi - w;
        if (threadIdx.x == 31) sc[at(kPlanMaxTiles)] = wi;   // grand total behind the counts
    }
    __syncthreads();
    uint32_t run = wsum[threadIdx.x >> 5] + incl - local;
    for (uint32_t i = b; i < e; ++i) {   // counts -> exclusive prefixes, in place
        const uint32_t v = sc[at(i)];
        sc[at(i)] = run;
        run += v;
    }
    __syncthreads();
    const unsigned long long base = *running;
    const uint32_t total = sc[at(kPlanMaxTiles)];
    for (uint32_t i0 = 0; i0 < ntiles; i0 += kScanBlock) {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t c = 0;
        if (i < ntiles) {
            offsets[i] = base + sc[at(i)];
            c = counts[i];
        }
#pragma unroll
        for (uint32_t cls = 0; cls < 2; ++cls) {
            const bool in = c != 0 && (c >= (uint32_t) kExpandDenseTile) == (cls == 1);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (m) {
                uint32_t pos = 0;
                if (lane_id() == (unsigned) (__ffs(m) - 1)) pos = atomicAdd(&nclass[cls], (uint32_t) __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1);
                if (in) lists[2 + cls * ntiles_cap + pos + __popc(m & lanemask_lt())] = i;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2) lists[threadIdx.x] = nclass[threadIdx.x];
    if (threadIdx.x == 0) *running = base + total;
}

// A thread loads kExpandWordsPerThread consecutive 64-bit words; the match counts are scanned across the
// CTA, which gives every thread the first output slot of its words. Expansion goes through a shared-memory
// window of kExpandWindow tile-relative ids: each thread walks the set bits of its own words (two 32-bit
// halves per word, FLO + clear-lowest per id) and drops the ids into its slots of the window; then the CTA
// copies the window out front to back, so every store instruction writes 32 consecutive uint64 ids.
// The earlier per-word warp broadcast spent 47 instructions per 27-id word at 10 % selectivity and was
// issue-bound (profiles/r01_ncu_scan_expand_v1.md); it survives only for tiles that are at least 3/4 full,
// where a word yields ~64 ids and the window's extra shared-memory round trip costs more than it saves.
// Slots are XOR-swizzled so that equal per-thread match counts (64 at 100 %) do not map all lanes of a warp
// to one bank.
constexpr int kExpandWindow = 8192;   // ids per window: 32 KiB of shared memory
__device__ __forceinline__ void st_stream_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// What the expansion kernels write for a match at chunk-relative position pos (the remaining SIMD512 variants share
// the row-id machinery, SIMD512.cpp:89-150,:289-336):
//   kEmitRowId  id_base + pos as uint64      implicit_index_scan
//   kEmitValue  the value itself as uint32   scan
//   kEmitDict   dict[value] as int64         dict_scan_8bit_64bit
//   kEmitExplicit  the caller's index entry  explicit_index_scan (SIMD512.cpp:152-208; single-pass kernel only)
enum { kEmitRowId = 0, kEmitValue = 1, kEmitDict = 2, kEmitExplicit = 3 };
struct EmitArgs {
    uint64_t id_base;          // row id of chunk position 0
    const uint8_t *data;       // the chunk of the column (kEmitValue, kEmitDict)
    const int64_t *dict;       // 256 entries (kEmitDict)
    const uint64_t *index;     // kEmitExplicit: the caller's index vectors, 8 entries per 512-bit register
};
template <int kEmit>
__device__ __forceinline__ void emit(void *out, uint64_t slot, uint64_t pos, const EmitArgs &e) {
    if (kEmit == kEmitRowId)
        st_stream_u64(static_cast<uint64_t *>(out) + slot, e.id_base + pos);
    else if (kEmit == kEmitValue)
        st_stream_u32(static_cast<uint32_t *>(out) + slot, __ldg(e.data + pos));
    else
        st_stream_u64(static_cast<uint64_t *>(out) + slot, (uint64_t) __ldg(e.dict + __ldg(e.data + pos)));
}
__device__ __forceinline__ uint32_t expand_phys(uint32_t slot) { return slot ^ ((slot >> 5) & 31u); }

template <int kEmit>
__global__ void __launch_bounds__(kScanThreads, 6)   // 6 CTAs/SM: the 32 KiB window allows no more, so cap the registers
expand_rowids_kernel(const uint64_t *__restrict__ bv, size_t nwords, const uint64_t *__restrict__ tile_offsets,
                     const uint32_t *__restrict__ tile_list, const uint32_t *__restrict__ tile_list_len, EmitArgs ea,
                     void *__restrict__ out, uint64_t out_capacity) {
    __shared__ uint32_t stage[kExpandWindow];
    __shared__ uint32_t wtot[kScanThreads / 32];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const size_t ntiles = (nwords + kExpandTileWords - 1) / kExpandTileWords;
    // the words of the NEXT tile are requested before the current tile is expanded
    auto load_words = [&](size_t tile, uint64_t (&dst)[kExpandWordsPerThread]) {
        const size_t w0 = tile * kExpandTileWords + (size_t) threadIdx.x * kExpandWordsPerThread;
        if (w0 + kExpandWordsPerThread <= nwords) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(bv + w0);
            const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(bv + w0 + 2);
            dst[0] = a.x; dst[1] = a.y; dst[2] = b.x; dst[3] = b.y;
        } else {
#pragma unroll
            for (int k = 0; k < kExpandWordsPerThread; ++k) dst[k] = w0 + k < nwords ? bv[w0 + k] : 0ull;
        }
    };
    // this kernel walks the list of tiles that are less than 3/4 full (built by tile_offsets_kernel); the words
    // of the next list entry are requested before the current tile is expanded
    uint64_t mn[kExpandWordsPerThread];
    const uint32_t nlist = *tile_list_len;
    uint32_t tile_n = 0;
    uint64_t gbase_n = 0;
    if (blockIdx.x < nlist) {
        tile_n = tile_list[blockIdx.x];
        load_words(tile_n, mn);
        gbase_n = tile_offsets[tile_n];
    }
    for (uint32_t li = blockIdx.x; li < nlist; li += gridDim.x) {
        const size_t tile = tile_n;
        const uint64_t gbase = gbase_n;
        uint32_t piece[2 * kExpandWordsPerThread];   // remaining bits, 32 per piece
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < kExpandWordsPerThread; ++k) {
            piece[2 * k] = (uint32_t) mn[k];
            piece[2 * k + 1] = (uint32_t) (mn[k] >> 32);
            cnt += __popcll(mn[k]);
        }
        if (li + gridDim.x < nlist) {
            tile_n = tile_list[li + gridDim.x];
            load_words(tile_n, mn);
            gbase_n = tile_offsets[tile_n];
        }
        const uint32_t incl = warp_incl_scan(cnt);
        __syncthreads();   // wtot and the window of the previous tile are consumed
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < kScanThreads / 32; ++k) {
            wbase += (k < (int) warp) ? wtot[k] : 0;
            total += wtot[k];
        }
        uint32_t slot = wbase + incl - cnt;                                  // next output slot of this thread
        const uint32_t rel0 = threadIdx.x * kExpandWordsPerThread * 64;     // tile-relative id of my first bit
        const uint64_t posb = (uint64_t) tile * kExpandTileVals;   // chunk-relative position of the tile
        const uint64_t idb = ea.id_base + posb;
        for (uint32_t wlo = 0; wlo < total; wlo += kExpandWindow) {
            const uint32_t whi = wlo + kExpandWindow;
#pragma unroll
            for (int q = 0; q < 2 * kExpandWordsPerThread; ++q) {
                uint32_t bits = piece[q];
                while (bits && slot < whi) {
                    const uint32_t b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    stage[expand_phys(slot - wlo)] = rel0 + q * 32 + b;
                    ++slot;
                }
                piece[q] = bits;
            }
            __syncthreads();
            const uint32_t nwin = min((uint32_t) kExpandWindow, total - wlo);
            const uint64_t g0 = gbase + wlo;
            const uint32_t nok = g0 >= out_capacity ? 0u : (uint32_t) min((uint64_t) nwin, out_capacity - g0);
            if (kEmit == kEmitRowId) {   // row ids: one 64-bit add per id
                for (uint32_t t = threadIdx.x; t < nok; t += kScanThreads)
                    st_stream_u64(static_cast<uint64_t *>(out) + g0 + t, idb + stage[expand_phys(t)]);
            } else {
                for (uint32_t t = threadIdx.x; t < nok; t += kScanThreads)
                    emit<kEmit>(out, g0 + t, posb + stage[expand_phys(t)], ea);
            }
            if (whi < total) __syncthreads();   // the window is reused
        }
    }
}

// Tiles that are at least 3/4 full: whole-warp expansion, one word at a time, straight from registers. Lane l
// owns bits l and l+32 of the broadcast word; its slot is the word's offset plus the set bits below it, so each
// store instruction writes one contiguous run (a full 256-byte line pair for a full word).
template <int kEmit>
__global__ void __launch_bounds__(kScanThreads)
expand_dense_rowids_kernel(const uint64_t *__restrict__ bv, size_t nwords, const uint64_t *__restrict__ tile_offsets,
                           const uint32_t *__restrict__ tile_list, const uint32_t *__restrict__ tile_list_len, EmitArgs ea,
                           void *__restrict__ out, uint64_t out_capacity) {
    __shared__ uint32_t wtot[kScanThreads / 32];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    const size_t ntiles = (nwords + kExpandTileWords - 1) / kExpandTileWords;
    auto load_words = [&](size_t tile, uint64_t (&dst)[kExpandWordsPerThread]) {
        const size_t w0 = tile * kExpandTileWords + (size_t) threadIdx.x * kExpandWordsPerThread;
        if (w0 + kExpandWordsPerThread <= nwords) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(bv + w0);
            const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(bv + w0 + 2);
            dst[0] = a.x; dst[1] = a.y; dst[2] = b.x; dst[3] = b.y;
        } else {
#pragma unroll
            for (int k = 0; k < kExpandWordsPerThread; ++k) dst[k] = w0 + k < nwords ? bv[w0 + k] : 0ull;
        }
    };
    // this kernel walks the list of tiles that are at least 3/4 full
    uint64_t mn[kExpandWordsPerThread];
    const uint32_t nlist = *tile_list_len;
    uint32_t tile_n = 0;
    uint64_t gbase_n = 0;
    if (blockIdx.x < nlist) {
        tile_n = tile_list[blockIdx.x];
        load_words(tile_n, mn);
        gbase_n = tile_offsets[tile_n];
    }
    for (uint32_t li = blockIdx.x; li < nlist; li += gridDim.x) {
        const size_t tile = tile_n;
        const uint64_t gbase = gbase_n;
        uint64_t m[kExpandWordsPerThread];
#pragma unroll
        for (int k = 0; k < kExpandWordsPerThread; ++k) m[k] = mn[k];
        if (li + gridDim.x < nlist) {
            tile_n = tile_list[li + gridDim.x];
            load_words(tile_n, mn);
            gbase_n = tile_offsets[tile_n];
        }
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < kExpandWordsPerThread; ++k) cnt += __popcll(m[k]);
        const uint32_t incl = warp_incl_scan(cnt);
        __syncthreads();   // wtot of the previous tile consumed
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int k = 0; k < kScanThreads / 32; ++k) wbase += (k < (int) warp) ? wtot[k] : 0;
        uint32_t off[kExpandWordsPerThread];
        unsigned nz[kExpandWordsPerThread], any = 0;
        {
            uint32_t run = wbase + incl - cnt;
#pragma unroll
            for (int k = 0; k < kExpandWordsPerThread; ++k) {
                off[k] = run;
                run += __popcll(m[k]);
                nz[k] = __ballot_sync(0xffffffffu, m[k] != 0);
                any |= nz[k];
            }
        }
        const uint64_t idw = (kEmit == kEmitRowId ? ea.id_base : 0ull) + (uint64_t) tile * kExpandTileVals +
                             (uint64_t) warp * 32 * kExpandWordsPerThread * 64;
        while (any) {
            const int src = __ffs(any) - 1;
            any &= any - 1;
#pragma unroll
            for (int k = 0; k < kExpandWordsPerThread; ++k) {
                if (!((nz[k] >> src) & 1u)) continue;   // warp-uniform
                const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t) m[k], src);
                const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t) (m[k] >> 32), src);
                const uint64_t o = gbase + __shfl_sync(0xffffffffu, off[k], src);
                const uint64_t id0 = idw + (uint64_t) (src * kExpandWordsPerThread + k) * 64 + lane;
                if ((lo >> lane) & 1u) {
                    uint64_t g = o + __popc(lo & lt);
                    if (g < out_capacity) {
                        if (kEmit == kEmitRowId) st_stream_u64(static_cast<uint64_t *>(out) + g, id0);
                        else emit<kEmit>(out, g, id0, ea);
                    }
                }
                if ((hi >> lane) & 1u) {
                    uint64_t g = o + __popc(lo) + __popc(hi & lt);
                    if (g < out_capacity) {
                        if (kEmit == kEmitRowId) st_stream_u64(static_cast<uint64_t *>(out) + g, id0 + 32);
                        else emit<kEmit>(out, g, id0 + 32, ea);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// single-pass row-id scan (the default): one kernel reads the column once and writes the ids, 1 + 8*sel bytes per
// value of HBM traffic and no bitvector scratch.
//
// A CTA takes macro-tiles in ticket order. Per macro-tile every warp streams its own contiguous sub-range in 1 KiB
// rounds (one 256-bit load per lane: 32 consecutive values -> one 32-bit match mask per lane, kept in shared
// memory), the CTA publishes its match count and learns the number of matches in front of it by decoupled look-back
// over the status words of earlier tiles, then every warp expands ITS OWN rounds on its own: a warp scan of the
// 32 popcounts gives each lane its first slot, the lane walks its set bits into the warp's private 1024-entry
// window (16-bit round-relative positions), and the warp copies the window out front to back, so every store
// instruction writes 32 consecutive ids. No block barrier inside the expansion (the two-pass expand_rowids_kernel
// spent 9 stall cycles per issue at its barriers); three block barriers per macro-tile in total.
// Look-back: status word = flag (2 bits: 0 invalid, 1 tile aggregate, 2 inclusive prefix) | value, one 64-bit
// relaxed store/load each, so flag and value travel together and no fence is needed. One warp reads
// kLookWindows x 32 predecessors per round trip. Macro-tiles are large (>= 64 Ki values) so that tiles are
// created more slowly than the prefix front can move (a first version with 16 Ki-value tiles was 3-5x slower
// than two passes for exactly that reason).
// ---------------------------------------------------------------------------------------------
struct U32x8 {
    uint32_t w[8];
};
__device__ __forceinline__ U32x8 ld_stream_v8_hint(const void *p, uint64_t pol) {   // SASS: LDG.E.256
    U32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p), "l"(pol));
    return r;
}
// 32-bit mask of a lane's 32 consecutive bytes: four 8-bit dp4a groups (each at bit 7) put together on the FMA pipe
template <int kVariant>
__device__ __forceinline__ uint32_t range_mask32(const U32x8 &v, const Pred &p) {
    uint32_t g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        g[k] = __dp4a(inrange_msb<kVariant>(v.w[2 * k], p), 0x08040201u, 0u);
        g[k] = __dp4a(inrange_msb<kVariant>(v.w[2 * k + 1], p), 0x80402010u, g[k]);
    }
    const uint32_t hi = g[3] * (1u << 17) + g[2] * (1u << 9);   // groups 2, 3 -> bits 16..31
    return __umulhi(g[1] * 256u + g[0], 1u << 25) + hi;        // groups 0, 1 -> bits 0..15
}

// status word of a tile: flag (2 bits: 1 = the tile's own count, 2 = inclusive prefix) | epoch (22 bits) | value (40 bits).
// A word counts only if its epoch is the current launch's: the status array is zeroed once when it is allocated and
// never again (a memset node in front of every scan cost ~2 us, 10 % of a 128 MiB shard's scan).
constexpr int kStatValueBits = 40, kStatEpochBits = 22;
constexpr uint64_t kStatAggregate = 1ull << 62, kStatInclusive = 2ull << 62, kStatValueMask = (1ull << kStatValueBits) - 1;
constexpr uint32_t kStatEpochMask = (1u << kStatEpochBits) - 1;
constexpr int kLookWindows = 4;   // status words per lane and round trip of the look-back
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// number of matches in tiles [0, tile): called by one whole warp, tile > 0
__device__ __forceinline__ uint64_t lookback_exclusive(const unsigned long long *status, uint32_t tile, uint32_t epoch) {
    const unsigned lane = lane_id();
    uint64_t sum = 0;
    int64_t d = (int64_t) tile - 1;   // index of the nearest status word not yet accounted for
    for (;;) {
        unsigned long long s[kLookWindows];
#pragma unroll
        for (int k = 0; k < kLookWindows; ++k) {
            const int64_t idx = d - k * 32 - (int64_t) lane;
            s[k] = idx >= 0 ? ld_relaxed_u64(status + idx) : (kStatInclusive | ((uint64_t) epoch << kStatValueBits));   // in front of tile 0: prefix 0
        }
        // windows nearest first; state 0 = all 32 were aggregates (go on), 1 = reached an inclusive prefix (done),
        // 2 = met a tile that has not published yet (poll again from there)
        int state = 0;
        auto window = [&](const unsigned long long sk, const int k) {
            const uint32_t flag = ((uint32_t) (sk >> kStatValueBits) & kStatEpochMask) == epoch ? (uint32_t) (sk >> 62) : 0u;   // stale epoch: not published
            const unsigned inval = __ballot_sync(0xffffffffu, flag == 0);
            const unsigned incl = __ballot_sync(0xffffffffu, flag == 2);
            const unsigned first_inval = inval ? (unsigned) __ffs(inval) - 1 : 32u;
            const unsigned first_incl = incl ? (unsigned) __ffs(incl) - 1 : 32u;
            const unsigned upto = min(first_inval, first_incl);   // aggregates nearer than this lane count
            const uint32_t val32 = (uint32_t) sk;                 // an aggregate is at most one tile's worth of matches (< 2^32)
            sum += __reduce_add_sync(0xffffffffu, lane < upto ? val32 : 0u);
            if (first_incl < first_inval) {
                const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t) sk, first_incl);
                const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t) (sk >> 32), first_incl);
                sum += (((uint64_t) hi << 32) | lo) & kStatValueMask;
                state = 1;
            } else if (first_inval < 32u) {
                d -= k * 32 + (int64_t) first_inval;
                state = 2;
            }
        };
        window(s[0], 0);
        if (state == 0) window(s[1], 1);
        if (state == 0) window(s[2], 2);
        if (state == 0) window(s[3], 3);
        static_assert(kLookWindows == 4, "unrolled by hand");
        if (state == 1) return sum;
        if (state == 0) d -= kLookWindows * 32;
        else __nanosleep(100);   // a predecessor is still streaming: do not hammer the L2 while it finishes
    }
}

constexpr size_t kFusedMinTileVals = 7 * 4 * 1024;   // smallest macro-tile any geometry uses (sizes the status array)
constexpr int kFusedRoundVals = 1024;   // values per warp round: 32 lanes x 32 bytes
constexpr int kFusedBatch = 4;          // 256-bit loads in flight per lane
// Where a group of matches goes: the group's first output slot and the value position its relative offsets count
// from are folded into two 64-bit bases once per group, so that a match costs one 32-bit multiply-add for the
// address (IMAD.WIDE) and one 64-bit add for the id.
template <int kEmit>
struct GroupOut {
    unsigned char *o;        // address of output slot g
    uint64_t idb;            // kEmitRowId: id of relative position 0
    const uint8_t *dp;       // kEmitValue / kEmitDict: the column at relative position 0
    const int64_t *dict;
    const uint64_t *ib;      // kEmitExplicit: index vector of the 64-value block at relative position 0
    __device__ __forceinline__ GroupOut(void *out, uint64_t g, uint64_t pos0, const EmitArgs &ea)
        : o(static_cast<unsigned char *>(out) + g * (kEmit == kEmitValue ? 4 : 8)), idb(ea.id_base + pos0), dp(ea.data + pos0),
          dict(ea.dict), ib(kEmit == kEmitExplicit ? ea.index + (pos0 >> 6) * 8 : nullptr) {}   // pos0 is a multiple of 64
    __device__ __forceinline__ void put(uint32_t t, uint32_t rel) const {   // match at relative position rel -> slot g + t
        if (kEmit == kEmitRowId) st_stream_u64(reinterpret_cast<uint64_t *>(o) + t, idb + rel);
        else if (kEmit == kEmitValue) st_stream_u32(reinterpret_cast<uint32_t *>(o) + t, __ldg(dp + rel));
        else if (kEmit == kEmitDict) st_stream_u64(reinterpret_cast<uint64_t *>(o) + t, (uint64_t) __ldg(dict + __ldg(dp + rel)));
        else   // block i = pos / 64, byte group j = (pos / 8) % 8, lane k = pos % 8 -> index_compressed[i + j] lane k (:176,:198)
            st_stream_u64(reinterpret_cast<uint64_t *>(o) + t, __ldg(ib + ((((rel >> 6) + ((rel >> 3) & 7u)) << 3) | (rel & 7u))));
    }
};

// mask words of one warp and tile: entry i (= round * 32 + lane) covers values [32 i, 32 i + 32) of the warp's
// sub-range; stored at i + i / 32 so that both access patterns are conflict-free: the streaming phase writes entry
// r * 32 + lane, the expansion reads G consecutive entries per lane (G = 1, 2, 4, 8, 16)
__device__ __forceinline__ uint32_t mask_slot(uint32_t i) { return i + (i >> 5); }
static size_t fused_smem_bytes(int warps, int rounds) {   // per worker warp: two mask buffers + one window
    return (size_t) (warps - 1) * (2 * rounds * 33 * sizeof(uint32_t) + kFusedRoundVals * sizeof(uint16_t));
}

// CTA = kWarps - 1 worker warps + 1 control warp, NO block barrier after start-up; the warps meet through mbarriers
// that are normally complete long before anyone waits on them:
//   control  iteration i: take the ticket of tile T[i+1] (-> bar_ticket), look back for T[i-1], whose count it
//            published an iteration ago (-> s_prefix, bar_prefix), wait for the workers' counts of T[i]
//            (bar_counts) and publish their sum to later tiles
//   worker   iteration i: stream its sub-range of T[i] (predicate -> masks in shared memory, count -> wtot,
//            arrive on bar_counts), THEN expand its sub-range of T[i-1] from the other mask buffer
// so a look-back has a whole tile's streaming time to finish (the first, unpipelined version lost 6-9 us per tile
// waiting for the slowest of its ~50 nearest predecessors; with the look-back in a worker warp between two block
// barriers 18 % of the stall samples of the 0.1 % case were barrier waits). Workers may drift apart by one tile.
// Slots (ticket, counts, prefix) live in a ring of kFusedStages; the control warp cannot run more than two
// iterations ahead of the slowest worker (it waits for all counts of T[i]), so four stages never alias.
//
// Streaming keeps kFusedBatch 256-bit loads per lane in flight on a rolling basis (a register is reloaded for
// round r+4 as soon as round r's mask is computed).
//
// Expansion of a warp's sub-range (kRounds * 1024 values, W matches): rounds are taken in groups of G = 16, 8, 4, 2
// or 1 - the largest group whose matches fit the warp's 1024-entry window - and lane l owns the G consecutive mask
// words l*G .. l*G+G-1 of the group, i.e. 32*G consecutive values: one popcount sum and ONE warp scan per group give
// every lane its first slot, the lane walks the set bits of its non-empty words into the window, the warp copies
// the window out in order. Sparse data costs ~150 instructions per 16 Ki values this way (a scan per 1 KiB round
// cost 1100 and made the 0.1 % case instruction-bound); clustered matches (the reference's tiled column: hi+1
// consecutive matches in every 256 values) are balanced whenever a lane covers a multiple of 256 values. For G = 1
// and very uneven lanes the whole warp takes one lane's word at a time instead (lane l tests bit l) and stores
// straight to global memory: a full word leaves as one 256-byte store.
constexpr int kFusedStages = 4;
constexpr int fused_ctas_per_sm(int warps) { return 1056 / (warps * 32); }   // ~1024 threads per SM: 60+ registers each
template <int kEmit, int kWarps, int kRounds, bool kV8>
__global__ void __launch_bounds__(kWarps * 32, fused_ctas_per_sm(kWarps))
rowid_scan_fused_kernel(const uint8_t *__restrict__ in, size_t n, Pred p, EmitArgs ea, void *__restrict__ out,
                        uint64_t out_capacity, unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket,
                        unsigned long long *__restrict__ d_count, uint32_t ntiles, uint32_t epoch) {
    constexpr int kWorkers = kWarps - 1;
    constexpr int kWarpVals = kRounds * kFusedRoundVals;
    constexpr size_t kTileVals = (size_t) kWorkers * kWarpVals;
    constexpr int kWarpWords = kRounds * 33;             // padded mask words per warp and buffer
    static_assert(kRounds % kFusedBatch == 0, "whole load batches");
    static_assert(kRounds <= 16, "a lane owns at most 16 mask words of a group");
    extern __shared__ __align__(16) unsigned char fused_smem[];
    __shared__ uint64_t bar_ticket[kFusedStages], bar_counts[kFusedStages], bar_prefix[kFusedStages];
    __shared__ uint32_t s_tile[kFusedStages], want[kFusedStages];
    __shared__ uint32_t wtot[kFusedStages][kWorkers];
    __shared__ unsigned long long s_prefix[kFusedStages];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kFusedStages; ++s) {
            mbar_init(&bar_ticket[s], 1);
            mbar_init(&bar_counts[s], kWorkers);
            mbar_init(&bar_prefix[s], 1);
            want[s] = 0;
        }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == (unsigned) kWorkers) {
        // ------------------------------------------------------------------ control warp
        uint32_t t_prev = 0, total_prev = 0;
        const uint64_t etag = (uint64_t) epoch << kStatValueBits;
        for (uint32_t i = 0;; ++i) {
            const uint32_t s = i % kFusedStages, ph = (i / kFusedStages) & 1u;
            if (i > 0) {            // matches in front of T[i-1], whose count went out at the end of the last iteration
                uint64_t prefix = 0;
                if (t_prev > 0) {
                    prefix = lookback_exclusive(status, t_prev, epoch);
                    if (lane == 0) st_relaxed_u64(status + t_prev, kStatInclusive | etag | (prefix + total_prev));
                }
                if (lane == 0) {
                    s_prefix[(i - 1) % kFusedStages] = prefix;
                    if (t_prev == ntiles - 1) *d_count = prefix + total_prev;
                    mbar_arrive(&bar_prefix[(i - 1) % kFusedStages]);
                }
            }
            mbar_wait(&bar_ticket[s], ph);   // the first worker that got this far took the ticket of T[i]
            const uint32_t t_cur = s_tile[s];
            if (t_cur >= ntiles) break;
            mbar_wait(&bar_counts[s], ph);   // every worker has streamed T[i]
            uint32_t total = 0;
            for (int k = lane; k < kWorkers; k += 32) total += wtot[s][k];
            total = __reduce_add_sync(0xffffffffu, total);
            if (lane == 0) st_relaxed_u64(status + t_cur, (t_cur == 0 ? kStatInclusive : kStatAggregate) | etag | total);
            t_prev = t_cur;
            total_prev = total;
        }
        // this CTA saw a ticket past the end and will not take another: the last CTA to get here re-arms the counters
        if (lane == 0) {
            __threadfence();
            if (atomicAdd(ticket + 1, 1u) == gridDim.x - 1) {
                ticket[0] = 0;
                ticket[1] = 0;
                __threadfence();
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- worker warps
    const unsigned lt = lanemask_lt();
    uint32_t *my_masks = reinterpret_cast<uint32_t *>(fused_smem) + (size_t) warp * 2 * kWarpWords;   // [buffer][padded entry]
    uint16_t *win = reinterpret_cast<uint16_t *>(reinterpret_cast<uint32_t *>(fused_smem) + (size_t) kWorkers * 2 * kWarpWords) +
                    (size_t) warp * kFusedRoundVals;
    const uint64_t pol = l2_evict_first_policy();
    auto load = [&](const uint8_t *q) {
        if (kV8) return ld_stream_v8_hint(q, pol);
        const uint4 a = ld_stream_v4_hint(reinterpret_cast<const uint4 *>(q), pol);
        const uint4 c = ld_stream_v4_hint(reinterpret_cast<const uint4 *>(q) + 1, pol);
        return U32x8{{a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w}};
    };

    uint32_t prev_tile = 0;
    for (uint32_t i = 0;; ++i) {
        const uint32_t s = i % kFusedStages, ph = (i / kFusedStages) & 1u;
        // the ticket of T[i] is taken by the first worker that needs it, i.e. as late as possible: tiles then start
        // streaming in ticket order and a look-back rarely meets a tile that is still being streamed (with tickets
        // taken an iteration ahead by the control warp, workers of the 50 % case spent 20 % of their instructions
        // spinning on bar_prefix: a ticket was held for two expansions before its tile was read)
        if (lane == 0) {
            const uint32_t a = atomicAdd(&want[s], 1u);
            if (a == 0) {
                s_tile[s] = atomicAdd(ticket, 1u);
                mbar_arrive(&bar_ticket[s]);
            }
            if (a == (uint32_t) kWorkers - 1) want[s] = 0;   // free for iteration i + kFusedStages
        }
        mbar_wait(&bar_ticket[s], ph);
        const uint32_t tile = s_tile[s];

        // ---- stream my sub-range of T[i]: predicate -> one 32-bit mask per lane and round
        if (tile < ntiles) {
            const size_t wbase = (size_t) tile * kTileVals + (size_t) warp * kWarpVals;
            uint32_t *wm = my_masks + (i & 1u) * kWarpWords + lane;   // entry r*32+lane -> + r*33
            const uint8_t *src = in + wbase + lane * 32;
            uint32_t cnt = 0;
            if (wbase + kWarpVals <= n) {   // whole sub-range inside the column: no per-load bounds
                U32x8 v[kFusedBatch];
#pragma unroll
                for (int j = 0; j < kFusedBatch; ++j) v[j] = load(src + (size_t) j * kFusedRoundVals);
                AQP_PRED_SWITCH(p,
                    _Pragma("unroll 1")
                    for (int b = 0; b < kRounds; b += kFusedBatch) {
                        _Pragma("unroll")
                        for (int j = 0; j < kFusedBatch; ++j) {
                            const uint32_t m = range_mask32<kVariant>(v[j], p);
                            if (b + kFusedBatch < kRounds) v[j] = load(src + (size_t) (b + kFusedBatch + j) * kFusedRoundVals);
                            wm[(b + j) * 33] = m;
                            cnt += __popc(m);
                        }
                    })
            } else {   // the column's last tile: a lane's 32 bytes are inside or outside as a whole (n % 64 == 0)
                AQP_PRED_SWITCH(p,
                    _Pragma("unroll 1")
                    for (int r = 0; r < kRounds; ++r) {
                        const size_t pos = wbase + (size_t) r * kFusedRoundVals + lane * 32;
                        uint32_t m = 0;
                        if (pos < n) m = range_mask32<kVariant>(load(in + pos), p);
                        wm[r * 33] = m;
                        cnt += __popc(m);
                    })
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) {
                wtot[s][warp] = cnt;
                mbar_arrive(&bar_counts[s]);   // release: the control warp and the other workers see wtot after their wait
            }
        }

        // ---- expand my sub-range of T[i-1]
        if (i > 0) {
            const uint32_t sp = (i - 1) % kFusedStages;
            mbar_wait(&bar_prefix[sp], ((i - 1) / kFusedStages) & 1u);   // also: every worker's count of T[i-1] is in wtot
            uint32_t before = 0;
#pragma unroll
            for (int k = 0; k < kWorkers; ++k) before += (k < (int) warp) ? wtot[sp][k] : 0u;
            const uint32_t wcount = wtot[sp][warp];
            uint64_t g = s_prefix[sp] + before;   // output slot of this warp's next id
            if (wcount != 0 && g < out_capacity) {
                const size_t wbase = (size_t) prev_tile * kTileVals + (size_t) warp * kWarpVals;
                const uint32_t *wm = my_masks + ((i - 1) & 1u) * kWarpWords;
                // largest group whose expected matches fit the window with some slack
                uint32_t G = kRounds;
                while (G > 1 && wcount * G > 896u * kRounds) G >>= 1;
                uint32_t r = 0;
                while (r < (uint32_t) kRounds) {
                    const uint32_t e0 = r * 32 + G * lane;   // my first entry of this group
                    uint32_t c = 0, nzk = 0;
                    for (uint32_t k = 0; k < G; ++k) {
                        const uint32_t m = wm[mask_slot(e0 + k)];
                        c += __popc(m);
                        nzk |= (m != 0 ? 1u : 0u) << k;
                    }
                    const uint32_t incl = warp_incl_scan(c);
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total > (uint32_t) kFusedRoundVals) {   // denser than the sub-range's average: smaller groups
                        G >>= 1;
                        continue;
                    }
                    const GroupOut<kEmit> go(out, g, wbase + (size_t) r * kFusedRoundVals, ea);   // slot and position of the group
                    const bool fits = g + total <= out_capacity;
                    const uint32_t room = fits ? 0xffffffffu : (uint32_t) (out_capacity - g);   // g < out_capacity here
                    if (total == (uint32_t) kFusedRoundVals && G == 1) {   // a full round needs no window
                        if (fits) {
#pragma unroll 8
                            for (uint32_t t = lane; t < (uint32_t) kFusedRoundVals; t += 32) go.put(t, t);
                        } else {
                            for (uint32_t t = lane; t < (uint32_t) kFusedRoundVals; t += 32)
                                if (t < room) go.put(t, t);
                        }
                    } else if (total) {
                        const uint32_t excl = incl - c;
                        bool by_word = false;
                        if (G == 1) {
                            const uint32_t maxc = __reduce_max_sync(0xffffffffu, c);
                            const uint32_t nz = __popc(__ballot_sync(0xffffffffu, c != 0));
                            by_word = 8u * nz < 9u * maxc + 7u * ((total + 31) >> 5) + 10u;
                        }
                        if (by_word) {   // one lane's word at a time, lane l tests bit l, straight to global memory
                            const uint32_t m = wm[mask_slot(e0)];
                            for (unsigned mm = __ballot_sync(0xffffffffu, m != 0); mm; mm &= mm - 1) {
                                const int src = __ffs(mm) - 1;
                                const uint32_t word = __shfl_sync(0xffffffffu, m, src);
                                const uint32_t off = __shfl_sync(0xffffffffu, excl, src);
                                if (word == 0xffffffffu) {   // 32 consecutive ids: one 256-byte store, no popcount
                                    if (off + lane < room) go.put(off + lane, src * 32 + lane);
                                } else if ((word >> lane) & 1u) {
                                    const uint32_t t = off + __popc(word & lt);
                                    if (t < room) go.put(t, src * 32 + lane);
                                }
                            }
                        } else {         // every lane walks the set bits of its own non-empty words
                            uint32_t slot = excl, bits = 0, vbase = 0;
                            for (;;) {
                                if (bits == 0) {
                                    if (nzk == 0) break;
                                    const uint32_t k = __ffs(nzk) - 1;
                                    nzk &= nzk - 1;
                                    bits = wm[mask_slot(e0 + k)];
                                    vbase = (G * lane + k) * 32;
                                    continue;
                                }
                                win[slot++] = (uint16_t) (vbase + __ffs(bits) - 1);
                                bits &= bits - 1;
                            }
                            __syncwarp();
                            if (fits) {
                                uint32_t t = lane;
                                for (; t + 96 < total; t += 128) {
                                    const uint32_t w0 = win[t], w1 = win[t + 32], w2 = win[t + 64], w3 = win[t + 96];
                                    go.put(t, w0);
                                    go.put(t + 32, w1);
                                    go.put(t + 64, w2);
                                    go.put(t + 96, w3);
                                }
                                for (; t < total; t += 32) go.put(t, win[t]);
                            } else {
                                for (uint32_t t = lane; t < total; t += 32)
                                    if (t < room) go.put(t, win[t]);
                            }
                            __syncwarp();
                        }
                    }
                    g += total;
                    r += G;
                }
            }
        }
        if (tile >= ntiles) break;
        prev_tile = tile;
    }
}

// ---------------------------------------------------------------------------------------------
// column generators (Scan-Micro-Benchmarks/shared_libraries/SharedHeaders/include/Allocator.hpp:95-109)
// ---------------------------------------------------------------------------------------------
__global__ void fill_tiled_kernel(uint8_t *data, size_t n, uint64_t pos_begin) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) data[i] = (uint8_t) ((pos_begin + i) & 255u);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void fill_skewed_kernel(uint8_t *data, size_t n, uint64_t pos_begin, uint32_t p_zero_ppm, uint64_t seed) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t h = splitmix64(seed ^ splitmix64(pos_begin + i));
        uint32_t r = (uint32_t) (h % 1000000u);
        uint8_t v = (uint8_t) (1u + (uint32_t) ((h >> 32) % 255u));
        data[i] = r < p_zero_ppm ? (uint8_t) 0 : v;
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers (device-pointer API)
// ---------------------------------------------------------------------------------------------
static int scan_grid(size_t nvec, int blocks_per_sm) {
    size_t ntiles = (nvec + kScanTileVec - 1) / kScanTileVec;
    size_t g = (size_t) kNumSMs * blocks_per_sm;
    if (ntiles < g) g = ntiles;
    return (int) (g ? g : 1);
}

int bitvector_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_out,
                          cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) != 0) {
        set_error("bitvector_scan: column must be 16-byte aligned");
        return -1;
    }
    size_t nvec = (n / 64) * 4;
    if (nvec == 0) return 0;
    // large columns: the TMA-fed kernel over whole 15 KiB stages, the LDG kernel for the rest (B200_AQP_BITVECTOR=ldg: A/B)
    static const bool ldg_only = getenv("B200_AQP_BITVECTOR") && !strcmp(getenv("B200_AQP_BITVECTOR"), "ldg");
    size_t done = 0;
    const size_t nstages = (n / 64 * 64) / kTmaStageBytes;
    // measured (profiles/r02_sweep_bitvector_tma.txt): 2^30 values 0.182 vs 0.202 ms, 2^29 0.096 vs 0.099, below that the
    // LDG kernel wins (a column of <= 2^28 bytes is partly L2-resident between runs, and the ring's fill / drain shows)
    if (!ldg_only && (n >> 29) != 0) {
        static unsigned attr_set = ~0u;
        if (attr_set != g_device_epoch) {
            AQP_CUDA_OK(cudaFuncSetAttribute(bitvector_scan_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) (kTmaStages * kTmaStageBytes)));
            attr_set = g_device_epoch;
        }
        bitvector_scan_tma_kernel<<<2 * kNumSMs, (kTmaWarps + 1) * 32, kTmaStages * kTmaStageBytes, st>>>(d_data, nstages, d_out,
                                                                                                        make_pred(lo, hi));
        AQP_LAUNCHED();
        done = nstages * kTmaStageBytes;
    }
    if (done < n / 64 * 64) {
        const size_t rest = (n / 64 * 64 - done) / 16;
        bitvector_scan_kernel<false><<<scan_grid(rest, 8), kScanThreads, 0, st>>>(reinterpret_cast<const uint4 *>(d_data + done),
                                                                                 rest, d_out + done / 64, make_pred(lo, hi), nullptr);
        AQP_LAUNCHED();
    }
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int scan_count_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_count, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) != 0) {
        set_error("scan_count: column must be 16-byte aligned");
        return -1;
    }
    AQP_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    size_t nvec = (n / 64) * 4;
    if (nvec == 0) return 0;
    scan_count_kernel<<<scan_grid(nvec, 8), kScanThreads, 0, st>>>(reinterpret_cast<const uint4 *>(d_data), nvec,
                                                                  reinterpret_cast<unsigned long long *>(d_count),
                                                                  make_pred(lo, hi));
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// scratch layout: [bitvector of one chunk][tile counts u32][tile offsets u64][tile lists: 2 lengths + 2 lists u32]
static size_t align256(size_t x) { return (x + 255) & ~(size_t) 255; }
static bool index_scan_two_pass() {
    static const bool two_pass = [] {
        const char *e = getenv("B200_AQP_INDEX_SCAN");   // A/B switch: "twopass" = bitvector scratch + expansion kernels
        return e && !strcmp(e, "twopass");
    }();
    return two_pass;
}
size_t index_scan_scratch_bytes(size_t n) {
    if (!index_scan_two_pass()) return 256;   // the single-pass kernel keeps its few KiB of state in its own buffer
    size_t chunk = n < kIndexChunkVals ? n : kIndexChunkVals;
    size_t tiles = (chunk + kExpandTileVals - 1) / kExpandTileVals + 1;
    return align256(chunk / 8 + 64) + align256(tiles * 4) + align256(tiles * 8) + align256((2 * tiles + 2) * 4) + 256;   // two-pass path only
}

// ---- single-pass path: geometry (warps per CTA x rounds per warp) picked by size, B200_AQP_SCAN_GEOM=WxR overrides
static struct {
    DevBuf buf;
    uint32_t epoch = 0;
} g_fused;
void scan_release_codes();
void scan_release() {   // b200_shutdown: the buffers belong to the device that is being left
    g_fused.buf.release();
    g_fused.epoch = 0;
    scan_release_codes();
}
template <int kEmit, int kWarps, int kRounds>
static int launch_fused(const uint8_t *d_data, size_t n, const Pred &p, const EmitArgs &ea, void *d_out, uint64_t cap,
                        uint64_t *d_count, void *d_scratch, cudaStream_t st) {
    constexpr size_t kTileVals = (size_t) (kWarps - 1) * kRounds * kFusedRoundVals;   // one warp of the CTA is the control warp
    static_assert(kTileVals >= kFusedMinTileVals, "scratch is sized for tiles of at least kFusedMinTileVals");
    const uint32_t ntiles = (uint32_t) ((n + kTileVals - 1) / kTileVals);
    // ticket + exit counter (256 bytes) and one status word per macro-tile live in a buffer this file owns: zeroed when
    // it is (re)allocated, tagged with a per-launch epoch afterwards, counters re-armed by the kernel's last CTA
    const size_t need = 256 + (size_t) ntiles * 8;
    if (need > g_fused.buf.cap || g_fused.epoch >= kStatEpochMask) {
        if (g_fused.buf.ensure(need)) return -1;
        AQP_CUDA_OK(cudaMemsetAsync(g_fused.buf.p, 0, g_fused.buf.cap, st));
        g_fused.epoch = 0;
    }
    const uint32_t epoch = ++g_fused.epoch;
    unsigned int *ticket = static_cast<unsigned int *>(g_fused.buf.p);
    unsigned long long *status = reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(g_fused.buf.p) + 256);
    const size_t smem = fused_smem_bytes(kWarps, kRounds);
    const bool v8 = (reinterpret_cast<uintptr_t>(d_data) & 31u) == 0;
    auto kern = v8 ? rowid_scan_fused_kernel<kEmit, kWarps, kRounds, true> : rowid_scan_fused_kernel<kEmit, kWarps, kRounds, false>;
    AQP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));   // per call: cheap, and right after a device change
    const uint32_t resident = (uint32_t) kNumSMs * fused_ctas_per_sm(kWarps);
    kern<<<ntiles < resident ? ntiles : resident, kWarps * 32, smem, st>>>(
        d_data, n, p, ea, d_out, cap, status, ticket, reinterpret_cast<unsigned long long *>(d_count), ntiles, epoch);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int kEmit>
static int fused_scan_device(const uint8_t *d_data, size_t n, const Pred &p, const EmitArgs &ea, void *d_out, uint64_t cap,
                             uint64_t *d_count, void *d_scratch, cudaStream_t st) {
    // measured (profiles/r02_sweep_scan_v7*.txt): 15 workers x 16 KiB tiles win from 2^26 values up, half-size tiles below
    int w = 16, r = n >= ((size_t) 1 << 26) ? 16 : 8;
    if (const char *e = getenv("B200_AQP_SCAN_GEOM")) sscanf(e, "%dx%d", &w, &r);
#define AQP_GEOM(W, R) \
    if (w == W && r == R) return launch_fused<kEmit, W, R>(d_data, n, p, ea, d_out, cap, d_count, d_scratch, st)
    AQP_GEOM(16, 16);
    AQP_GEOM(16, 8);
    AQP_GEOM(16, 4);
    AQP_GEOM(8, 8);
    AQP_GEOM(8, 16);
    AQP_GEOM(32, 8);
    AQP_GEOM(11, 16);
    AQP_GEOM(6, 16);
    AQP_GEOM(11, 8);
#undef AQP_GEOM
    set_error("B200_AQP_SCAN_GEOM: unknown geometry");
    return -1;
}

// emit = kEmitRowId / kEmitValue / kEmitDict (d_dict: 256 int64 on the device); d_out holds cap elements of the
// emitted type
static int emit_scan_device(int emit_kind, uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t id_base,
                            const int64_t *d_dict, void *d_out, uint64_t cap, uint64_t *d_count, void *d_scratch,
                            cudaStream_t st, const uint64_t *d_index = nullptr) {
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) != 0) {
        set_error("scan: column must be 16-byte aligned");
        return -1;
    }
    n = n / 64 * 64;
    if (n == 0) {
        AQP_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
        return 0;
    }
    if (!index_scan_two_pass()) {
        const Pred pf = make_pred(lo, hi);
        const EmitArgs eaf{id_base, d_data, d_dict, d_index};
        if (emit_kind == kEmitRowId) return fused_scan_device<kEmitRowId>(d_data, n, pf, eaf, d_out, cap, d_count, d_scratch, st);
        if (emit_kind == kEmitValue) return fused_scan_device<kEmitValue>(d_data, n, pf, eaf, d_out, cap, d_count, d_scratch, st);
        if (emit_kind == kEmitExplicit) return fused_scan_device<kEmitExplicit>(d_data, n, pf, eaf, d_out, cap, d_count, d_scratch, st);
        return fused_scan_device<kEmitDict>(d_data, n, pf, eaf, d_out, cap, d_count, d_scratch, st);   // writes *d_count itself
    }
    if (emit_kind == kEmitExplicit) {
        set_error("explicit_index_scan is served by the single-pass kernel only (unset B200_AQP_INDEX_SCAN=twopass)");
        return -1;
    }
    AQP_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    const size_t chunk_cap = n < kIndexChunkVals ? n : kIndexChunkVals;
    const size_t tiles_cap = (chunk_cap + kExpandTileVals - 1) / kExpandTileVals + 1;
    unsigned char *sb = static_cast<unsigned char *>(d_scratch);
    uint64_t *bv = reinterpret_cast<uint64_t *>(sb);
    uint32_t *counts = reinterpret_cast<uint32_t *>(sb + align256(chunk_cap / 8 + 64));
    uint64_t *offsets = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(counts) + align256(tiles_cap * 4));
    uint32_t *lists = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(offsets) + align256(tiles_cap * 8));
    unsigned long long *running = reinterpret_cast<unsigned long long *>(d_count);
    const Pred p = make_pred(lo, hi);
    static_assert(kIndexChunkVals / kExpandTileVals <= kPlanMaxTiles, "one planning CTA per chunk");
    static unsigned attr_set = ~0u;   // device epoch the opt-in was made for (b200_shutdown + b200_init(other device) re-arms it)
    if (attr_set != g_device_epoch) {
        AQP_CUDA_OK(cudaFuncSetAttribute(tile_offsets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) kPlanSmemBytes));
        attr_set = g_device_epoch;
    }
    for (size_t begin = 0; begin < n; begin += kIndexChunkVals) {
        const size_t len = n - begin < kIndexChunkVals ? n - begin : kIndexChunkVals;
        const size_t nvec = len / 16, nwords = len / 64;
        const uint32_t ntiles = (uint32_t) ((nwords + kExpandTileWords - 1) / kExpandTileWords);
        AQP_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t) ntiles * 4, st));
        bitvector_scan_kernel<true><<<scan_grid(nvec, 8), kScanThreads, 0, st>>>(
            reinterpret_cast<const uint4 *>(d_data + begin), nvec, bv, p, counts);
        AQP_LAUNCHED();
        tile_offsets_kernel<<<1, kScanBlock, kPlanSmemBytes, st>>>(counts, ntiles, offsets, running, lists,
                                                                   (uint32_t) tiles_cap);
        AQP_LAUNCHED();
        const EmitArgs ea{id_base + begin, d_data + begin, d_dict, nullptr};
        size_t g = (size_t) kNumSMs * 6;   // 6 CTAs/SM resident (32 KiB window each)
        const unsigned g1 = (unsigned) (ntiles < g ? ntiles : g);
        g = (size_t) kNumSMs * 8;
        const unsigned g2 = (unsigned) (ntiles < g ? ntiles : g);
#define AQP_EXPAND(KIND)                                                                                             \
    do {                                                                                                             \
        expand_rowids_kernel<KIND><<<g1, kScanThreads, 0, st>>>(bv, nwords, offsets, lists + 2, lists, ea, d_out, cap); \
        AQP_LAUNCHED();                                                                                              \
        expand_dense_rowids_kernel<KIND><<<g2, kScanThreads, 0, st>>>(bv, nwords, offsets, lists + 2 + tiles_cap,      \
                                                                      lists + 1, ea, d_out, cap);                    \
        AQP_LAUNCHED();                                                                                              \
    } while (0)
        if (emit_kind == kEmitRowId) AQP_EXPAND(kEmitRowId);
        else if (emit_kind == kEmitValue) AQP_EXPAND(kEmitValue);
        else AQP_EXPAND(kEmitDict);
#undef AQP_EXPAND
    }
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int index_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t id_base,
                      uint64_t *d_out, uint64_t cap, uint64_t *d_count, void *d_scratch, cudaStream_t st) {
    return emit_scan_device(kEmitRowId, lo, hi, d_data, n, id_base, nullptr, d_out, cap, d_count, d_scratch, st);
}

int value_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint32_t *d_out, uint64_t cap,
                      uint64_t *d_count, void *d_scratch, cudaStream_t st) {
    return emit_scan_device(kEmitValue, lo, hi, d_data, n, 0, nullptr, d_out, cap, d_count, d_scratch, st);
}

int dict_scan_device(uint8_t code_lo, uint8_t code_hi, const int64_t *d_dict, const uint8_t *d_data, size_t n,
                     int64_t *d_out, uint64_t cap, uint64_t *d_count, void *d_scratch, cudaStream_t st) {
    return emit_scan_device(kEmitDict, code_lo, code_hi, d_data, n, 0, d_dict, d_out, cap, d_count, d_scratch, st);
}

int explicit_index_scan_device(uint8_t lo, uint8_t hi, const uint64_t *d_index, const uint8_t *d_data, size_t n,
                               uint64_t *d_out, uint64_t cap, uint64_t *d_count, cudaStream_t st) {
    return emit_scan_device(kEmitExplicit, lo, hi, d_data, n, 0, nullptr, d_out, cap, d_count, nullptr, st, d_index);
}

// the < 64 values behind the last whole block, for the scalar twin of the row-id scan (ScalarScan.hpp:8-20 walks all
// n values, the SIMD kernels only n / 64 blocks): appended behind the ids of the blocks, one thread (at most 63 values)
__global__ void index_scan_tail_kernel(const uint8_t *data, size_t begin, size_t n, uint8_t lo, uint8_t hi, uint64_t id_base,
                                       uint64_t *out, uint64_t cap, unsigned long long *count) {
    unsigned long long c = *count;
    for (size_t i = begin; i < n; ++i) {
        const uint8_t v = data[i];
        if (v >= lo && v <= hi) {
            if (c < cap) out[c] = id_base + i;
            ++c;
        }
    }
    *count = c;
}
int scalar_index_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t id_base, uint64_t *d_out,
                             uint64_t cap, uint64_t *d_count, void *d_scratch, cudaStream_t st) {
    if (index_scan_device(lo, hi, d_data, n, id_base, d_out, cap, d_count, d_scratch, st)) return -1;
    if (n % 64) {
        index_scan_tail_kernel<<<1, 1, 0, st>>>(d_data, n / 64 * 64, n, lo, hi, id_base, d_out, cap,
                                                reinterpret_cast<unsigned long long *>(d_count));
        AQP_LAUNCHED();
        AQP_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// dictionary scans over 16- and 32-bit codes (SIMD512::dict_scan_16bit_64bit / dict_scan_32bit_64bit,
// SIMD512.cpp:531-622): emit dict[code] (int64) for every code with code_lo <= code <= code_hi (unsigned), in
// column order; only whole 512-bit registers are processed (32 resp. 16 codes). Not on the headline path
// (SURVEY 8f rank 4): a count pass, an exclusive scan of the per-tile counts and an emit pass that re-reads the
// column - 2 reads + 8 B per match. A tile is one 128-bit load per thread (8 resp. 4 codes).
// ---------------------------------------------------------------------------------------------
template <typename Code>
__device__ __forceinline__ uint32_t code_mask(uint4 v, uint32_t lo, uint32_t hi) {   // bit k <-> k-th code of the vector
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (sizeof(Code) == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t c = (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
            m |= (c >= lo && c <= hi ? 1u : 0u) << k;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) m |= (w[k] >= lo && w[k] <= hi ? 1u : 0u) << k;
    }
    return m;
}
template <typename Code>
__global__ void __launch_bounds__(256)
code_scan_count_kernel(const uint4 *__restrict__ in, size_t nvec, uint32_t lo, uint32_t hi, uint32_t *__restrict__ tile_counts) {
    __shared__ uint32_t wsum[8];
    for (size_t tile = blockIdx.x; tile * 256 < nvec; tile += gridDim.x) {
        const size_t q = tile * 256 + threadIdx.x;
        uint32_t c = q < nvec ? __popc(code_mask<Code>(ld_stream_v4(in + q), lo, hi)) : 0u;
        c = warp_sum(c);
        if (lane_id() == 0) wsum[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) tile_counts[tile] = wsum[0] + wsum[1] + wsum[2] + wsum[3] + wsum[4] + wsum[5] + wsum[6] + wsum[7];
        __syncthreads();
    }
}
template <typename Code>
__global__ void __launch_bounds__(256)
code_scan_emit_kernel(const uint4 *__restrict__ in, size_t nvec, uint32_t lo, uint32_t hi, const uint32_t *__restrict__ tile_off,
                      const int64_t *__restrict__ dict, int64_t *__restrict__ out, uint64_t cap,
                      unsigned long long *__restrict__ count) {
    __shared__ uint32_t wsum[8];
    constexpr int kPer = 16 / sizeof(Code);
    const size_t ntiles = (nvec + 255) / 256;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t q = tile * 256 + threadIdx.x;
        uint4 v = make_uint4(0, 0, 0, 0);
        uint32_t m = 0;
        if (q < nvec) {
            v = ld_stream_v4(in + q);
            m = code_mask<Code>(v, lo, hi);
        }
        const uint32_t c = __popc(m), incl = warp_incl_scan(c);
        if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            before += k < (int) (threadIdx.x >> 5) ? wsum[k] : 0u;
            total += wsum[k];
        }
        uint64_t slot = (uint64_t) tile_off[tile] + before + incl - c;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            if ((m >> k) & 1u) {
                const uint32_t code = sizeof(Code) == 2 ? (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu : w[k];
                if (slot < cap) out[slot] = __ldg(dict + code);
                ++slot;
            }
        }
        if (tile == ntiles - 1 && threadIdx.x == 0) *count = (unsigned long long) tile_off[tile] + total;
        __syncthreads();
    }
}
static struct {
    DevBuf counts, offsets;
} g_code;
template <typename Code>
static int code_dict_scan_device(uint32_t code_lo, uint32_t code_hi, const int64_t *d_dict, const Code *d_data, size_t n,
                                 int64_t *d_out, uint64_t cap, uint64_t *d_count, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) != 0) {
        set_error("dict scan: column must be 16-byte aligned");
        return -1;
    }
    constexpr size_t kBlock = 64 / sizeof(Code);   // codes per 512-bit register
    n = n / kBlock * kBlock;
    AQP_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    if (n == 0) return 0;
    const size_t nvec = n * sizeof(Code) / 16, ntiles = (nvec + 255) / 256;
    if (ntiles >= 0xFFFFFFFFull) {
        set_error("dict scan: column too large");
        return -1;
    }
    if (g_code.counts.ensure(ntiles * 4 + 64) || g_code.offsets.ensure((ntiles + 1) * 4 + 64)) return -1;
    uint32_t *counts = static_cast<uint32_t *>(g_code.counts.p), *offs = static_cast<uint32_t *>(g_code.offsets.p);
    const unsigned grid = (unsigned) (ntiles < (size_t) kNumSMs * 8 ? ntiles : (size_t) kNumSMs * 8);
    const uint4 *in = reinterpret_cast<const uint4 *>(d_data);
    code_scan_count_kernel<Code><<<grid, 256, 0, st>>>(in, nvec, code_lo, code_hi, counts);
    AQP_LAUNCHED();
    if (exclusive_scan_u32_device(counts, (uint32_t) ntiles, offs, st)) return -1;
    code_scan_emit_kernel<Code><<<grid, 256, 0, st>>>(in, nvec, code_lo, code_hi, offs, d_dict, d_out, cap,
                                                      reinterpret_cast<unsigned long long *>(d_count));
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}
void scan_release_codes() {
    g_code.counts.release();
    g_code.offsets.release();
}
int dict_scan16_device(uint32_t code_lo, uint32_t code_hi, const int64_t *d_dict, const uint16_t *d_data, size_t n,
                       int64_t *d_out, uint64_t cap, uint64_t *d_count, cudaStream_t st) {
    return code_dict_scan_device<uint16_t>(code_lo, code_hi, d_dict, d_data, n, d_out, cap, d_count, st);
}
int dict_scan32_device(uint32_t code_lo, uint32_t code_hi, const int64_t *d_dict, const uint32_t *d_data, size_t n,
                       int64_t *d_out, uint64_t cap, uint64_t *d_count, cudaStream_t st) {
    return code_dict_scan_device<uint32_t>(code_lo, code_hi, d_dict, d_data, n, d_out, cap, d_count, st);
}

int scan_sum_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_sum, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) != 0) {
        set_error("scan_sum: column must be 16-byte aligned");
        return -1;
    }
    AQP_CUDA_OK(cudaMemsetAsync(d_sum, 0, sizeof(uint64_t), st));
    size_t nvec = (n / 64) * 4;
    if (nvec == 0) return 0;
    scan_sum_kernel<<<scan_grid(nvec, 8), kScanThreads, 0, st>>>(reinterpret_cast<const uint4 *>(d_data), nvec,
                                                                reinterpret_cast<unsigned long long *>(d_sum),
                                                                make_pred(lo, hi));
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int fill_tiled_column_device(uint8_t *d, size_t n, uint64_t pos_begin, cudaStream_t st) {
    if (n == 0) return 0;
    fill_tiled_kernel<<<kNumSMs * 8, 256, 0, st>>>(d, n, pos_begin);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int fill_skewed_column_device(uint8_t *d, size_t n, uint64_t pos_begin, uint32_t ppm, uint64_t seed, cudaStream_t st) {
    if (n == 0) return 0;
    fill_skewed_kernel<<<kNumSMs * 8, 256, 0, st>>>(d, n, pos_begin, ppm, seed);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace aqp
