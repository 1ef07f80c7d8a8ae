// scan.cu — range-predicate scans over a packed uint8 column (sm_100a).
//
// Replaces the AVX-512 loops of Scan-Micro-Benchmarks/shared_libraries/SimdScan/src/SIMD512.cpp:
//   count                :7-32     -> scan_count_kernel
//   bitvector_scan       :210-222  -> bitvector_scan_kernel
//   implicit_index_scan  :225-287  -> index_scan_kernel (single pass, decoupled look-back)
// Semantics kept: unsigned inclusive range lo <= v <= hi, only n/64 whole blocks are processed,
// bit k of word i <-> value 64*i+k, row ids ascending uint64 positions.
//
// All three are HBM-bound streaming kernels: 16-byte coalesced loads, 4 independent loads in
// flight per thread, the byte compare done 4 values at a time with carry-free SWAR arithmetic
// (there is no per-byte compare instruction on sm_100; __vcmpgeu4 expands to more ops).
#include "common.cuh"

namespace aqp {

constexpr int kScanThreads = 256;
constexpr int kScanUnroll = 4;                                  // uint4 loads in flight per thread
constexpr int kScanTileVec = kScanThreads * kScanUnroll;        // 1024 uint4 = 16 KiB per tile
constexpr int kScanTileVals = kScanTileVec * 16;                // 16384 values per tile

// bit 7 of every byte of the result is set iff lo <= byte <= hi (unsigned). lo7 = lo4 & 0x7f7f7f7f,
// hiH = hi4 | 0x80808080 are loop invariants. Per byte: (x|0x80) - lo_low never borrows across
// bytes and its bit 7 says x_low >= lo_low; x >= lo  <=>  x7 > lo7 or (x7 == lo7 and that bit).
__device__ __forceinline__ uint32_t inrange_msb(uint32_t x, uint32_t lo4, uint32_t hi4, uint32_t lo7,
                                                uint32_t hiH) {
    const uint32_t H = 0x80808080u;
    uint32_t t = (x | H) - lo7;
    uint32_t ge = (x & ~lo4) | (~(x ^ lo4) & t);
    uint32_t u = hiH - (x & ~H);
    uint32_t le = (hi4 & ~x) | (~(hi4 ^ x) & u);
    return ge & le & H;
}

// 16-bit mask of the 16 bytes of v (bit i <-> byte i in memory order). The multiply gathers the
// four bit-7 flags of a word into its top nibble (bit 7+8k lands on 28+k, no carries collide).
__device__ __forceinline__ uint32_t range_mask16(uint4 v, uint32_t lo4, uint32_t hi4, uint32_t lo7,
                                                 uint32_t hiH) {
    const uint32_t M = 0x00204081u;
    uint32_t p0 = inrange_msb(v.x, lo4, hi4, lo7, hiH) * M;
    uint32_t p1 = inrange_msb(v.y, lo4, hi4, lo7, hiH) * M;
    uint32_t p2 = inrange_msb(v.z, lo4, hi4, lo7, hiH) * M;
    uint32_t p3 = inrange_msb(v.w, lo4, hi4, lo7, hiH) * M;
    return (p0 >> 28) | ((p1 >> 24) & 0xF0u) | ((p2 >> 20) & 0xF00u) | ((p3 >> 16) & 0xF000u);
}

struct Pred {
    uint32_t lo4, hi4, lo7, hiH;
};
static Pred make_pred(uint8_t lo, uint8_t hi) {
    Pred p;
    p.lo4 = 0x01010101u * lo;
    p.hi4 = 0x01010101u * hi;
    p.lo7 = p.lo4 & 0x7f7f7f7fu;
    p.hiH = p.hi4 | 0x80808080u;
    return p;
}

// ---------------------------------------------------------------------------------------------
// bitvector scan
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads)
bitvector_scan_kernel(const uint4 *__restrict__ in, size_t nvec, uint64_t *__restrict__ out, Pred p) {
    const size_t ntiles = (nvec + kScanTileVec - 1) / kScanTileVec;
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
            uint32_t m16 = range_mask16(v[j], p.lo4, p.hi4, p.lo7, p.hiH);
            // four neighbouring lanes hold the four 16-bit quarters of one output word
            uint32_t m32 = m16 | (__shfl_down_sync(0xffffffffu, m16, 1) << 16);
            uint32_t hi = __shfl_down_sync(0xffffffffu, m32, 2);
            if ((threadIdx.x & 3) == 0 && q < nvec) out[q >> 2] = (uint64_t) m32 | ((uint64_t) hi << 32);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// count
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads)
scan_count_kernel(const uint4 *__restrict__ in, size_t nvec, unsigned long long *__restrict__ count, Pred p) {
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
            uint32_t m16 = range_mask16(v[j], p.lo4, p.hi4, p.lo7, p.hiH);
            c += q < nvec ? __popc(m16) : 0;
        }
    }
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
// row-id list scan: one pass over the column; tiles are taken in ticket order and chained with
// a decoupled look-back so every tile learns the number of matches before it without a second
// read of the input. Matches are compacted through shared memory so the uint64 ids leave the SM
// as fully coalesced 256-byte warp stores.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kFlagShift = 62;
constexpr uint64_t kFlagAgg = 1ull << kFlagShift;   // tile aggregate available
constexpr uint64_t kFlagIncl = 2ull << kFlagShift;  // inclusive prefix available
constexpr uint64_t kValMask = (1ull << kFlagShift) - 1;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int kLookWidth = 2;   // predecessors inspected per thread and look-back round (window = 512 tiles)

__global__ void __launch_bounds__(kScanThreads)
index_scan_kernel(const uint4 *__restrict__ in, size_t nvec, uint64_t id_base, uint64_t *__restrict__ out,
                  uint64_t out_capacity, unsigned long long *__restrict__ count_out,
                  uint64_t *__restrict__ tile_state, unsigned int *__restrict__ ticket, Pred p) {
    __shared__ __align__(8) uint16_t smask[kScanTileVec];   // 16-bit masks, one per uint4 of the tile
    __shared__ uint16_t stage[kScanTileVals];               // compacted tile-local positions
    __shared__ uint32_t wtot[kScanThreads / 32];
    __shared__ uint32_t w_p[kScanThreads / 32], w_inv[kScanThreads / 32];
    __shared__ unsigned long long w_sum[kScanThreads / 32];
    __shared__ uint32_t s_tile;

    const uint32_t ntiles = (uint32_t) ((nvec + kScanTileVec - 1) / kScanTileVec);
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    constexpr uint32_t kNone = 0xffffffffu;

    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();   // also protects stage/smask reuse from the previous iteration
        const uint32_t tile = s_tile;
        if (tile >= ntiles) break;

        const size_t base = (size_t) tile * kScanTileVec + threadIdx.x;
        uint4 v[kScanUnroll];
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            v[j] = q < nvec ? ld_stream_v4(in + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < kScanUnroll; ++j) {
            size_t q = base + (size_t) j * kScanThreads;
            uint32_t m16 = q < nvec ? range_mask16(v[j], p.lo4, p.hi4, p.lo7, p.hiH) : 0u;
            smask[j * kScanThreads + threadIdx.x] = (uint16_t) m16;
        }
        __syncthreads();
        // thread t now owns the 64 consecutive values [64t, 64t+64) of the tile
        uint64_t m = reinterpret_cast<const uint64_t *>(smask)[threadIdx.x];
        const uint32_t cnt = __popcll(m);
        const uint32_t incl = warp_incl_scan(cnt);
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) {
            uint32_t t = wtot[w];
            wbase += (w < (int) warp) ? t : 0;
            total += t;
        }
        // publish this tile's aggregate as early as possible
        if (threadIdx.x == 0) st_relaxed_u64(tile_state + tile, (tile == 0 ? kFlagIncl : kFlagAgg) | total);

        // compact this thread's matches into the staging buffer (tile-local positions)
        uint32_t pos = wbase + incl - cnt;
        const uint32_t vbase = threadIdx.x * 64;
        while (m) {
            uint32_t b = __ffsll((long long) m) - 1;
            m &= m - 1;
            stage[pos++] = (uint16_t) (vbase + b);
        }

        // Block-wide decoupled look-back: every thread inspects kLookWidth predecessors per round, so a
        // round covers 512 tiles for the price of one memory round trip. The sum of the aggregates
        // back to (and including) the nearest tile with a known inclusive prefix is this tile's offset.
        unsigned long long excl = 0;
        if (tile > 0) {
            int64_t look = (int64_t) tile - 1;
            while (true) {
                uint64_t sv[kLookWidth];
                uint32_t nearest_p, nearest_inv;
                do {
                    uint32_t my_p = kNone, my_inv = kNone;
#pragma unroll
                    for (int k = kLookWidth - 1; k >= 0; --k) {
                        uint32_t d = k * kScanThreads + threadIdx.x;   // distance behind `look`
                        int64_t idx = look - (int64_t) d;
                        sv[k] = idx >= 0 ? ld_relaxed_u64(tile_state + idx) : kFlagIncl;
                        uint32_t f = (uint32_t) (sv[k] >> kFlagShift);
                        if (f == 2) my_p = d;
                        if (f == 0) my_inv = d;
                    }
                    my_p = __reduce_min_sync(0xffffffffu, my_p);
                    my_inv = __reduce_min_sync(0xffffffffu, my_inv);
                    __syncthreads();   // previous round's w_p / w_inv readers are done
                    if (lane == 0) {
                        w_p[warp] = my_p;
                        w_inv[warp] = my_inv;
                    }
                    __syncthreads();
                    nearest_p = kNone;
                    nearest_inv = kNone;
#pragma unroll
                    for (int w = 0; w < kScanThreads / 32; ++w) {
                        nearest_p = min(nearest_p, w_p[w]);
                        nearest_inv = min(nearest_inv, w_inv[w]);
                    }
                } while (nearest_inv < nearest_p);   // a tile in front of the nearest prefix has not published yet
                unsigned long long part = 0;
#pragma unroll
                for (int k = 0; k < kLookWidth; ++k) {
                    uint32_t d = k * kScanThreads + threadIdx.x;
                    if (d <= nearest_p) part += sv[k] & kValMask;   // nearest_p == kNone: take the whole window
                }
                part = warp_sum(part);
                if (lane == 0) w_sum[warp] = part;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < kScanThreads / 32; ++w) excl += w_sum[w];
                if (nearest_p != kNone) break;
                look -= (int64_t) kLookWidth * kScanThreads;
            }
            if (threadIdx.x == 0) st_relaxed_u64(tile_state + tile, kFlagIncl | (excl + total));
        }
        if (threadIdx.x == 0 && tile == ntiles - 1) *count_out = excl + total;
        __syncthreads();   // staging complete (and w_sum consumed)

        const uint64_t idb = id_base + (uint64_t) tile * kScanTileVals;
        for (uint32_t s = threadIdx.x; s < total; s += kScanThreads) {
            uint64_t g = excl + s;
            if (g < out_capacity) out[g] = idb + stage[s];
        }
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
    bitvector_scan_kernel<<<scan_grid(nvec, 8), kScanThreads, 0, st>>>(reinterpret_cast<const uint4 *>(d_data), nvec,
                                                                      d_out, make_pred(lo, hi));
    AQP_LAUNCHED();
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

// scratch = tile_state[ntiles] (uint64) followed by one uint32 ticket; must hold
// index_scan_scratch_bytes(n) bytes.
size_t index_scan_scratch_bytes(size_t n) {
    size_t nvec = (n / 64) * 4;
    size_t ntiles = (nvec + kScanTileVec - 1) / kScanTileVec;
    return (ntiles + 1) * sizeof(uint64_t);
}

int index_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t id_base,
                      uint64_t *d_out, uint64_t cap, uint64_t *d_count, void *d_scratch, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) != 0) {
        set_error("index_scan: column must be 16-byte aligned");
        return -1;
    }
    size_t nvec = (n / 64) * 4;
    if (nvec == 0) {
        AQP_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
        return 0;
    }
    size_t ntiles = (nvec + kScanTileVec - 1) / kScanTileVec;
    AQP_CUDA_OK(cudaMemsetAsync(d_scratch, 0, index_scan_scratch_bytes(n), st));
    uint64_t *tile_state = reinterpret_cast<uint64_t *>(d_scratch);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(tile_state + ntiles);
    index_scan_kernel<<<scan_grid(nvec, 5), kScanThreads, 0, st>>>(
        reinterpret_cast<const uint4 *>(d_data), nvec, id_base, d_out, cap,
        reinterpret_cast<unsigned long long *>(d_count), tile_state, ticket, make_pred(lo, hi));
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
