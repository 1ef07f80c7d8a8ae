// tpch.cu — TPC-H-style Q3 / Q12 / Q19 pipelines on device-resident columns (sm_100a).
//
// GPU counterpart of Join-Benchmarks/lib/TPCH-Queries/src/tpch.cpp (tpch_q3 :37-117, tpch_q12 :219-253,
// tpch_q19 :255-309): selections (filters.hpp:31 parallel_filter + Q*Predicates.hpp) feeding run_join("RHO"),
// the join-result-to-table transformer copy_Sp_Sp (result_transformers.hpp:51-54,:77-98) and Q19's post-join
// predicate (Q19Predicates.hpp:58-78,:144-192). Everything between the input columns and the final row count
// stays in HBM: filters compact straight into the join's input relations, the materialised join output is
// consumed on the device instead of as a host chunked table.
#include <chrono>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "join_internal.cuh"
#include "aqp/b200_tpch.h"

namespace aqp {

// TpcHTypes.hpp:31-41 (epoch seconds, UTC)
constexpr uint64_t kTs1994_01_01 = 757382400ull;
constexpr uint64_t kTs1995_01_01 = 788918400ull;
constexpr uint64_t kTs1995_03_15 = 795225600ull;
constexpr uint64_t kTs1995_03_16 = 795312000ull;
constexpr uint64_t kTs1992_01_01 = 694224000ull;
constexpr uint32_t kOrderDateDays = 2406;   // 1992-01-01 .. 1998-08-02 (TPC-H 4.2.3: STARTDATE .. ENDDATE - 151 days)

// ---------------------------------------------------------------------------------------------
// device tables
// ---------------------------------------------------------------------------------------------
struct DeviceTables {
    uint64_t nl = 0, no = 0, nc = 0, np = 0;
    uint64_t no_total = 0;           // orders of the whole data set (this rank holds a shard when world > 1)
    uint32_t world = 1, rank = 0;
    DevBuf l_orderkey, l_shipdate, l_commitdate, l_receiptdate, l_shipmode, l_partkey, l_quantity, l_shipinstruct,
        l_returnflag;
    DevBuf o_orderkey, o_orderdate, o_custkey;
    DevBuf c_custkey, c_mktsegment, c_nationkey;
    DevBuf p_partkey, p_brand, p_size, p_container;
    DevBuf f1, f2, u, triples, counters;   // filter outputs, intermediate table, materialised join, counters
    void release_all() {
        for (DevBuf *b : {&l_orderkey, &l_shipdate, &l_commitdate, &l_receiptdate, &l_shipmode, &l_partkey, &l_quantity,
                          &l_shipinstruct, &l_returnflag, &o_orderkey, &o_orderdate, &o_custkey, &c_custkey,
                          &c_mktsegment, &c_nationkey, &p_partkey, &p_brand, &p_size, &p_container, &f1, &f2, &u,
                          &triples, &counters})
            b->release();
        nl = no = nc = np = 0;
    }
};
static DeviceTables T;
static std::recursive_mutex t_mu;

template <typename X>
static X *ptr(DevBuf &b) { return static_cast<X *>(b.p); }

// ---------------------------------------------------------------------------------------------
// synthetic generator (see include/aqp/b200_tpch.h for the distributions)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t h64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint32_t draw(uint64_t seed, uint64_t row, uint32_t column, uint32_t n) {
    return (uint32_t) (h64(seed ^ h64(row * 32 + column)) % n);
}
// dbgen order keys use 8 of every 32 values
__host__ __device__ __forceinline__ uint32_t sparse_orderkey(uint64_t i) { return (uint32_t) ((i >> 3) * 32 + (i & 7) + 1); }
// dbgen order keys use 8 of every 32 values: two of the low five key bits carry no information, so
// the joins on o_orderkey / l_orderkey ask the planner for two more radix bits (same raw-bit digit, balanced sizes)
constexpr uint32_t kOrderKeyDeadBits = 2;
__device__ __forceinline__ uint64_t order_date(uint64_t seed, uint64_t order) {
    return kTs1992_01_01 + 86400ull * draw(seed, order, 1, kOrderDateDays);
}

// Every generator kernel writes rows [r0, r0 + n) of its table to local indices 0..n-1: values depend on the GLOBAL row
// only, so the shards of a multi-GPU run together are exactly the single-GPU tables.
__global__ void gen_customer_kernel(uint2 *custkey, uint8_t *mkt, uint32_t *nation, uint64_t r0, uint64_t n, uint64_t seed) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t i = r0 + j;
        custkey[j] = make_uint2((uint32_t) i + 1, (uint32_t) i);
        mkt[j] = draw(seed, i, 20, 5) == 0 ? B200_MKT_BUILDING : 0;   // 5 segments, only BUILDING is coded
        nation[j] = draw(seed, i, 21, 25);
    }
}

__global__ void gen_orders_kernel(uint2 *orderkey, uint64_t *orderdate, uint32_t *custkey, uint64_t r0, uint64_t n, uint64_t ncust,
                                  uint64_t seed) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t i = r0 + j;
        orderkey[j] = make_uint2(sparse_orderkey(i), (uint32_t) i);
        orderdate[j] = order_date(seed, i);
        // dbgen: customer keys divisible by 3 place no orders
        uint32_t two_thirds = (uint32_t) (ncust - ncust / 3);
        uint32_t k = draw(seed, i, 2, two_thirds ? two_thirds : 1);
        custkey[j] = ncust >= 3 ? k + k / 2 + 1 : 1;   // k-th key not divisible by 3: 1,2,4,5,7,8,...
    }
}

__global__ void gen_lineitem_kernel(uint2 *orderkey, uint64_t *shipdate, uint64_t *commitdate, uint64_t *receiptdate,
                                    uint8_t *shipmode, uint32_t *partkey, float *quantity, uint8_t *shipinstruct,
                                    char *returnflag, uint64_t r0, uint64_t n, uint64_t npart, uint64_t seed) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t i = r0 + j;
        const uint64_t order = i >> 2;   // 4 line items per order
        const uint64_t od = order_date(seed, order);
        orderkey[j] = make_uint2(sparse_orderkey(order), (uint32_t) i);
        const uint64_t sd = od + 86400ull * (1 + draw(seed, i, 3, 121));
        shipdate[j] = sd;
        commitdate[j] = od + 86400ull * (30 + draw(seed, i, 4, 61));
        receiptdate[j] = sd + 86400ull * (1 + draw(seed, i, 5, 30));
        // 7 modes REG AIR, AIR, RAIL, SHIP, TRUCK, MAIL, FOB; the loader codes MAIL, SHIP, AIR and "AIR REG" (never)
        const uint32_t m = draw(seed, i, 6, 7);
        shipmode[j] = m == 5 ? B200_L_SHIPMODE_MAIL : (m == 3 ? B200_L_SHIPMODE_SHIP : (m == 1 ? B200_L_SHIPMODE_AIR : 0));
        partkey[j] = 1 + draw(seed, i, 7, (uint32_t) npart);
        quantity[j] = (float) (1 + draw(seed, i, 8, 50));
        shipinstruct[j] = draw(seed, i, 9, 4) == 0 ? B200_L_SHIPINSTRUCT_DELIVER_IN_PERSON : 0;
        const uint32_t rf = draw(seed, i, 10, 3);
        returnflag[j] = rf == 0 ? 'R' : (rf == 1 ? 'A' : 'N');
    }
}

__global__ void gen_part_kernel(uint2 *partkey, uint8_t *brand, uint32_t *size, uint8_t *container, uint64_t r0, uint64_t n,
                                uint64_t seed) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t i = r0 + j;
        partkey[j] = make_uint2((uint32_t) i + 1, (uint32_t) i);
        // Brand#MN, M,N in 1..5: the loader codes 12 -> 1, 23 -> 2, 34 -> 3 (TpcHTypes.hpp:16-18)
        const uint32_t mn = (1 + draw(seed, i, 11, 5)) * 10 + 1 + draw(seed, i, 12, 5);
        brand[j] = mn == 12 ? 1 : (mn == 23 ? 2 : (mn == 34 ? 3 : 0));
        size[j] = 1 + draw(seed, i, 13, 50);
        // 5 x 8 containers; coded: SM {CASE,BOX,PACK,PKG} = 1..4, MED {BAG,BOX,PKG,PACK} = 5..8, LG {CASE,BOX,PACK,PKG} = 9..12
        const uint32_t s1 = draw(seed, i, 14, 5), s2 = draw(seed, i, 15, 8);   // s1: SM, LG, MED, JUMBO, WRAP; s2: CASE, BOX, BAG, JAR, PKG, PACK, CAN, DRUM
        uint8_t c = 0;
        if (s1 == 0) c = s2 == 0 ? 1 : (s2 == 1 ? 2 : (s2 == 5 ? 3 : (s2 == 4 ? 4 : 0)));
        if (s1 == 2) c = s2 == 2 ? 5 : (s2 == 1 ? 6 : (s2 == 4 ? 7 : (s2 == 5 ? 8 : 0)));
        if (s1 == 1) c = s2 == 0 ? 9 : (s2 == 1 ? 10 : (s2 == 5 ? 11 : (s2 == 4 ? 12 : 0)));
        container[j] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// selection kernels: predicate + compaction into a row_t relation. A CTA evaluates a 1024-row tile with
// coalesced column reads, ranks the survivors with a block scan, reserves its output range with ONE global
// atomicAdd and writes the survivors through shared memory as coalesced stores. The order of the output
// rows is unspecified (the joins that consume them do not depend on it).
// ---------------------------------------------------------------------------------------------
constexpr int kFilterThreads = 256, kFilterRows = 4, kFilterTile = kFilterThreads * kFilterRows;
#ifndef AQP_Q12_BYTE_PREFILTER
#define AQP_Q12_BYTE_PREFILTER 1
#endif

// bit k of the result = byte k of the 16-byte vector m is non-zero, for m made of 0x00 / 0xff bytes (__vcmp*4 results)
__device__ __forceinline__ uint32_t byte_flags16(uint4 m) {
    auto nib = [](uint32_t w) { return ((w & 0x01010101u) * 0x01020408u) >> 24; };
    return nib(m.x) | (nib(m.y) << 4) | (nib(m.z) << 8) | (nib(m.w) << 12);
}
__device__ __forceinline__ uint4 eq_bytes16(uint4 v, uint8_t code) {
    const uint32_t c = 0x01010101u * code;
    return make_uint4(__vcmpeq4(v.x, c), __vcmpeq4(v.y, c), __vcmpeq4(v.z, c), __vcmpeq4(v.w, c));
}
__device__ __forceinline__ uint4 or16(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
__device__ __forceinline__ uint4 and16(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ uint4 ld_bytes16(const uint8_t *col, uint64_t i0) {
    return __ldcs(reinterpret_cast<const uint4 *>(col + i0));
}

struct Q12Lineitem {   // Q12Predicates.hpp:22-37
    static constexpr bool kBytePrefilter = AQP_Q12_BYTE_PREFILTER, kRefine = true;
    __device__ uint32_t prefilter16(uint64_t i0) const {   // l_shipmode in (MAIL, SHIP)
        const uint4 m = ld_bytes16(shipmode, i0);
        return byte_flags16(or16(eq_bytes16(m, B200_L_SHIPMODE_MAIL), eq_bytes16(m, B200_L_SHIPMODE_SHIP)));
    }
    // second level: 2 rows in 7 are candidates, 1 in 7 of those has its receipt date in 1994. The receipt dates of all
    // candidates of the thread's 16 rows (one 128-byte line) are requested together - predicated, independent loads -
    // instead of one dependent load per trip of the candidate loop.
    __device__ uint32_t refine16(uint64_t i0, uint32_t cand) const {
        uint64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = (cand >> k) & 1u ? receipt[i0 + k] : 0ull;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) out |= (uint32_t) (r[k] >= kTs1994_01_01 && r[k] < kTs1995_01_01) << k;
        return out;
    }
    const uint2 *orderkey; const uint8_t *shipmode; const uint64_t *commit, *ship, *receipt;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        const uint8_t m = shipmode[i];
        if (!(m == B200_L_SHIPMODE_MAIL || m == B200_L_SHIPMODE_SHIP)) return false;
        // the receipt-date year first: 1 candidate in 7 survives it, so the sectors of the other two date columns are
        // mostly never touched (the conjunction is the reference's, its order is free)
        const uint64_t r = receipt[i];
        if (!(r >= kTs1994_01_01 && r < kTs1995_01_01)) return false;
        const uint64_t c = commit[i], s = ship[i];
        if (!(c < r && s < c)) return false;
        out = orderkey[i];
        return true;
    }
};
struct Q3Customer {    // Q3Predicates.hpp:25-33
    static constexpr bool kBytePrefilter = false;
    const uint2 *custkey; const uint8_t *mkt;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        if (mkt[i] != B200_MKT_BUILDING) return false;
        out = custkey[i];
        return true;
    }
};
struct Q3Orders {      // Q3Predicates.hpp:35-44: key = o_custkey, payload = o_orderkey
    static constexpr bool kBytePrefilter = false;
    const uint2 *orderkey; const uint64_t *orderdate; const uint32_t *custkey;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        if (!(orderdate[i] < kTs1995_03_15)) return false;
        out = make_uint2(custkey[i], orderkey[i].x);
        return true;
    }
};
struct Q3Lineitem {    // Q3Predicates.hpp:46-54
    static constexpr bool kBytePrefilter = false;
    const uint2 *orderkey; const uint64_t *shipdate;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        if (!(shipdate[i] >= kTs1995_03_16)) return false;
        out = orderkey[i];
        return true;
    }
};
struct Q19Part {       // Q19Predicates.hpp:41-52
    static constexpr bool kBytePrefilter = false;
    const uint2 *partkey; const uint8_t *brand, *container; const uint32_t *size;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        const uint8_t b = brand[i], c = container[i];
        const uint32_t s = size[i];
        if (!((b >= 1 && b <= 3) && (c >= 1 && c <= 12) && (s >= 1 && s <= 15))) return false;
        out = partkey[i];
        return true;
    }
};
struct Q19Lineitem {   // Q19Predicates.hpp:27-39: key = l_partkey, payload = lineitem row id
    static constexpr bool kBytePrefilter = true, kRefine = false;
    __device__ uint32_t prefilter16(uint64_t i0) const {   // l_shipmode in (AIR, AIR REG) and l_shipinstruct = DELIVER IN PERSON
        const uint4 m = ld_bytes16(shipmode, i0), si = ld_bytes16(shipinstruct, i0);
        return byte_flags16(and16(or16(eq_bytes16(m, B200_L_SHIPMODE_AIR), eq_bytes16(m, B200_L_SHIPMODE_AIR_REG)),
                                  eq_bytes16(si, B200_L_SHIPINSTRUCT_DELIVER_IN_PERSON)));
    }
    const uint2 *orderkey; const uint32_t *partkey; const float *quantity; const uint8_t *shipmode, *shipinstruct;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        const uint8_t m = shipmode[i];
        if (!(m == B200_L_SHIPMODE_AIR || m == B200_L_SHIPMODE_AIR_REG)) return false;
        if (shipinstruct[i] != B200_L_SHIPINSTRUCT_DELIVER_IN_PERSON) return false;
        const float q = quantity[i];
        if (!(q >= 1.0f && q <= 30.0f)) return false;
        out = make_uint2(partkey[i], orderkey[i].y);
        return true;
    }
};

template <typename Pred>
__global__ void __launch_bounds__(kFilterThreads)
filter_compact_kernel(uint64_t n, Pred pred, uint2 *__restrict__ out, unsigned long long *__restrict__ counter) {
    __shared__ uint2 stage[kFilterTile];
    __shared__ uint32_t wtot[kFilterThreads / 32];
    __shared__ unsigned long long s_base;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t ntiles = (n + kFilterTile - 1) / kFilterTile;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint2 v[kFilterRows];
        bool keep[kFilterRows];
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < kFilterRows; ++j) {
            const uint64_t i = tile * kFilterTile + (uint64_t) j * kFilterThreads + threadIdx.x;
            keep[j] = i < n && pred(i, v[j]);
            cnt += keep[j];
        }
        const uint32_t incl = warp_incl_scan(cnt);
        __syncthreads();   // previous tile's stage / wtot consumed
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < kFilterThreads / 32; ++k) {
            uint32_t t = wtot[k];
            wbase += (k < (int) warp) ? t : 0;
            total += t;
        }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(counter, (unsigned long long) total);
        uint32_t pos = wbase + incl - cnt;
#pragma unroll
        for (int j = 0; j < kFilterRows; ++j)
            if (keep[j]) stage[pos++] = v[j];
        __syncthreads();
        const unsigned long long base = s_base;
        for (uint32_t s = threadIdx.x; s < total; s += kFilterThreads) out[base + s] = stage[s];
    }
}

// Selections whose first conjuncts test one-byte codes (Q19: l_shipmode, l_shipinstruct - 1 row in 14 survives them):
// the row-at-a-time kernel above spends a dependent chain of 1-byte loads per row on them and runs at a quarter of the
// bandwidth it needs. Here a thread takes 16 CONSECUTIVE rows, reads each code column with one 16-byte load, compares
// the 16 codes with byte-SIMD instructions (Pred::prefilter16 -> candidate bit mask), and evaluates the full predicate
// only on the candidates. The wide columns are then touched for candidate rows only, as before. Two passes over the
// candidates (count, then - after the block-wide scan - evaluate again and store) keep the kernel free of staging
// memory; the second evaluation hits L1/L2 and concerns few rows.
constexpr int kByteRows = 16, kByteTile = kFilterThreads * kByteRows;

template <typename Pred>
__global__ void __launch_bounds__(kFilterThreads)
filter_compact_bytes_kernel(uint64_t n, Pred pred, uint2 *__restrict__ out, unsigned long long *__restrict__ counter) {
    __shared__ uint32_t wtot[kFilterThreads / 32];
    __shared__ unsigned long long s_base;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t ntiles = (n + kByteTile - 1) / kByteTile;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t i0 = tile * kByteTile + (uint64_t) threadIdx.x * kByteRows;
        uint32_t cand = 0;
        if (i0 + kByteRows <= n)
            cand = pred.prefilter16(i0);
        else if (i0 < n)
            cand = (1u << (uint32_t) (n - i0)) - 1;   // ragged end: every row is a candidate
        if constexpr (Pred::kRefine) cand = pred.refine16(i0, cand);
        uint32_t keep = 0;
        uint2 v;
        for (uint32_t m = cand; m; m &= m - 1) {
            const uint32_t k = __ffs(m) - 1;
            if (pred(i0 + k, v)) keep |= 1u << k;
        }
        const uint32_t cnt = __popc(keep);
        const uint32_t incl = warp_incl_scan(cnt);
        __syncthreads();   // previous tile's wtot / s_base consumed
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < kFilterThreads / 32; ++k) {
            uint32_t t = wtot[k];
            wbase += (k < (int) warp) ? t : 0;
            total += t;
        }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(counter, (unsigned long long) total);
        __syncthreads();
        unsigned long long pos = s_base + wbase + incl - cnt;
        for (uint32_t m = keep; m; m &= m - 1) {
            pred(i0 + (__ffs(m) - 1), v);
            out[pos++] = v;
        }
    }
}

template <typename Pred>
static int run_filter(uint64_t n, Pred pred, row_t *d_out, unsigned long long *d_counter, cudaStream_t st) {
    AQP_CUDA_OK(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    if (n == 0) return 0;
    if constexpr (Pred::kBytePrefilter) {
        uint64_t tiles = (n + kByteTile - 1) / kByteTile;
        unsigned grid = (unsigned) (tiles < (uint64_t) kNumSMs * 8 ? tiles : (uint64_t) kNumSMs * 8);
        filter_compact_bytes_kernel<<<grid, kFilterThreads, 0, st>>>(n, pred, reinterpret_cast<uint2 *>(d_out), d_counter);
    } else {
        uint64_t tiles = (n + kFilterTile - 1) / kFilterTile;
        unsigned grid = (unsigned) (tiles < (uint64_t) kNumSMs * 8 ? tiles : (uint64_t) kNumSMs * 8);
        filter_compact_kernel<<<grid, kFilterThreads, 0, st>>>(n, pred, reinterpret_cast<uint2 *>(d_out), d_counter);
    }
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// copy_Sp_Sp (result_transformers.hpp:51-54): join result triple -> {Spayload, Spayload}
__global__ void triples_to_sp_sp_kernel(const output_triple_t *__restrict__ t, uint64_t n, uint2 *__restrict__ out) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t sp = t[i].Spayload;
        out[i] = make_uint2(sp, sp);
    }
}

// q19FinalPredicate (Q19Predicates.hpp:58-78)
__device__ __forceinline__ bool q19_final_pred(uint32_t b, uint32_t k, uint32_t s, float q) {
    const bool p1 = b == 1 && (k >= 1 && k <= 4) && (s >= 1 && s <= 5) && (q >= 1.0f && q <= 11.0f);
    const bool p2 = b == 2 && (k >= 5 && k <= 8) && (s >= 1 && s <= 10) && (q >= 10.0f && q <= 20.0f);
    const bool p3 = b == 3 && (k >= 9 && k <= 12) && (s >= 1 && s <= 15) && (q >= 20.0f && q <= 30.0f);
    return p1 || p2 || p3;
}

// Sharded Q19: after the exchange a match sits on the GPU that owns its part key, not on the ones that hold the two
// rows, so the attributes the final predicate reads travel IN the payloads instead of being gathered by row id:
// part -> brand | container << 8 | size << 16, lineitem -> the bits of l_quantity.
struct Q19PartPacked {
    static constexpr bool kBytePrefilter = false;
    const uint2 *partkey; const uint8_t *brand, *container; const uint32_t *size;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        const uint8_t b = brand[i], c = container[i];
        const uint32_t s = size[i];
        if (!((b >= 1 && b <= 3) && (c >= 1 && c <= 12) && (s >= 1 && s <= 15))) return false;
        out = make_uint2(partkey[i].x, (uint32_t) b | ((uint32_t) c << 8) | (s << 16));
        return true;
    }
};
struct Q19LineitemPacked {
    static constexpr bool kBytePrefilter = true, kRefine = false;
    __device__ uint32_t prefilter16(uint64_t i0) const {   // l_shipmode in (AIR, AIR REG) and l_shipinstruct = DELIVER IN PERSON
        const uint4 m = ld_bytes16(shipmode, i0), si = ld_bytes16(shipinstruct, i0);
        return byte_flags16(and16(or16(eq_bytes16(m, B200_L_SHIPMODE_AIR), eq_bytes16(m, B200_L_SHIPMODE_AIR_REG)),
                                  eq_bytes16(si, B200_L_SHIPINSTRUCT_DELIVER_IN_PERSON)));
    }
    const uint32_t *partkey; const float *quantity; const uint8_t *shipmode, *shipinstruct;
    __device__ bool operator()(uint64_t i, uint2 &out) const {
        const uint8_t m = shipmode[i];
        if (!(m == B200_L_SHIPMODE_AIR || m == B200_L_SHIPMODE_AIR_REG)) return false;
        if (shipinstruct[i] != B200_L_SHIPINSTRUCT_DELIVER_IN_PERSON) return false;
        const float q = quantity[i];
        if (!(q >= 1.0f && q <= 30.0f)) return false;
        out = make_uint2(partkey[i], __float_as_uint(q));
        return true;
    }
};
__global__ void q19_final_packed_kernel(const output_triple_t *__restrict__ t, uint64_t n, unsigned long long *__restrict__ counter) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    uint32_t c = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t rp = t[i].Rpayload;
        c += q19_final_pred(rp & 0xFFu, (rp >> 8) & 0xFFu, rp >> 16, __uint_as_float(t[i].Spayload));
    }
    c = warp_sum(c);
    if (lane_id() == 0 && c) atomicAdd(counter, (unsigned long long) c);
}

// the same over the single-GPU matches, attributes gathered by row id
__global__ void q19_final_kernel(const output_triple_t *__restrict__ t, uint64_t n, const uint8_t *__restrict__ brand,
                                 const uint8_t *__restrict__ container, const uint32_t *__restrict__ size,
                                 const float *__restrict__ quantity, unsigned long long *__restrict__ counter) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    uint32_t c = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t rp = t[i].Rpayload, rl = t[i].Spayload;   // row ids in part / lineitem
        const uint8_t b = brand[rp], k = container[rp];
        const uint32_t s = size[rp];
        const float q = quantity[rl];
        c += q19_final_pred(b, k, s, q);
    }
    c = warp_sum(c);
    if (lane_id() == 0 && c) atomicAdd(counter, (unsigned long long) c);
}

// ---------------------------------------------------------------------------------------------
// pipelines
// ---------------------------------------------------------------------------------------------
struct Timer {
    cudaEvent_t e[6];
    Timer() { for (auto &x : e) cudaEventCreate(&x); }
    ~Timer() { for (auto &x : e) cudaEventDestroy(x); }
    float ms(int a, int b) { float m = 0; cudaEventElapsedTime(&m, e[a], e[b]); return m; }
};

static int read_counter(unsigned long long *d, uint64_t *h, cudaStream_t st) {
    unsigned long long v = 0;
    AQP_CUDA_OK(cudaMemcpyAsync(&v, d, sizeof v, cudaMemcpyDeviceToHost, st));
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    *h = v;
    return 0;
}

static int need(bool ok, const char *what) {
    if (!ok) {
        set_error(std::string("tpch: table not loaded: ") + what + " (call b200_tpch_generate_device or b200_tpch_upload)");
        return -1;
    }
    return 0;
}

// materialising join whose result size is only bounded by a guess (exact for a unique build key): if the build side
// has duplicate keys the join reports more matches than fit, and it is run again with room for all of them - the
// reference simply materialises whatever run_join produces (tpch.cpp:64-68,:281-282)
static int materialising_join(const row_t *dR, uint64_t nR, const row_t *dS, uint64_t nS, uint64_t guess, b200_join_stats_t *js,
                              float *ms_join, cudaStream_t st) {
    if (T.triples.ensure((guess + 1) * sizeof(output_triple_t))) return -1;
    if (join_device_internal(dR, nR, dS, nS, ptr<output_triple_t>(T.triples), guess + 1, js, st)) return -1;
    *ms_join += js->ms_total;
    if ((uint64_t) js->matches > guess + 1) {
        const uint64_t cap = (uint64_t) js->matches;
        if (T.triples.ensure(cap * sizeof(output_triple_t))) return -1;
        if (join_device_internal(dR, nR, dS, nS, ptr<output_triple_t>(T.triples), cap, js, st)) return -1;
        *ms_join += js->ms_total;
    }
    return 0;
}

static int q12_device(b200_tpch_stats_t *out) {
    if (need(T.nl && T.no, "lineitem, orders")) return -1;
    cudaStream_t st = library_stream();
    const unsigned long long l0 = g_kernel_launches;
    Timer tm;
    if (T.f1.ensure(T.nl * 8 + 64) || T.counters.ensure(64)) return -1;
    unsigned long long *ctr = ptr<unsigned long long>(T.counters);
    b200_tpch_stats_t s{};
    cudaEventRecord(tm.e[0], st);
    Q12Lineitem pred{ptr<uint2>(T.l_orderkey), ptr<uint8_t>(T.l_shipmode), ptr<uint64_t>(T.l_commitdate),
                     ptr<uint64_t>(T.l_shipdate), ptr<uint64_t>(T.l_receiptdate)};
    if (run_filter(T.nl, pred, ptr<row_t>(T.f1), ctr, st)) return -1;   // selection 1 (tpch.cpp:230-233)
    cudaEventRecord(tm.e[1], st);
    if (read_counter(ctr, &s.filtered[0], st)) return -1;
    b200_join_stats_t js{};
    // join orders (all rows, key = o_orderkey) with the selected line items, count only (tpch.cpp:240-241)
    if (join_device_internal(ptr<row_t>(T.o_orderkey), T.no, ptr<row_t>(T.f1), s.filtered[0], nullptr, 0, &js, st,
                             kOrderKeyDeadBits))
        return -1;
    cudaEventRecord(tm.e[2], st);
    cudaEventSynchronize(tm.e[2]);
    s.result_rows = (uint64_t) js.matches;
    s.input_rows = T.nl + T.no;
    s.ms_filter = tm.ms(0, 1);
    s.ms_join = js.ms_total;
    s.ms_total = tm.ms(0, 2);
    s.ms_other = s.ms_total - s.ms_filter - s.ms_join;
    s.kernel_launches = (uint32_t) (g_kernel_launches - l0);
    *out = s;
    return 0;
}

static int q3_device(b200_tpch_stats_t *out) {
    if (need(T.nl && T.no && T.nc, "lineitem, orders, customer")) return -1;
    cudaStream_t st = library_stream();
    const unsigned long long l0 = g_kernel_launches;
    Timer tm;
    if (T.f1.ensure((T.nc > T.nl ? T.nc : T.nl) * 8 + 64) || T.f2.ensure(T.no * 8 + 64) || T.counters.ensure(64)) return -1;
    unsigned long long *ctr = ptr<unsigned long long>(T.counters);
    b200_tpch_stats_t s{};
    float ms_join = 0;
    cudaEventRecord(tm.e[0], st);
    // selections 1 and 2 (tpch.cpp:52-55)
    if (run_filter(T.nc, Q3Customer{ptr<uint2>(T.c_custkey), ptr<uint8_t>(T.c_mktsegment)}, ptr<row_t>(T.f1), ctr, st)) return -1;
    if (run_filter(T.no, Q3Orders{ptr<uint2>(T.o_orderkey), ptr<uint64_t>(T.o_orderdate), ptr<uint32_t>(T.o_custkey)},
                   ptr<row_t>(T.f2), ctr + 1, st))
        return -1;
    cudaEventRecord(tm.e[1], st);
    if (read_counter(ctr, &s.filtered[0], st) || read_counter(ctr + 1, &s.filtered[1], st)) return -1;
    // join 1: customers x orders, materialised (tpch.cpp:64-68); every order has one customer -> <= |orders| matches
    b200_join_stats_t js{};
    if (materialising_join(ptr<row_t>(T.f1), s.filtered[0], ptr<row_t>(T.f2), s.filtered[1], s.filtered[1], &js, &ms_join, st))
        return -1;
    s.join1_rows = (uint64_t) js.matches;
    // transform to the build side of join 2: {o_orderkey, o_orderkey} (tpch.cpp:76-83)
    cudaEventRecord(tm.e[2], st);
    if (T.u.ensure(s.join1_rows * 8 + 64)) return -1;
    if (s.join1_rows) {
        triples_to_sp_sp_kernel<<<kNumSMs * 4, 256, 0, st>>>(ptr<output_triple_t>(T.triples), s.join1_rows, ptr<uint2>(T.u));
        AQP_LAUNCHED();
    }
    // selection 3 (tpch.cpp:92-93)
    if (run_filter(T.nl, Q3Lineitem{ptr<uint2>(T.l_orderkey), ptr<uint64_t>(T.l_shipdate)}, ptr<row_t>(T.f1), ctr + 2, st)) return -1;
    cudaEventRecord(tm.e[3], st);
    if (read_counter(ctr + 2, &s.filtered[2], st)) return -1;
    // join 2: U x lineitem, count only (tpch.cpp:100-101)
    if (join_device_internal(ptr<row_t>(T.u), s.join1_rows, ptr<row_t>(T.f1), s.filtered[2], nullptr, 0, &js, st,
                             kOrderKeyDeadBits))
        return -1;
    ms_join += js.ms_total;
    cudaEventRecord(tm.e[4], st);
    cudaEventSynchronize(tm.e[4]);
    s.result_rows = (uint64_t) js.matches;
    s.input_rows = T.nl + T.no + T.nc;
    s.ms_filter = tm.ms(0, 1) + tm.ms(2, 3);
    s.ms_join = ms_join;
    s.ms_total = tm.ms(0, 4);
    s.ms_other = s.ms_total - s.ms_filter - s.ms_join;
    s.kernel_launches = (uint32_t) (g_kernel_launches - l0);
    *out = s;
    return 0;
}

static int q19_device(b200_tpch_stats_t *out) {
    if (need(T.nl && T.np, "lineitem, part")) return -1;
    cudaStream_t st = library_stream();
    const unsigned long long l0 = g_kernel_launches;
    Timer tm;
    if (T.f1.ensure(T.np * 8 + 64) || T.f2.ensure(T.nl * 8 + 64) || T.counters.ensure(64)) return -1;
    unsigned long long *ctr = ptr<unsigned long long>(T.counters);
    b200_tpch_stats_t s{};
    cudaEventRecord(tm.e[0], st);
    // selections 1 and 2 (tpch.cpp:268-274)
    if (run_filter(T.np, Q19Part{ptr<uint2>(T.p_partkey), ptr<uint8_t>(T.p_brand), ptr<uint8_t>(T.p_container), ptr<uint32_t>(T.p_size)},
                   ptr<row_t>(T.f1), ctr, st))
        return -1;
    if (run_filter(T.nl, Q19Lineitem{ptr<uint2>(T.l_orderkey), ptr<uint32_t>(T.l_partkey), ptr<float>(T.l_quantity),
                                     ptr<uint8_t>(T.l_shipmode), ptr<uint8_t>(T.l_shipinstruct)},
                   ptr<row_t>(T.f2), ctr + 1, st))
        return -1;
    cudaEventRecord(tm.e[1], st);
    if (read_counter(ctr, &s.filtered[0], st) || read_counter(ctr + 1, &s.filtered[1], st)) return -1;
    // join 1: part x lineitem, materialised (tpch.cpp:281-282); part keys are unique -> <= |lineitem'| matches
    b200_join_stats_t js{};
    float ms_join = 0;
    if (materialising_join(ptr<row_t>(T.f1), s.filtered[0], ptr<row_t>(T.f2), s.filtered[1], s.filtered[1], &js, &ms_join, st))
        return -1;
    s.join1_rows = (uint64_t) js.matches;
    // selection 3: re-check the combined predicate on the matches by row id (tpch.cpp:288-299)
    cudaEventRecord(tm.e[2], st);
    AQP_CUDA_OK(cudaMemsetAsync(ctr + 2, 0, sizeof(unsigned long long), st));
    if (s.join1_rows) {
        q19_final_kernel<<<kNumSMs * 4, 256, 0, st>>>(ptr<output_triple_t>(T.triples), s.join1_rows, ptr<uint8_t>(T.p_brand),
                                                      ptr<uint8_t>(T.p_container), ptr<uint32_t>(T.p_size),
                                                      ptr<float>(T.l_quantity), ctr + 2);
        AQP_LAUNCHED();
    }
    cudaEventRecord(tm.e[3], st);
    if (read_counter(ctr + 2, &s.filtered[2], st)) return -1;
    s.result_rows = s.filtered[2];
    s.input_rows = T.nl + T.np;
    s.ms_filter = tm.ms(0, 1);
    s.ms_join = ms_join;
    s.ms_total = tm.ms(0, 3);
    s.ms_other = s.ms_total - s.ms_filter - s.ms_join;
    s.kernel_launches = (uint32_t) (g_kernel_launches - l0);
    *out = s;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// table management
// ---------------------------------------------------------------------------------------------
static int upload(DevBuf &b, const void *h, size_t bytes, cudaStream_t st) {
    if (b.ensure(bytes + 64)) return -1;
    AQP_CUDA_OK(cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}
template <typename X>
static int download(X *&h, DevBuf &b, size_t count, cudaStream_t st) {
    h = static_cast<X *>(malloc((count ? count : 1) * sizeof(X)));
    if (!h) {
        set_error("out of host memory");
        return -1;
    }
    AQP_CUDA_OK(cudaMemcpyAsync(h, b.p, count * sizeof(X), cudaMemcpyDeviceToHost, st));
    return 0;
}

static int upload_tables(const LineItemTable *l, const OrdersTable *o, const CustomerTable *c, const PartTable *p) {
    cudaStream_t st = library_stream();
    if (!st) return -1;
    if (l) {
        const size_t n = l->numTuples;
        T.nl = n;
        if (l->l_orderkey && upload(T.l_orderkey, l->l_orderkey, n * 8, st)) return -1;
        if (l->l_shipdate && upload(T.l_shipdate, l->l_shipdate, n * 8, st)) return -1;
        if (l->l_commitdate && upload(T.l_commitdate, l->l_commitdate, n * 8, st)) return -1;
        if (l->l_receiptdate && upload(T.l_receiptdate, l->l_receiptdate, n * 8, st)) return -1;
        if (l->l_shipmode && upload(T.l_shipmode, l->l_shipmode, n, st)) return -1;
        if (l->l_partkey && upload(T.l_partkey, l->l_partkey, n * 4, st)) return -1;
        if (l->l_quantity && upload(T.l_quantity, l->l_quantity, n * 4, st)) return -1;
        if (l->l_shipinstruct && upload(T.l_shipinstruct, l->l_shipinstruct, n, st)) return -1;
        if (l->l_returnflag && upload(T.l_returnflag, l->l_returnflag, n, st)) return -1;
    }
    if (o) {
        const size_t n = o->numTuples;
        T.no = n;
        if (o->o_orderkey && upload(T.o_orderkey, o->o_orderkey, n * 8, st)) return -1;
        if (o->o_orderdate && upload(T.o_orderdate, o->o_orderdate, n * 8, st)) return -1;
        if (o->o_custkey && upload(T.o_custkey, o->o_custkey, n * 4, st)) return -1;
    }
    if (c) {
        const size_t n = c->numTuples;
        T.nc = n;
        if (c->c_custkey && upload(T.c_custkey, c->c_custkey, n * 8, st)) return -1;
        if (c->c_mktsegment && upload(T.c_mktsegment, c->c_mktsegment, n, st)) return -1;
        if (c->c_nationkey && upload(T.c_nationkey, c->c_nationkey, n * 4, st)) return -1;
    }
    if (p) {
        const size_t n = p->numTuples;
        T.np = n;
        if (p->p_partkey && upload(T.p_partkey, p->p_partkey, n * 8, st)) return -1;
        if (p->p_brand && upload(T.p_brand, p->p_brand, n, st)) return -1;
        if (p->p_size && upload(T.p_size, p->p_size, n * 4, st)) return -1;
        if (p->p_container && upload(T.p_container, p->p_container, n, st)) return -1;
    }
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

[[noreturn]] static void die_tpch(const char *what) {
    fprintf(stderr, "[b200aqp][ERROR] %s: %s\n", what, g_last_error.c_str());
    exit(EXIT_FAILURE);
}

static void fill_result(result_t *res, const b200_tpch_stats_t &s, const joinconfig_t *cfg) {
    res->totalresults = (int64_t) s.result_rows;
    res->nthreads = cfg ? cfg->NTHREADS : 0;
    res->materialized = 0;
    res->result_type = 1;
    chunked_table_t *ct = static_cast<chunked_table_t *>(calloc(1, sizeof(chunked_table_t)));
    ct->chunks = static_cast<table_chunk_t **>(malloc(sizeof(table_chunk_t *)));
    ct->current_chunk = (uint64_t) -1;
    res->result = ct;
    // tpch.cpp:116: throughput = input rows / total cycles; here rows per microsecond of device time
    res->throughput = s.ms_total > 0 ? (double) s.input_rows / (s.ms_total * 1e3) : 0.0;
}

static void check_algorithm(const char *algorithm) {
    if (!algorithm || strcmp(algorithm, "RHO") != 0) {
        fprintf(stderr, "[b200aqp][ERROR] Algorithm not found: %s (this library serves RHO only)\n",
                algorithm ? algorithm : "(null)");
        exit(EXIT_FAILURE);
    }
}

}  // namespace aqp

using namespace aqp;

extern "C" {

// rows [total * rank / world, total * (rank + 1) / world) of every table (orders in whole orders: 4 line items each)
static int generate_shard(double sf, uint64_t seed, uint32_t rank, uint32_t world) {
    cudaStream_t st = library_stream();
    if (!st) return -1;
    if (!(sf > 0) || sf > 300 || world == 0 || rank >= world) {
        set_error("b200_tpch_generate_device: scale factor must be in (0, 300], rank < world");
        return -1;
    }
    const uint64_t nc_t = (uint64_t) (150000.0 * sf), no_t = (uint64_t) (1500000.0 * sf), np_t = (uint64_t) (200000.0 * sf);
    if (no_t * 4 >= 0xFFFF0000ull || nc_t < 3 || np_t < 1) {
        set_error("b200_tpch_generate_device: scale factor out of range");
        return -1;
    }
    auto lo = [&](uint64_t t) { return t * rank / world; };
    auto hi = [&](uint64_t t) { return t * (rank + 1) / world; };
    const uint64_t c0 = lo(nc_t), nc = hi(nc_t) - c0, o0 = lo(no_t), no = hi(no_t) - o0, l0 = o0 * 4, nl = no * 4,
                   p0 = lo(np_t), np = hi(np_t) - p0;
    if (T.l_orderkey.ensure(nl * 8 + 64) || T.l_shipdate.ensure(nl * 8 + 64) || T.l_commitdate.ensure(nl * 8 + 64) ||
        T.l_receiptdate.ensure(nl * 8 + 64) || T.l_shipmode.ensure(nl + 64) || T.l_partkey.ensure(nl * 4 + 64) ||
        T.l_quantity.ensure(nl * 4 + 64) || T.l_shipinstruct.ensure(nl + 64) || T.l_returnflag.ensure(nl + 64) ||
        T.o_orderkey.ensure(no * 8 + 64) || T.o_orderdate.ensure(no * 8 + 64) || T.o_custkey.ensure(no * 4 + 64) ||
        T.c_custkey.ensure(nc * 8 + 64) || T.c_mktsegment.ensure(nc + 64) || T.c_nationkey.ensure(nc * 4 + 64) ||
        T.p_partkey.ensure(np * 8 + 64) || T.p_brand.ensure(np + 64) || T.p_size.ensure(np * 4 + 64) ||
        T.p_container.ensure(np + 64))
        return -1;
    gen_customer_kernel<<<kNumSMs * 4, 256, 0, st>>>(ptr<uint2>(T.c_custkey), ptr<uint8_t>(T.c_mktsegment),
                                                     ptr<uint32_t>(T.c_nationkey), c0, nc, seed);
    AQP_LAUNCHED();
    gen_orders_kernel<<<kNumSMs * 8, 256, 0, st>>>(ptr<uint2>(T.o_orderkey), ptr<uint64_t>(T.o_orderdate),
                                                   ptr<uint32_t>(T.o_custkey), o0, no, nc_t, seed);
    AQP_LAUNCHED();
    gen_lineitem_kernel<<<kNumSMs * 8, 256, 0, st>>>(ptr<uint2>(T.l_orderkey), ptr<uint64_t>(T.l_shipdate),
                                                     ptr<uint64_t>(T.l_commitdate), ptr<uint64_t>(T.l_receiptdate),
                                                     ptr<uint8_t>(T.l_shipmode), ptr<uint32_t>(T.l_partkey),
                                                     ptr<float>(T.l_quantity), ptr<uint8_t>(T.l_shipinstruct),
                                                     ptr<char>(T.l_returnflag), l0, nl, np_t, seed);
    AQP_LAUNCHED();
    gen_part_kernel<<<kNumSMs * 4, 256, 0, st>>>(ptr<uint2>(T.p_partkey), ptr<uint8_t>(T.p_brand), ptr<uint32_t>(T.p_size),
                                                 ptr<uint8_t>(T.p_container), p0, np, seed);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    T.nl = nl;
    T.no = no;
    T.nc = nc;
    T.np = np;
    T.no_total = no_t;
    T.world = world;
    T.rank = rank;
    return 0;
}

int b200_tpch_generate_device(double sf, uint64_t seed) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    return generate_shard(sf, seed, 0, 1);
}

int b200_tpch_generate_shard_device(double sf, uint64_t seed, uint32_t rank, uint32_t world) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    return generate_shard(sf, seed, rank, world);
}

// Q12 across `world` GPUs: every rank filters ITS line items (selection 1 is a row-range scan, no exchange) and the
// join orders x selected line items is the sharded join of csrc/mg.cu. The caller initialises the multi-GPU host with
// b200_tpch_mg_init (capacities follow from the shard sizes; the order keys' two dead bits go into the radix plan).
int b200_tpch_mg_init(int rank, int world, const unsigned char *id) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    if (need(T.nl && T.no && T.world == (uint32_t) world && T.rank == (uint32_t) rank, "lineitem, orders of this rank's shard")) return -1;
    const uint64_t cap_o = (T.no_total + world - 1) / world + 1;
    return b200_mg_init_caps(rank, world, id, T.no_total, cap_o, cap_o * 4, kOrderKeyDeadBits);
}

int b200_tpch_q12_mg(struct b200_tpch_stats_t *out) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    if (need(T.nl && T.no, "lineitem, orders")) return -1;
    cudaStream_t st = library_stream();
    const unsigned long long l0 = g_kernel_launches;
    Timer tm;
    if (T.f1.ensure(T.nl * 8 + 64) || T.counters.ensure(64)) return -1;
    unsigned long long *ctr = ptr<unsigned long long>(T.counters);
    b200_tpch_stats_t s{};
    cudaEventRecord(tm.e[0], st);
    Q12Lineitem pred{ptr<uint2>(T.l_orderkey), ptr<uint8_t>(T.l_shipmode), ptr<uint64_t>(T.l_commitdate),
                     ptr<uint64_t>(T.l_shipdate), ptr<uint64_t>(T.l_receiptdate)};
    if (run_filter(T.nl, pred, ptr<row_t>(T.f1), ctr, st)) return -1;
    cudaEventRecord(tm.e[1], st);
    if (read_counter(ctr, &s.filtered[0], st)) return -1;   // this rank's selected line items
    b200_mg_result_t r{};
    if (b200_mg_join(ptr<row_t>(T.o_orderkey), T.no, ptr<row_t>(T.f1), s.filtered[0], &r)) return -1;
    s.result_rows = r.matches;                              // global
    s.input_rows = T.nl + T.no;                             // this rank's rows
    s.ms_filter = tm.ms(0, 1);
    s.ms_join = r.ms_total;
    s.ms_total = s.ms_filter + s.ms_join;
    s.kernel_launches = (uint32_t) (g_kernel_launches - l0);
    *out = s;
    return 0;
}

// Q3 across `world` GPUs (tpch.cpp:40-110 on row-range shards): both selections of join 1 are local scans, the join
// customers x orders is the materialising sharded join - its matches stay on the GPU that owns the customer key -,
// each rank turns ITS matches into {o_orderkey, o_orderkey} and they are the (again row-sharded) build side of the
// second sharded join with the selected line items. No table row ever moves except through the two exchanges.
int b200_tpch_q3_mg(struct b200_tpch_stats_t *out) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    if (need(T.nl && T.no && T.nc, "lineitem, orders, customer")) return -1;
    cudaStream_t st = library_stream();
    const unsigned long long l0 = g_kernel_launches;
    Timer tm;
    if (T.f1.ensure((T.nc > T.nl ? T.nc : T.nl) * 8 + 64) || T.f2.ensure(T.no * 8 + 64) || T.counters.ensure(64)) return -1;
    unsigned long long *ctr = ptr<unsigned long long>(T.counters);
    b200_tpch_stats_t s{};
    cudaEventRecord(tm.e[0], st);
    if (run_filter(T.nc, Q3Customer{ptr<uint2>(T.c_custkey), ptr<uint8_t>(T.c_mktsegment)}, ptr<row_t>(T.f1), ctr, st)) return -1;
    if (run_filter(T.no, Q3Orders{ptr<uint2>(T.o_orderkey), ptr<uint64_t>(T.o_orderdate), ptr<uint32_t>(T.o_custkey)},
                   ptr<row_t>(T.f2), ctr + 1, st))
        return -1;
    cudaEventRecord(tm.e[1], st);
    if (read_counter(ctr, &s.filtered[0], st) || read_counter(ctr + 1, &s.filtered[1], st)) return -1;
    b200_mg_result_t r1{}, r2{};
    const output_triple_t *trip = nullptr;
    uint64_t rows = 0;
    if (b200_mg_join_materialize(ptr<row_t>(T.f1), s.filtered[0], ptr<row_t>(T.f2), s.filtered[1], &trip, &rows, &r1)) return -1;
    s.join1_rows = r1.matches;   // global
    cudaEventRecord(tm.e[2], st);
    if (T.u.ensure(rows * 8 + 64)) return -1;
    if (rows) {
        triples_to_sp_sp_kernel<<<kNumSMs * 4, 256, 0, st>>>(trip, rows, ptr<uint2>(T.u));
        AQP_LAUNCHED();
    }
    if (run_filter(T.nl, Q3Lineitem{ptr<uint2>(T.l_orderkey), ptr<uint64_t>(T.l_shipdate)}, ptr<row_t>(T.f1), ctr + 2, st)) return -1;
    cudaEventRecord(tm.e[3], st);
    if (read_counter(ctr + 2, &s.filtered[2], st)) return -1;
    if (b200_mg_join(ptr<row_t>(T.u), rows, ptr<row_t>(T.f1), s.filtered[2], &r2)) return -1;
    s.result_rows = r2.matches;   // global
    s.input_rows = T.nl + T.no + T.nc;
    s.ms_filter = tm.ms(0, 1) + tm.ms(2, 3);
    s.ms_join = r1.ms_total + r2.ms_total;
    s.ms_total = s.ms_filter + s.ms_join;
    s.kernel_launches = (uint32_t) (g_kernel_launches - l0);
    *out = s;
    return 0;
}

// Q19 across `world` GPUs (tpch.cpp:256-306): local selections with the final predicate's attributes packed into the
// payloads, materialising sharded join part x lineitem, final predicate over this rank's matches, one all-reduce.
int b200_tpch_q19_mg(struct b200_tpch_stats_t *out) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    if (need(T.nl && T.np, "lineitem, part")) return -1;
    cudaStream_t st = library_stream();
    const unsigned long long l0 = g_kernel_launches;
    Timer tm;
    if (T.f1.ensure(T.np * 8 + 64) || T.f2.ensure(T.nl * 8 + 64) || T.counters.ensure(64)) return -1;
    unsigned long long *ctr = ptr<unsigned long long>(T.counters);
    b200_tpch_stats_t s{};
    cudaEventRecord(tm.e[0], st);
    if (run_filter(T.np, Q19PartPacked{ptr<uint2>(T.p_partkey), ptr<uint8_t>(T.p_brand), ptr<uint8_t>(T.p_container), ptr<uint32_t>(T.p_size)},
                   ptr<row_t>(T.f1), ctr, st))
        return -1;
    if (run_filter(T.nl, Q19LineitemPacked{ptr<uint32_t>(T.l_partkey), ptr<float>(T.l_quantity), ptr<uint8_t>(T.l_shipmode),
                                           ptr<uint8_t>(T.l_shipinstruct)},
                   ptr<row_t>(T.f2), ctr + 1, st))
        return -1;
    cudaEventRecord(tm.e[1], st);
    if (read_counter(ctr, &s.filtered[0], st) || read_counter(ctr + 1, &s.filtered[1], st)) return -1;
    b200_mg_result_t r{};
    const output_triple_t *trip = nullptr;
    uint64_t rows = 0;
    if (b200_mg_join_materialize(ptr<row_t>(T.f1), s.filtered[0], ptr<row_t>(T.f2), s.filtered[1], &trip, &rows, &r)) return -1;
    s.join1_rows = r.matches;   // global
    cudaEventRecord(tm.e[2], st);
    AQP_CUDA_OK(cudaMemsetAsync(ctr + 2, 0, sizeof(unsigned long long), st));
    if (rows) {
        q19_final_packed_kernel<<<kNumSMs * 4, 256, 0, st>>>(trip, rows, ctr + 2);
        AQP_LAUNCHED();
    }
    cudaEventRecord(tm.e[3], st);
    if (read_counter(ctr + 2, &s.filtered[2], st)) return -1;   // this rank's qualifying matches
    uint64_t total = s.filtered[2];
    if (b200_mg_allreduce_u64(&total, 1)) return -1;
    s.result_rows = total;       // global
    s.input_rows = T.nl + T.np;
    s.ms_filter = tm.ms(0, 1);
    s.ms_join = r.ms_total;
    s.ms_other = tm.ms(2, 3);
    s.ms_total = s.ms_filter + s.ms_join + s.ms_other;
    s.kernel_launches = (uint32_t) (g_kernel_launches - l0);
    *out = s;
    return 0;
}

int b200_tpch_upload(const struct LineItemTable *l, const struct OrdersTable *o, const struct CustomerTable *c,
                     const struct PartTable *p) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    return upload_tables(l, o, c, p);
}

int b200_tpch_download(struct LineItemTable *l, struct OrdersTable *o, struct CustomerTable *c, struct PartTable *p) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    cudaStream_t st = library_stream();
    if (!st) return -1;
    if (l) {
        memset(l, 0, sizeof *l);
        l->numTuples = T.nl;
        if (download(l->l_orderkey, T.l_orderkey, T.nl, st) || download(l->l_shipdate, T.l_shipdate, T.nl, st) ||
            download(l->l_commitdate, T.l_commitdate, T.nl, st) || download(l->l_receiptdate, T.l_receiptdate, T.nl, st) ||
            download(l->l_shipmode, T.l_shipmode, T.nl, st) || download(l->l_partkey, T.l_partkey, T.nl, st) ||
            download(l->l_quantity, T.l_quantity, T.nl, st) || download(l->l_shipinstruct, T.l_shipinstruct, T.nl, st) ||
            download(l->l_returnflag, T.l_returnflag, T.nl, st))
            return -1;
    }
    if (o) {
        memset(o, 0, sizeof *o);
        o->numTuples = T.no;
        if (download(o->o_orderkey, T.o_orderkey, T.no, st) || download(o->o_orderdate, T.o_orderdate, T.no, st) ||
            download(o->o_custkey, T.o_custkey, T.no, st))
            return -1;
    }
    if (c) {
        memset(c, 0, sizeof *c);
        c->numTuples = T.nc;
        if (download(c->c_custkey, T.c_custkey, T.nc, st) || download(c->c_mktsegment, T.c_mktsegment, T.nc, st) ||
            download(c->c_nationkey, T.c_nationkey, T.nc, st))
            return -1;
    }
    if (p) {
        memset(p, 0, sizeof *p);
        p->numTuples = T.np;
        if (download(p->p_partkey, T.p_partkey, T.np, st) || download(p->p_brand, T.p_brand, T.np, st) ||
            download(p->p_size, T.p_size, T.np, st) || download(p->p_container, T.p_container, T.np, st))
            return -1;
    }
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

void b200_tpch_free_host(struct LineItemTable *l, struct OrdersTable *o, struct CustomerTable *c, struct PartTable *p) {
    if (l) {
        free(l->l_orderkey); free(l->l_shipdate); free(l->l_commitdate); free(l->l_receiptdate); free(l->l_shipmode);
        free(l->l_partkey); free(l->l_quantity); free(l->l_shipinstruct); free(l->l_returnflag);
        memset(l, 0, sizeof *l);
    }
    if (o) {
        free(o->o_orderkey); free(o->o_orderdate); free(o->o_custkey);
        memset(o, 0, sizeof *o);
    }
    if (c) {
        free(c->c_custkey); free(c->c_mktsegment); free(c->c_nationkey);
        memset(c, 0, sizeof *c);
    }
    if (p) {
        free(p->p_partkey); free(p->p_brand); free(p->p_size); free(p->p_container);
        memset(p, 0, sizeof *p);
    }
}

void b200_tpch_free_device(void) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    T.release_all();
}

int b200_tpch_q3_device(struct b200_tpch_stats_t *stats) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    return q3_device(stats);
}
int b200_tpch_q12_device(struct b200_tpch_stats_t *stats) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    return q12_device(stats);
}
int b200_tpch_q19_device(struct b200_tpch_stats_t *stats) {
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    return q19_device(stats);
}

void tpch_q3(struct result_t *result, const struct CustomerTable *c, const struct OrdersTable *o,
             const struct LineItemTable *l, const char *algorithm, struct joinconfig_t *config) {
    check_algorithm(algorithm);
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    // only the columns Q3 reads travel to the device
    LineItemTable lq{};
    lq.numTuples = l->numTuples; lq.l_orderkey = l->l_orderkey; lq.l_shipdate = l->l_shipdate;
    CustomerTable cq{};
    cq.numTuples = c->numTuples; cq.c_custkey = c->c_custkey; cq.c_mktsegment = c->c_mktsegment;
    b200_tpch_stats_t s{};
    if (upload_tables(&lq, o, &cq, nullptr) || q3_device(&s)) die_tpch("tpch_q3");
    fill_result(result, s, config);
}

void tpch_q12(struct result_t *result, const struct LineItemTable *l, const struct OrdersTable *o,
              const char *algorithm, struct joinconfig_t *config) {
    check_algorithm(algorithm);
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    LineItemTable lq{};
    lq.numTuples = l->numTuples; lq.l_orderkey = l->l_orderkey; lq.l_shipmode = l->l_shipmode;
    lq.l_commitdate = l->l_commitdate; lq.l_shipdate = l->l_shipdate; lq.l_receiptdate = l->l_receiptdate;
    OrdersTable oq{};
    oq.numTuples = o->numTuples; oq.o_orderkey = o->o_orderkey;
    b200_tpch_stats_t s{};
    if (upload_tables(&lq, &oq, nullptr, nullptr) || q12_device(&s)) die_tpch("tpch_q12");
    fill_result(result, s, config);
}

void tpch_q19(struct result_t *result, const struct LineItemTable *l, const struct PartTable *p,
              const char *algorithm, struct joinconfig_t *config) {
    check_algorithm(algorithm);
    std::lock_guard<std::recursive_mutex> lk(t_mu);
    LineItemTable lq{};
    lq.numTuples = l->numTuples; lq.l_orderkey = l->l_orderkey; lq.l_partkey = l->l_partkey; lq.l_quantity = l->l_quantity;
    lq.l_shipmode = l->l_shipmode; lq.l_shipinstruct = l->l_shipinstruct;
    b200_tpch_stats_t s{};
    if (upload_tables(&lq, nullptr, nullptr, p) || q19_device(&s)) die_tpch("tpch_q19");
    fill_result(result, s, config);
}

}  // extern "C"
