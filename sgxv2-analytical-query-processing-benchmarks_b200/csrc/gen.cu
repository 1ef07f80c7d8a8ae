// gen.cu — relation generators that write straight into HBM (sm_100a).
//
// Device-side counterpart of Join-Benchmarks/lib/AppUtilities/src/generator.cpp:
//   create_relation_pk :352-377 (random_unique_gen :143-153 + knuth_shuffle :99-109)
//   create_relation_fk :474-512
// The reference's shuffle is a strictly sequential glibc-rand() Fisher-Yates walk; here every
// tuple is computed independently from its row index with a keyed bijection on [0, n), so the
// key *distribution* is the same (PK: a permutation of 1..n; FK: floor(n/maxid) independent
// permutations of 1..maxid laid end to end, then 1..rem) but the permutation itself differs.
// Bit-identical reference inputs come from the host generators in host_gen.cpp.
#include "common.cuh"
#include "join_internal.cuh"

namespace aqp {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// keyed bijection on [0, 2^b): odd multiplies push entropy up, xor-shifts pull it back down
__device__ __forceinline__ uint64_t permute_pow2(uint64_t x, uint32_t b, uint64_t k) {
    const uint64_t mask = b >= 64 ? ~0ull : ((1ull << b) - 1);
    const uint32_t s1 = b > 1 ? b / 2 : 1, s2 = b > 2 ? (b + 2) / 3 : 1;
    x = (x * 0x9E3779B97F4A7C15ull + k) & mask;
    x ^= x >> s1;
    x = (x * 0xD6E8FEB86659FD93ull + (k >> 17)) & mask;
    x ^= x >> s2;
    x = (x * 0xCA5A826395121157ull + (k >> 31)) & mask;
    x ^= x >> s1;
    x = (x * 0x2545F4914F6CDD1Dull) & mask;
    x ^= x >> s2;
    return x;
}

// bijection on [0, n) by cycle walking inside the enclosing power of two
__device__ __forceinline__ uint64_t permute(uint64_t x, uint64_t n, uint32_t b, uint64_t k) {
    do {
        x = permute_pow2(x, b, k);
    } while (x >= n);
    return x;
}

static uint32_t ceil_log2(uint64_t n) {
    uint32_t b = 0;
    while (b < 63 && (1ull << b) < n) ++b;
    return b;
}

__global__ void gen_pk_kernel(uint2 *rel, uint64_t n_total, uint32_t b, uint64_t row_begin, uint64_t n, uint64_t key) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t row = row_begin + i;
        rel[i] = make_uint2((uint32_t) (permute(row, n_total, b, key) + 1), (uint32_t) row);
    }
}

__global__ void gen_fk_kernel(uint2 *rel, uint64_t n_total, uint64_t maxid, uint32_t b_full, uint32_t b_rem,
                              uint64_t row_begin, uint64_t n, uint64_t seed) {
    const uint64_t iters = n_total / maxid, rem = n_total % maxid;
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t row = row_begin + i;
        uint64_t blk = row / maxid, j = row - blk * maxid;
        uint64_t k = mix64(seed ^ mix64(blk));
        uint64_t v = blk < iters ? permute(j, maxid, b_full, k) : permute(j, rem, b_rem, k);
        rel[i] = make_uint2((uint32_t) (v + 1), (uint32_t) row);
    }
}

__global__ void set_rowid_payload_kernel(uint2 *rel, uint64_t row_begin, uint64_t n) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        rel[i].y = (uint32_t) (row_begin + i);
}

// ---------------------------------------------------------------------------------------------
// Zipf (Join-Benchmarks/lib/AppUtilities/src/genzipf.cpp): alphabet = random permutation of 1..maxid
// (:33-51, here the keyed bijection), cumulative distribution table lut[i] = sum_{k<=i+1} k^-z / total
// (:58-83), per tuple a uniform r in [0,1) and a binary search for the first lut[pos] >= r (:113-137).
// The table is built on the device with a three-phase fp64 scan (block sums, scan of sums, rescan).
// ---------------------------------------------------------------------------------------------
constexpr int kLutBlock = 256, kLutPerThread = 16, kLutChunk = kLutBlock * kLutPerThread;

__device__ __forceinline__ double block_sum_f64(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < kLutBlock / 32; ++w) t += sh[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kLutBlock) zipf_chunk_sums_kernel(uint64_t n, double z, double *sums) {
    __shared__ double sh[kLutBlock / 32];
    uint64_t base = (uint64_t) blockIdx.x * kLutChunk + (uint64_t) threadIdx.x * kLutPerThread;
    double v = 0;
    for (int k = 0; k < kLutPerThread; ++k)
        if (base + k < n) v += 1.0 / pow((double) (base + k + 1), z);
    double t = block_sum_f64(v, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = t;
}

__global__ void zipf_scan_sums_kernel(double *sums, uint32_t nchunks, double *total) {   // one thread: <= 2^15 chunks
    double run = 0;
    for (uint32_t i = 0; i < nchunks; ++i) {
        double c = sums[i];
        sums[i] = run;
        run += c;
    }
    *total = run;
}

__global__ void __launch_bounds__(kLutBlock)
zipf_lut_kernel(uint64_t n, double z, const double *sums, const double *total, double *lut) {
    __shared__ double sh[kLutBlock / 32];
    __shared__ double wsum[kLutBlock / 32];
    uint64_t base = (uint64_t) blockIdx.x * kLutChunk + (uint64_t) threadIdx.x * kLutPerThread;
    double w[kLutPerThread], v = 0;
    for (int k = 0; k < kLutPerThread; ++k) {
        w[k] = base + k < n ? 1.0 / pow((double) (base + k + 1), z) : 0.0;
        v += w[k];
    }
    // exclusive prefix of the per-thread sums inside the block
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    double before = 0;
    for (int k = 0; k < (int) (threadIdx.x >> 5); ++k) before += wsum[k];
    (void) sh;
    double run = sums[blockIdx.x] + before + incl - v;
    const double inv = 1.0 / *total;
    for (int k = 0; k < kLutPerThread; ++k) {
        run += w[k];
        if (base + k < n) lut[base + k] = run * inv;
    }
}

__global__ void gen_zipf_kernel(uint2 *rel, const double *__restrict__ lut, uint64_t maxid, uint32_t b,
                                uint64_t row_begin, uint64_t n, uint64_t seed, uint64_t alpha_key) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t row = row_begin + i;
        double r = (double) (mix64(seed ^ mix64(row)) >> 11) * (1.0 / 9007199254740992.0);   // [0,1), 53 bits
        uint64_t pos;
        if (lut[0] >= r) {
            pos = 0;
        } else {
            uint64_t left = 0, right = maxid - 1;
            while (right - left > 1) {
                uint64_t m = (left + right) / 2;
                if (lut[m] < r) left = m; else right = m;
            }
            pos = right;
        }
        rel[i] = make_uint2((uint32_t) (permute(pos, maxid, b, alpha_key) + 1), (uint32_t) row);
    }
}

static uint64_t host_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

int gen_pk_device(row_t *d_rel, uint64_t n_total, uint64_t row_begin, uint64_t n, uint64_t seed, cudaStream_t st) {
    if (n == 0) return 0;
    if (n_total > 0xFFFFFFFFull) {
        set_error("gen_pk: keys are 32-bit, n_total must be < 2^32");
        return -1;
    }
    gen_pk_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<uint2 *>(d_rel), n_total, ceil_log2(n_total), row_begin, n,
                                              host_mix64(seed));
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int gen_fk_device(row_t *d_rel, uint64_t n_total, uint64_t maxid, uint64_t row_begin, uint64_t n, uint64_t seed,
                  cudaStream_t st) {
    if (n == 0) return 0;
    if (maxid == 0 || maxid > 0xFFFFFFFFull) {
        set_error("gen_fk: maxid must be in [1, 2^32)");
        return -1;
    }
    uint64_t rem = n_total % maxid;
    gen_fk_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<uint2 *>(d_rel), n_total, maxid, ceil_log2(maxid),
                                              ceil_log2(rem ? rem : 1), row_begin, n, seed);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

int gen_zipf_device(row_t *d_rel, uint64_t maxid, double z, uint64_t row_begin, uint64_t n, uint64_t seed,
                    cudaStream_t st) {
    if (n == 0) return 0;
    if (maxid == 0 || maxid > 0xFFFFFFFFull || maxid > ((uint64_t) 1 << 15) * kLutChunk) {
        set_error("gen_zipf: maxid must be in [1, 2^27]");
        return -1;
    }
    const uint32_t nchunks = (uint32_t) ((maxid + kLutChunk - 1) / kLutChunk);
    double *lut = nullptr, *sums = nullptr;
    AQP_CUDA_OK(cudaMallocAsync(&lut, maxid * sizeof(double), st));
    AQP_CUDA_OK(cudaMallocAsync(&sums, ((size_t) nchunks + 1) * sizeof(double), st));
    zipf_chunk_sums_kernel<<<nchunks, kLutBlock, 0, st>>>(maxid, z, sums);
    AQP_LAUNCHED();
    zipf_scan_sums_kernel<<<1, 1, 0, st>>>(sums, nchunks, sums + nchunks);
    AQP_LAUNCHED();
    zipf_lut_kernel<<<nchunks, kLutBlock, 0, st>>>(maxid, z, sums, sums + nchunks, lut);
    AQP_LAUNCHED();
    gen_zipf_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<uint2 *>(d_rel), lut, maxid, ceil_log2(maxid), row_begin,
                                                n, seed, host_mix64(seed ^ 0x5a17f00dull));
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    AQP_CUDA_OK(cudaFreeAsync(lut, st));
    AQP_CUDA_OK(cudaFreeAsync(sums, st));
    return 0;
}

int set_rowid_payload_device(row_t *d_rel, uint64_t row_begin, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    set_rowid_payload_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<uint2 *>(d_rel), row_begin, n);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace aqp
