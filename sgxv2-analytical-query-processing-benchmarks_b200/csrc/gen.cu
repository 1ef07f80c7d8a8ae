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

int set_rowid_payload_device(row_t *d_rel, uint64_t row_begin, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    set_rowid_payload_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<uint2 *>(d_rel), row_begin, n);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace aqp
