// scanbench.cu — micro-benchmark behind the scan kernels' load path (standalone: nvcc scanbench.cu -o scanbench).
// Question: how should a persistent CTA stream a uint8 column so that HBM stays saturated while every 32 bytes cost
// ~60 ALU/FMA instructions (the range predicate)? Every variant evaluates the predicate of csrc/scan.cu on the whole
// column and counts the matches (so the ALU load is the real one) and reads each byte exactly once:
//   warp16k   every warp owns a contiguous 16 KiB piece of the CTA's tile, 1 KiB per step, 4 x 256-bit loads
//             per lane in flight on a rolling basis (the fused row-id kernel's first layout)
//   interl    same loads, but step j of warp w reads KiB (j * warps + w) of the tile: the CTA's loads of one step are
//             adjacent in memory
//   cta16k    the CTA reads 16 KiB per step with 4 x 128-bit loads per thread (bitvector_scan_kernel's layout)
//   tma       one elected thread feeds a ring of 16 KiB stages with cp.async.bulk (TMA), the warps read the stages
//             from shared memory with 128-bit loads (the north-star's "TMA bulk loads")
// Output: GB/s of column bytes per variant.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct Pred { uint32_t addK, hiH, one, neg_one; };
__device__ __forceinline__ uint32_t mad_fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t inrange_msb(uint32_t x, const Pred &p) {   // variant lo < 128, hi < 128
    const uint32_t H = 0x80808080u;
    const uint32_t b = x & ~H;
    return ~x & mad_fma(b, p.one, p.addK) & mad_fma(b, p.neg_one, p.hiH) & H;
}
__device__ __forceinline__ uint32_t mask16(uint4 v, const Pred &p) {
    uint32_t g0 = __dp4a(inrange_msb(v.x, p), 0x08040201u, 0u);
    g0 = __dp4a(inrange_msb(v.y, p), 0x80402010u, g0);
    uint32_t g1 = __dp4a(inrange_msb(v.z, p), 0x08040201u, 0u);
    g1 = __dp4a(inrange_msb(v.w, p), 0x80402010u, g1);
    return __umulhi(g1 * 256u + g0, 1u << 25);
}
struct U32x8 { uint32_t w[8]; };
__device__ __forceinline__ U32x8 ld256(const void *p) {
    U32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld128(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t mask32(const U32x8 &v, const Pred &p) {
    return mask16(make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]), p) | (mask16(make_uint4(v.w[4], v.w[5], v.w[6], v.w[7]), p) << 16);
}

constexpr int kRounds = 16;
template <int kWarps, bool kInterleave>
__global__ void __launch_bounds__(kWarps * 32, 1024 / (kWarps * 32))
k_warp(const uint8_t *in, size_t n, Pred p, unsigned long long *count) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile_bytes = (size_t) kWarps * kRounds * 1024;
    uint32_t cnt = 0;
    for (size_t tile = blockIdx.x; tile * tile_bytes < n; tile += gridDim.x) {
        const uint8_t *base = in + tile * tile_bytes;
        auto addr = [&](int r) {
            return base + (kInterleave ? ((size_t) r * kWarps + warp) * 1024 : ((size_t) warp * kRounds + r) * 1024) + lane * 32;
        };
        U32x8 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = ld256(addr(j));
#pragma unroll 1
        for (int b = 0; b < kRounds; b += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t m = mask32(v[j], p);
                if (b + 4 < kRounds) v[j] = ld256(addr(b + 4 + j));
                cnt += __popc(m);
            }
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(count, (unsigned long long) cnt);
}

__global__ void __launch_bounds__(256)
k_cta16k(const uint4 *in, size_t nvec, Pred p, unsigned long long *count) {
    const size_t ntiles = nvec / 1024;
    uint32_t c = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * 1024 + threadIdx.x;
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = ld128(in + base + j * 256);
#pragma unroll
        for (int j = 0; j < 4; ++j) c += __popc(mask16(v[j], p));
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long) c);
}

// ---- TMA ring
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// kWarps consumer warps + 1 producer warp; stage = kWarps KiB (1 KiB per consumer warp), kStages stages
template <int kWarps, int kStages>
__global__ void __launch_bounds__((kWarps + 1) * 32)
k_tma(const uint8_t *in, size_t n, Pred p, unsigned long long *count) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ uint64_t full[kStages], empty[kStages];
    constexpr uint32_t kStageBytes = kWarps * 1024;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nstages_total = n / kStageBytes;   // column split into stage-sized pieces, dealt round-robin to CTAs in runs
    const size_t per_cta = (nstages_total + gridDim.x - 1) / gridDim.x;
    const size_t s_begin = (size_t) blockIdx.x * per_cta, s_end = s_begin + per_cta < nstages_total ? s_begin + per_cta : nstages_total;
    if (warp == kWarps) {
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t q = s_begin; q < s_end; ++q, ++it) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                if (it >= (uint32_t) kStages) mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], kStageBytes);
                tma_load_1d(ring + (size_t) s * kStageBytes, in + q * kStageBytes, kStageBytes, &full[s]);
            }
        }
        return;
    }
    uint32_t cnt = 0, it = 0;
    for (size_t q = s_begin; q < s_end; ++q, ++it) {
        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
        mbar_wait(&full[s], ph);
        const uint4 *src = reinterpret_cast<const uint4 *>(ring + (size_t) s * kStageBytes + warp * 1024);
        const uint4 a = src[lane], b = src[lane + 32];   // conflict-free: 32 lanes x 16 B consecutive
        cnt += __popc(mask16(a, p)) + __popc(mask16(b, p));
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(count, (unsigned long long) cnt);
}

__global__ void k_fill(uint8_t *d, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) d[i] = (uint8_t) i;
}

int main(int argc, char **argv) {
    const size_t n = (size_t) 1 << (argc > 1 ? atoi(argv[1]) : 30);
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    uint8_t *d;
    unsigned long long *cnt;
    CK(cudaMalloc(&d, n));
    CK(cudaMalloc(&cnt, 8));
    k_fill<<<148 * 8, 256>>>(d, n);
    Pred p{0x01010101u * 0x80u, 0x01010101u * (0x80u | 26u), 1u, 0xffffffffu};   // [0, 26]
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](const char *name, auto launch) {
        unsigned long long h = 0;
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaMemset(cnt, 0, 8));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(&h, cnt, 8, cudaMemcpyDeviceToHost));
        printf("%-28s %8.4f ms  %7.1f GB/s  count/rep %llu (expect %llu)\n", name, ms / reps, n / (ms / reps) / 1e6, h / reps,
               (unsigned long long) (n / 256 * 27));
    };
    run("warp16k 16 warps x2/SM", [&] { k_warp<16, false><<<296, 512>>>(d, n, p, cnt); });
    run("interl  16 warps x2/SM", [&] { k_warp<16, true><<<296, 512>>>(d, n, p, cnt); });
    run("warp16k 8 warps x4/SM", [&] { k_warp<8, false><<<592, 256>>>(d, n, p, cnt); });
    run("interl  8 warps x4/SM", [&] { k_warp<8, true><<<592, 256>>>(d, n, p, cnt); });
    run("warp16k 32 warps x1/SM", [&] { k_warp<32, false><<<148, 1024>>>(d, n, p, cnt); });
    run("interl  32 warps x1/SM", [&] { k_warp<32, true><<<148, 1024>>>(d, n, p, cnt); });
    run("cta16k 256 thr x8/SM", [&] { k_cta16k<<<148 * 8, 256>>>(reinterpret_cast<const uint4 *>(d), n / 16, p, cnt); });
    {
        auto k = k_tma<15, 4>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 15 * 1024 * 4));
        run("tma 15+1 warps x2/SM 4 st", [&] { k<<<296, 512, 15 * 1024 * 4>>>(d, n, p, cnt); });
    }
    {
        auto k = k_tma<15, 6>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 15 * 1024 * 6));
        run("tma 15+1 warps x2/SM 6 st", [&] { k<<<296, 512, 15 * 1024 * 6>>>(d, n, p, cnt); });
    }
    {
        auto k = k_tma<31, 4>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 31 * 1024 * 4));
        run("tma 31+1 warps x1/SM 4 st", [&] { k<<<148, 1024, 31 * 1024 * 4>>>(d, n, p, cnt); });
    }
    {
        auto k = k_tma<7, 4>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 1024 * 4));
        run("tma 7+1 warps x4/SM 4 st", [&] { k<<<592, 256, 7 * 1024 * 4>>>(d, n, p, cnt); });
    }
    {
        auto k = k_tma<7, 8>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 1024 * 8));
        run("tma 7+1 warps x4/SM 8 st", [&] { k<<<592, 256, 7 * 1024 * 8>>>(d, n, p, cnt); });
    }
    return 0;
}
