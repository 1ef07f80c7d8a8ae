// build_probe.cu — bucket-chained hash join of co-partitions in shared memory (sm_100a).
//
// Replaces bucket_chaining_join (Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp:359-458) and the
// join-task queue that feeds it (:1174-1222, :818-836, :1320-1334):
//   * N = next_pow2(|R_p|) buckets, hash = the key bits ABOVE all radix bits,
//     (key >> num_radix_bits) & (N-1)                                   (:373-378)
//   * build: next[i] = bucket[h]; bucket[h] = i+1 (1-based chain heads)    (:386-412) — done here by
//     all threads at once with an atomic exchange, which yields the same chains in another order
//   * probe: walk the chain, compare keys, count / emit {key, Rpayload, Spayload}  (:428-447)
//   * co-partitions where R or S is empty are skipped                      (:1196,:820)
// A work item is (co-partition, chunk of its S side): a CTA loads the R side into shared memory,
// builds bucket/next there and streams its S chunk past it. Splitting S by chunks keeps every SM
// busy when partitions are few (small R) or one partition is huge (Zipf-skewed S); R sides larger
// than the table capacity are processed in several build rounds.
#include "common.cuh"
#include "join_internal.cuh"
#include "block_scan.cuh"

namespace aqp {

// ---------------------------------------------------------------------------------------------
// work list: items[i] = {partition, S-chunk}
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanBlock)
join_items_kernel(const uint32_t *__restrict__ offR, const uint32_t *__restrict__ endR, const uint32_t *__restrict__ offS,
                  const uint32_t *__restrict__ endS, uint32_t nparts, uint32_t *__restrict__ item_start,
                  uint2 *__restrict__ items) {
    auto chunks_of = [&](uint32_t p) {
        uint32_t nr = endR[p] - offR[p], ns = endS[p] - offS[p];
        return (nr > 0 && ns > 0) ? (ns + kProbeChunk - 1) / kProbeChunk : 0u;
    };
    uint32_t total = block_exclusive_scan(nparts, chunks_of, [&](uint32_t p, uint32_t v) { item_start[p] = v; });
    if (threadIdx.x == 0) item_start[nparts] = total;
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < nparts; p += kScanBlock) {
        uint32_t b = item_start[p], c = chunks_of(p);
        for (uint32_t k = 0; k < c; ++k) items[b + k] = make_uint2(p, k);
    }
}

// partition p of a relation is [off[p], end[p]); end = nullptr means the partitions are dense: end[p] = off[p + 1]
int join_items_device(const uint32_t *d_offR, const uint32_t *d_offS, uint32_t nparts, uint32_t *d_item_start,
                      uint2 *d_items, cudaStream_t st, const uint32_t *d_endR, const uint32_t *d_endS) {
    join_items_kernel<<<1, kScanBlock, 0, st>>>(d_offR, d_endR ? d_endR : d_offR + 1, d_offS, d_endS ? d_endS : d_offS + 1,
                                                nparts, d_item_start, d_items);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// build + probe
// ---------------------------------------------------------------------------------------------
constexpr int kProbeUnroll = 4;   // S tuples per thread and round; two rounds are in flight (double buffering)
constexpr uint32_t kChainFlag = 0x80000000u;   // set in a bucket head when its chain holds more than one tuple
constexpr size_t kJoinSmemBytes = (size_t) kBuildCap * (sizeof(uint2) + sizeof(uint32_t) + sizeof(uint16_t));

// pull a 128-byte line into L2 ahead of the loads that will use it: the registers only cover two S rounds in flight,
// the L2 prefetch turns the HBM latency of the rounds after those (and of the next item's R side) into L2 latency
__device__ __forceinline__ void prefetch_l2(const void *p) {
#ifndef AQP_NO_PROBE_PREFETCH
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
#ifndef AQP_PROBE_PREFETCH_ROUNDS
#define AQP_PROBE_PREFETCH_ROUNDS 3
#endif
constexpr uint32_t kPrefetchRounds = AQP_PROBE_PREFETCH_ROUNDS;   // S rounds (16 KiB each) prefetched ahead of the register pipeline

template <bool kMaterialize>
__global__ void __launch_bounds__(kJoinThreads, 2)
build_probe_kernel(const uint2 *__restrict__ R, const uint32_t *__restrict__ offR, const uint32_t *__restrict__ endR,
                   const uint2 *__restrict__ S, const uint32_t *__restrict__ offS, const uint32_t *__restrict__ endS,
                   const uint32_t *__restrict__ item_start,
                   const uint2 *__restrict__ items, uint32_t nparts, uint32_t hash_shift,
                   JoinResult *__restrict__ res, output_triple_t *__restrict__ out, unsigned long long out_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2 *rt = reinterpret_cast<uint2 *>(smem_raw);                                     // R tuples
    uint32_t *bucket = reinterpret_cast<uint32_t *>(rt + kBuildCap);                     // chain heads (1-based)
    uint16_t *next = reinterpret_cast<uint16_t *>(bucket + kBuildCap);                   // chain links (1-based)

    __shared__ uint32_t s_collisions;
    __shared__ uint32_t s_wsum[kJoinThreads / 32];   // materialising probe: per-warp match counts of a round
    __shared__ unsigned long long s_obase;          // and the round's first output slot
    const uint32_t nitems = item_start[nparts];
    unsigned long long matches = 0, checksum = 0, keysum = 0;

    for (uint32_t it = blockIdx.x; it < nitems; it += gridDim.x) {
        uint32_t m32 = 0;   // matches of this thread in this item (<= 64 probe tuples x 8192 build tuples): one register
        const uint2 item = items[it];
        const uint32_t p = item.x;
        const uint32_t rbeg = offR[p], rend = endR[p];
        const uint32_t sbeg = offS[p] + item.y * kProbeChunk;
        const uint32_t send = min(sbeg + (uint32_t) kProbeChunk, endS[p]);

        constexpr uint32_t kRound = kJoinThreads * kProbeUnroll;
        // lines of S round `base` .. : 16 tuples per line, kRound / 16 lines per round
        auto prefetch_round = [&](uint32_t base) {
            if (threadIdx.x < kRound / 16) {
                const uint32_t i = base + threadIdx.x * 16;
                if (i < send) prefetch_l2(S + i);
            }
        };
        if (it + gridDim.x < nitems) {   // the next item's R side (one line per thread covers 8192 tuples)
            const uint2 nx = items[it + gridDim.x];
            const uint32_t nrb = offR[nx.x], nre = endR[nx.x];
            for (uint32_t i = nrb + threadIdx.x * 16; i < nre && i < nrb + kBuildCap; i += kJoinThreads * 16) prefetch_l2(R + i);
        }
        auto load_round = [&](uint32_t base, uint2 (&dst)[kProbeUnroll]) {
#pragma unroll
            for (int j = 0; j < kProbeUnroll; ++j) {
                uint32_t i = base + j * kJoinThreads + threadIdx.x;
                if (i < send) dst[j] = ld_stream_v2(S + i);
            }
        };

        for (uint32_t rb = rbeg; rb < rend; rb += kBuildCap) {
            const uint32_t nr = min((uint32_t) kBuildCap, rend - rb);
            uint32_t N = 1;
            while (N < nr) N <<= 1;
            const uint32_t hmask = N - 1;
            // the first round of S tuples is put in flight BEFORE the build so its latency hides behind it
            uint2 sn[kProbeUnroll] = {};
            load_round(sbeg, sn);
#pragma unroll
            for (uint32_t r = 1; r <= kPrefetchRounds; ++r) prefetch_round(sbeg + r * kRound);
            // all R tuples of this thread are requested up front, then inserted
            constexpr int kBuildPerThread = kBuildCap / kJoinThreads;
            uint2 rv[kBuildPerThread];
#pragma unroll
            for (int k = 0; k < kBuildPerThread; ++k) {
                uint32_t i = k * kJoinThreads + threadIdx.x;
                if (i < nr) rv[k] = R[rb + i];
            }
            __syncthreads();   // previous round's probe done before the table is cleared
            constexpr uint32_t kRoundS = kJoinThreads * kProbeUnroll;
            if (!kMaterialize && N >= 2) {
                // ---- collision-free fast path (what a dense primary key gives: every R tuple of the co-partition
                // hashes to its own bucket). The table is direct-mapped: tuple -> rt[hash]; an empty slot holds a key
                // of ANOTHER bucket, which no probe key of this bucket can equal, so the probe is ONE 8-byte
                // shared-memory load and a compare — no chain head, no link (9.5 -> 6 wavefronts per 32 probes on
                // the shared-memory pipe that bounds this kernel, profiles/r01_ncu_join_kernels_final.md).
                // Any collision during the build sends the round to the chained path below.
                if (threadIdx.x == 0) s_collisions = 0;
                for (uint32_t i = threadIdx.x; i < N; i += kJoinThreads) rt[i].x = (i ^ 1u) << hash_shift;
                __syncthreads();
                bool clash = false;
#pragma unroll
                for (int k = 0; k < kBuildPerThread; ++k) {
                    uint32_t i = k * kJoinThreads + threadIdx.x;
                    if (i < nr) {
                        const uint32_t h = (rv[k].x >> hash_shift) & hmask;
                        const uint32_t empty = (h ^ 1u) << hash_shift;
                        if (atomicCAS(&rt[h].x, empty, rv[k].x) == empty)
                            rt[h].y = rv[k].y;
                        else
                            clash = true;
                    }
                }
                if (clash) s_collisions = 1;
                __syncthreads();
                if (s_collisions == 0) {
                    for (uint32_t base = sbeg; base < send; base += kRoundS) {
                        uint2 s[kProbeUnroll];
#pragma unroll
                        for (int j = 0; j < kProbeUnroll; ++j) s[j] = sn[j];
                        if (base + kRoundS < send) load_round(base + kRoundS, sn);
                        prefetch_round(base + (1 + kPrefetchRounds) * kRoundS);
                        // (one table load next to its compare per tuple: issuing the four loads of a round together
                        // costs registers this kernel does not have - 40 bytes of spills, 1.086 -> 1.132 ms - and buys
                        // nothing here, the shared-memory pipe is already the limit)
#pragma unroll
                        for (int j = 0; j < kProbeUnroll; ++j) {
                            if (base + j * kJoinThreads + threadIdx.x < send) {
                                const uint2 r = rt[(s[j].x >> hash_shift) & hmask];
                                if (r.x == s[j].x) {
                                    ++m32;
                                    checksum += (unsigned long long) r.y + s[j].y;
                                    keysum += s[j].x;
                                }
                            }
                        }
                    }
                    continue;   // next build round
                }
                // collisions: the table is rebuilt with chains (the R tuples are still in registers)
            }
            for (uint32_t i = threadIdx.x; i < N; i += kJoinThreads) bucket[i] = 0;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kBuildPerThread; ++k) {
                uint32_t i = k * kJoinThreads + threadIdx.x;
                if (i < nr) {
                    rt[i] = rv[k];
                    uint32_t *b = &bucket[(rv[k].x >> hash_shift) & hmask];
                    uint32_t old = atomicExch(b, i + 1);
                    next[i] = (uint16_t) old;             // low 16 bits: previous head (0 = end of chain)
                    if (old) atomicOr(b, kChainFlag);     // whoever is head in the end carries the flag
                }
            }
            __syncthreads();

            // uniform trip count so the materialising variant can use block-wide primitives; the next round's
            // loads are issued before the current round is probed (register double buffering)
            for (uint32_t base = sbeg; base < send; base += kRound) {
                uint2 s[kProbeUnroll];
#pragma unroll
                for (int j = 0; j < kProbeUnroll; ++j) s[j] = sn[j];
                if (base + kRound < send) load_round(base + kRound, sn);
                prefetch_round(base + (1 + kPrefetchRounds) * kRound);
                if (!kMaterialize) {
                    // the four chain heads first, then the chains: behind a per-tuple bounds branch the compiler kept each
                    // head load next to its use and every probe paid two dependent shared-memory round trips in a row.
                    // (Also issuing the four first chain elements together spills: the direct-mapped path above shares
                    // this kernel's 64 registers and loses 6 %.) Only a flagged chain (more than one tuple) is walked.
                    uint32_t heads[kProbeUnroll];
#pragma unroll
                    for (int j = 0; j < kProbeUnroll; ++j) heads[j] = bucket[(s[j].x >> hash_shift) & hmask];
#pragma unroll
                    for (int j = 0; j < kProbeUnroll; ++j) {
                        const bool valid = base + j * kJoinThreads + threadIdx.x < send;
                        uint32_t hit = valid ? heads[j] & 0xFFFFu : 0u;
                        if (!hit) continue;
                        uint2 r = rt[hit - 1];
                        if (r.x == s[j].x) {
                            ++m32;
                            checksum += (unsigned long long) r.y + s[j].y;
                            keysum += s[j].x;
                        }
                        if (heads[j] & kChainFlag) {
                            hit = next[hit - 1];
                            while (hit) {
                                r = rt[hit - 1];
                                if (r.x == s[j].x) {
                                    ++m32;
                                    checksum += (unsigned long long) r.y + s[j].y;
                                    keysum += s[j].x;
                                }
                                hit = next[hit - 1];
                            }
                        }
                    }
                }
                if (kMaterialize) {
                    // ---- materialise: ONE output reservation per CTA and round. (Reserving per warp and chain step -
                    // the first version - put 2.3 M atomics on the single out_count word for 73 M probe tuples and
                    // made this kernel 15x slower than the counting one: profiles/r02_tpch_q3_launches.md.)
                    // Phase 1 walks the chains, does the accounting and remembers each tuple's first match; a block-wide
                    // exclusive scan of the per-thread match counts gives every thread a dense slot range; phase 2 writes
                    // the single matches from registers and walks a chain again only for a duplicate build key.
                    uint32_t cnt[kProbeUnroll], first_pay[kProbeUnroll], c = 0;
#pragma unroll
                    for (int j = 0; j < kProbeUnroll; ++j) {
                        const bool valid = base + j * kJoinThreads + threadIdx.x < send;
                        uint32_t hit = valid ? (bucket[(s[j].x >> hash_shift) & hmask] & 0xFFFFu) : 0u;
                        cnt[j] = 0;
                        first_pay[j] = 0;
                        while (hit) {
                            const uint2 r = rt[hit - 1];
                            hit = next[hit - 1];
                            if (r.x == s[j].x) {
                                if (!cnt[j]) first_pay[j] = r.y;
                                ++cnt[j];
                                checksum += (unsigned long long) r.y + s[j].y;
                                keysum += s[j].x;
                            }
                        }
                        c += cnt[j];
                    }
                    m32 += c;
                    const uint32_t incl = warp_incl_scan(c);
                    if (lane_id() == 31) s_wsum[threadIdx.x >> 5] = incl;
                    __syncthreads();
                    if (threadIdx.x < 32) {
                        const uint32_t v = threadIdx.x < kJoinThreads / 32 ? s_wsum[threadIdx.x] : 0u;
                        const uint32_t iv = warp_incl_scan(v);
                        if (threadIdx.x < kJoinThreads / 32) s_wsum[threadIdx.x] = iv - v;
                        if (threadIdx.x == 31) s_obase = iv ? atomicAdd(&res->out_count, (unsigned long long) iv) : 0ull;
                    }
                    __syncthreads();
                    unsigned long long slot = s_obase + s_wsum[threadIdx.x >> 5] + (incl - c);
#pragma unroll
                    for (int j = 0; j < kProbeUnroll; ++j) {
                        if (!cnt[j]) continue;
                        output_triple_t t;
                        t.key = s[j].x;
                        t.Spayload = s[j].y;
                        if (cnt[j] == 1) {
                            t.Rpayload = first_pay[j];
                            if (slot < out_cap) out[slot] = t;
                            ++slot;
                            continue;
                        }
                        uint32_t hit = bucket[(s[j].x >> hash_shift) & hmask] & 0xFFFFu;   // duplicate build keys: walk again
                        while (hit) {
                            const uint2 r = rt[hit - 1];
                            hit = next[hit - 1];
                            if (r.x == s[j].x) {
                                t.Rpayload = r.y;
                                if (slot < out_cap) out[slot] = t;
                                ++slot;
                            }
                        }
                    }
                }
            }
        }
        matches += m32;
    }

    // block reduction of the three accumulators -> 3 global atomics per CTA
    __shared__ unsigned long long red[3][kJoinThreads / 32];
    matches = warp_sum(matches);
    checksum = warp_sum(checksum);
    keysum = warp_sum(keysum);
    __syncthreads();
    if (lane_id() == 0) {
        red[0][threadIdx.x >> 5] = matches;
        red[1][threadIdx.x >> 5] = checksum;
        red[2][threadIdx.x >> 5] = keysum;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int w = 0; w < kJoinThreads / 32; ++w) t += red[threadIdx.x][w];
        unsigned long long *dst = threadIdx.x == 0 ? &res->matches : (threadIdx.x == 1 ? &res->checksum : &res->keysum);
        if (t) atomicAdd(dst, t);
    }
}

int build_probe_device(const row_t *d_R, const uint32_t *d_offR, const row_t *d_S, const uint32_t *d_offS,
                       const uint32_t *d_item_start, const uint2 *d_items, uint32_t nparts, uint64_t max_items,
                       uint32_t hash_shift, JoinResult *d_res, output_triple_t *d_out, uint64_t out_cap,
                       cudaStream_t st, const uint32_t *d_endR, const uint32_t *d_endS) {
    if (!d_endR) d_endR = d_offR + 1;   // dense partitions
    if (!d_endS) d_endS = d_offS + 1;
    static unsigned attr_set = ~0u;   // device epoch the opt-in was made for (b200_shutdown + b200_init(other device) re-arms it)
    if (attr_set != g_device_epoch) {
        AQP_CUDA_OK(cudaFuncSetAttribute(build_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) kJoinSmemBytes));
        AQP_CUDA_OK(cudaFuncSetAttribute(build_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) kJoinSmemBytes));
        attr_set = g_device_epoch;
    }
    if (max_items == 0) return 0;
    int grid = (int) (max_items < (uint64_t) kNumSMs * 2 ? max_items : (uint64_t) kNumSMs * 2);
    const uint2 *R = reinterpret_cast<const uint2 *>(d_R), *S = reinterpret_cast<const uint2 *>(d_S);
    if (d_out)
        build_probe_kernel<true><<<grid, kJoinThreads, kJoinSmemBytes, st>>>(R, d_offR, d_endR, S, d_offS, d_endS, d_item_start,
                                                                             d_items, nparts, hash_shift, d_res, d_out, out_cap);
    else
        build_probe_kernel<false><<<grid, kJoinThreads, kJoinSmemBytes, st>>>(R, d_offR, d_endR, S, d_offS, d_endS, d_item_start,
                                                                              d_items, nparts, hash_shift, d_res, nullptr, 0);
    AQP_LAUNCHED();
    AQP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace aqp
