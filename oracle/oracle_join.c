/*
 * oracle_join.c — single-threaded restatement of the reference RHO radix hash join.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Reference: Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp (+ include/radix/prj_params.h).
 *
 * What is restated bit-for-bit: the radix-bit / pass derivation (:295-337), the digit function
 * (key & MASK) >> R on raw key bits (:47), pass 1 on bits [0,b1) and pass 2 on [b1,b1+b2)
 * (:1118-1119,:1075-1076,:1262), skipping of empty co-partitions (:1196,:820), and the
 * bucket-chaining build/probe with N = next_pow2(|R_p|), hash on the bits above all radix bits
 * (:378), 1-based push-front chains (:386-412) and probe chain walk (:428-447).
 * What is NOT restated because it cannot change results: threads, queues, barriers, padding
 * between partitions (:339-345), timers.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* prj_params.h:58-66: L2_CACHE_SIZE 1280 KiB, L2_CACHE_TUPLES = L2_CACHE_SIZE / sizeof(row_t) */
#define L2_CACHE_TUPLES ((1280u * 1024u) / 8u)

uint32_t oracle_calc_num_radix_bits(uint64_t num_r, uint64_t nthreads) {
    uint64_t max_tuples_in_cache = L2_CACHE_TUPLES / 4;
    uint64_t parts = (num_r + max_tuples_in_cache - 1) / max_tuples_in_cache;
    if (parts < nthreads) parts = nthreads;
    uint32_t bits = 0;
    while ((1ull << bits) < parts) ++bits;
    return bits;
}

uint32_t oracle_calc_num_passes(uint32_t bits) { return bits <= 13 ? 1 : 2; }

void oracle_radix_partition(const oracle_row_t *in, uint64_t n, uint32_t shift, uint32_t bits,
                            oracle_row_t *out, uint64_t *offsets) {
    uint32_t fanout = 1u << bits;
    uint32_t mask = (fanout - 1) << shift;
    uint64_t *dst = (uint64_t *) calloc(fanout + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) dst[((in[i].key & mask) >> shift) + 1]++;
    for (uint32_t p = 0; p < fanout; ++p) dst[p + 1] += dst[p];
    memcpy(offsets, dst, (fanout + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) out[dst[(in[i].key & mask) >> shift]++] = in[i];
    free(dst);
}

static uint32_t next_pow_2(uint32_t v) {
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
    return v;
}

int64_t oracle_bucket_chaining_join(const oracle_row_t *R, uint64_t nR, const oracle_row_t *S, uint64_t nS,
                                    uint32_t num_radix_bits, uint64_t *checksum, uint64_t *keysum,
                                    oracle_triple_t *out, uint64_t out_cap, uint64_t *out_n) {
    uint32_t N = next_pow_2((uint32_t) nR);
    uint32_t mask = (N - 1) << num_radix_bits;
    uint32_t *next = (uint32_t *) malloc(sizeof(uint32_t) * (nR ? nR : 1));
    uint32_t *bucket = (uint32_t *) calloc(N ? N : 1, sizeof(uint32_t));
    int64_t matches = 0;
    for (uint32_t i = 0; i < nR;) {
        uint32_t idx = (R[i].key & mask) >> num_radix_bits;
        next[i] = bucket[idx];
        bucket[idx] = ++i;
    }
    for (uint64_t i = 0; i < nS; ++i) {
        uint32_t idx = (S[i].key & mask) >> num_radix_bits;
        for (uint32_t hit = bucket[idx]; hit > 0; hit = next[hit - 1]) {
            if (S[i].key == R[hit - 1].key) {
                ++matches;
                if (checksum) *checksum += (uint64_t) R[hit - 1].payload + (uint64_t) S[i].payload;
                if (keysum) *keysum += S[i].key;
                if (out && out_n) {
                    if (*out_n < out_cap) {
                        out[*out_n].key = S[i].key;
                        out[*out_n].Rpayload = R[hit - 1].payload;
                        out[*out_n].Spayload = S[i].payload;
                    }
                    ++*out_n;
                }
            }
        }
    }
    free(bucket);
    free(next);
    return matches;
}

int64_t oracle_rho(const oracle_row_t *R, uint64_t nR, const oracle_row_t *S, uint64_t nS,
                   int nthreads, int force_2_passes, uint64_t *checksum, uint64_t *keysum,
                   oracle_triple_t *out, uint64_t out_cap) {
    uint32_t bits = oracle_calc_num_radix_bits(nR, (uint64_t) nthreads);
    uint32_t passes = force_2_passes ? 2 : oracle_calc_num_passes(bits);
    uint32_t b1 = bits / passes, b2 = bits - b1;          /* :331-337 */
    uint32_t f1 = 1u << b1, f2 = 1u << b2;
    uint64_t cs = 0, ks = 0, out_n = 0;
    int64_t matches = 0;

    oracle_row_t *tR = (oracle_row_t *) malloc(sizeof(oracle_row_t) * (nR ? nR : 1));
    oracle_row_t *tS = (oracle_row_t *) malloc(sizeof(oracle_row_t) * (nS ? nS : 1));
    uint64_t *oR = (uint64_t *) malloc(sizeof(uint64_t) * (f1 + 1));
    uint64_t *oS = (uint64_t *) malloc(sizeof(uint64_t) * (f1 + 1));
    oracle_radix_partition(R, nR, 0, b1, tR, oR);
    oracle_radix_partition(S, nS, 0, b1, tS, oS);

    if (passes == 1) {
        for (uint32_t p = 0; p < f1; ++p) {
            uint64_t nr = oR[p + 1] - oR[p], ns = oS[p + 1] - oS[p];
            if (nr > 0 && ns > 0)
                matches += oracle_bucket_chaining_join(tR + oR[p], nr, tS + oS[p], ns, bits, &cs, &ks,
                                                       out, out_cap, &out_n);
        }
    } else {
        uint64_t *o2R = (uint64_t *) malloc(sizeof(uint64_t) * (f2 + 1));
        uint64_t *o2S = (uint64_t *) malloc(sizeof(uint64_t) * (f2 + 1));
        for (uint32_t p = 0; p < f1; ++p) {
            uint64_t nr = oR[p + 1] - oR[p], ns = oS[p + 1] - oS[p];
            if (!(nr > 0 && ns > 0)) continue;
            oracle_row_t *t2R = (oracle_row_t *) malloc(sizeof(oracle_row_t) * nr);
            oracle_row_t *t2S = (oracle_row_t *) malloc(sizeof(oracle_row_t) * ns);
            oracle_radix_partition(tR + oR[p], nr, b1, b2, t2R, o2R);
            oracle_radix_partition(tS + oS[p], ns, b1, b2, t2S, o2S);
            for (uint32_t q = 0; q < f2; ++q) {
                uint64_t nr2 = o2R[q + 1] - o2R[q], ns2 = o2S[q + 1] - o2S[q];
                if (nr2 > 0 && ns2 > 0)
                    matches += oracle_bucket_chaining_join(t2R + o2R[q], nr2, t2S + o2S[q], ns2, bits,
                                                           &cs, &ks, out, out_cap, &out_n);
            }
            free(t2R);
            free(t2S);
        }
        free(o2R);
        free(o2S);
    }
    free(tR); free(tS); free(oR); free(oS);
    if (checksum) *checksum = cs;
    if (keysum) *keysum = ks;
    return matches;
}
