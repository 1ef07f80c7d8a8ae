/*
 * oracle_scan.c — scalar restatement of the reference AVX-512 uint8 column scans.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Reference: Scan-Micro-Benchmarks/shared_libraries/SimdScan/src/SIMD512.cpp.
 * Semantics restated: unsigned 8-bit inclusive range lo <= v <= hi (cmpge & cmple, :218-220);
 * only input_size/64 whole 64-value blocks are processed (:216,:234,:264); bit k of output word
 * i is value 64*i+k (mask64 LSB = lowest lane); row ids are size_t positions relative to `in`.
 */
#include "oracle.h"

uint64_t oracle_scan_count(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n) {
    uint64_t c = 0;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks * 64; ++i) c += (in[i] >= lo && in[i] <= hi);
    return c;
}

void oracle_bitvector_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint64_t *out) {
    size_t blocks = n / 64;
    for (size_t b = 0; b < blocks; ++b) {
        uint64_t m = 0;
        for (int k = 0; k < 64; ++k) {
            uint8_t v = in[b * 64 + k];
            m |= (uint64_t) (v >= lo && v <= hi) << k;
        }
        out[b] = m;
    }
}

uint64_t oracle_index_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint64_t *out) {
    uint64_t w = 0;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks * 64; ++i)
        if (in[i] >= lo && in[i] <= hi) out[w++] = i;
    return w;
}

uint64_t oracle_scalar_index_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint64_t *out) {
    uint64_t w = 0;
    for (size_t i = 0; i < n; ++i)
        if (in[i] >= lo && in[i] <= hi) out[w++] = i;
    return w;
}

void oracle_fill_tiled_column(uint8_t *data, size_t n) {
    /* Allocator.hpp:95-109 copies the 0..255 pattern num/256 times; a tail (n % 256) is left
     * uninitialised by the reference — the oracle zero-fills it. */
    size_t copies = n / 256;
    for (size_t i = 0; i < copies * 256; ++i) data[i] = (uint8_t) (i & 255);
    for (size_t i = copies * 256; i < n; ++i) data[i] = 0;
}
