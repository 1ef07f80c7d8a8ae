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

/* ---- the remaining SIMD512 variants on the 8-bit column ------------------------------ */

/* SIMD512::sum (:34-86): sum of the values in range over the n/64 whole blocks. */
uint64_t oracle_scan_sum(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n) {
    uint64_t s = 0;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks * 64; ++i)
        if (in[i] >= lo && in[i] <= hi) s += in[i];
    return s;
}

/* SIMD512::scan (:89-150): the matching values, widened to uint32, in input order; returns the count. */
uint64_t oracle_value_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint32_t *out) {
    uint64_t w = 0;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks * 64; ++i)
        if (in[i] >= lo && in[i] <= hi) out[w++] = in[i];
    return w;
}

/* Code range of a value predicate over a sorted 256-entry dictionary, exactly as dict_scan_8bit_64bit derives it
 * (:297-305): low code = index of the first entry >= predicate_low (std::find_if), high code = index of the first
 * entry > predicate_high at or after it, minus one; both are then narrowed to uint8. The narrowing is kept as is:
 * a predicate above every entry gives low = (uint8) 256 = 0 and high = 255, one below every entry gives
 * high = (uint8) -1 = 255 — in both cases the reference selects everything. */
void oracle_dict_code_range(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, uint8_t *lo, uint8_t *hi) {
    long l = 0;
    while (l < 256 && !(dict[l] >= predicate_low)) ++l;
    long h = l;
    while (h < 256 && !(dict[h] > predicate_high)) ++h;
    h -= 1;
    *lo = (uint8_t) l;
    *hi = (uint8_t) (h & 0xff);
}

/* SIMD512::dict_scan_8bit_64bit (:289-336, cut = true): dict[code] of every code in the code range, in input
 * order; returns the number of values written. */
uint64_t oracle_dict_scan_8_64(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint8_t *in,
                               size_t n, int64_t *out) {
    uint8_t lo, hi;
    oracle_dict_code_range(predicate_low, predicate_high, dict, &lo, &hi);
    uint64_t w = 0;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks * 64; ++i)
        if (in[i] >= lo && in[i] <= hi) out[w++] = dict[in[i]];
    return w;
}

/* SIMD512::explicit_index_scan (:152-208): for every 64-value block i and every byte group j (8 values) with a
 * match, the matching lanes of index register i + j (8 uint64 each) are compress-stored - block i's group j reads
 * index[(i + j) * 8 + k], NOT (8 i + j): restated as written. Returns the count. */
uint64_t oracle_explicit_index_scan(uint8_t lo, uint8_t hi, const uint64_t *index, const uint8_t *in, size_t n, uint64_t *out) {
    uint64_t w = 0;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i)
        for (size_t j = 0; j < 8; ++j)
            for (size_t k = 0; k < 8; ++k) {
                uint8_t v = in[64 * i + 8 * j + k];
                if (v >= lo && v <= hi) out[w++] = index[(i + j) * 8 + k];
            }
    return w;
}

/* Code range of dict_scan_16bit_64bit / dict_scan_32bit_64bit (:539-547,:585-593): std::find_if for the first entry
 * >= predicate_low, then for the first entry > predicate_high from there, minus one; BOTH are narrowed with
 * static_cast<uint16_t> - in the 32-bit scan too - and then compared as unsigned codes of the column's width. */
void oracle_wide_code_range(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, size_t dict_size, uint32_t *lo,
                            uint32_t *hi) {
    size_t l = 0;
    while (l < dict_size && !(dict[l] >= predicate_low)) ++l;
    size_t h = l;
    while (h < dict_size && !(dict[h] > predicate_high)) ++h;
    *lo = (uint16_t) l;
    *hi = (uint16_t) (h - 1);
}

/* SIMD512::dict_scan_16bit_64bit (:531-577): 65536-entry dictionary, whole registers of 32 codes. */
uint64_t oracle_dict_scan_16_64(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint16_t *in, size_t n,
                                int64_t *out) {
    uint32_t lo, hi;
    oracle_wide_code_range(predicate_low, predicate_high, dict, (size_t) 1 << 16, &lo, &hi);
    uint64_t w = 0;
    size_t m = n / 32 * 32;
    for (size_t i = 0; i < m; ++i)
        if (in[i] >= lo && in[i] <= hi) out[w++] = dict[in[i]];
    return w;
}

/* SIMD512::dict_scan_32bit_64bit (:579-622): dict_size entries, whole registers of 16 codes. */
uint64_t oracle_dict_scan_32_64(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, size_t dict_size,
                                const uint32_t *in, size_t n, int64_t *out) {
    uint32_t lo, hi;
    oracle_wide_code_range(predicate_low, predicate_high, dict, dict_size, &lo, &hi);
    uint64_t w = 0;
    size_t m = n / 16 * 16;
    for (size_t i = 0; i < m; ++i)
        if (in[i] >= lo && in[i] <= hi) out[w++] = dict[in[i]];
    return w;
}
