/*
 * oracle/ — CPU restatement of the reference's RHO radix hash join, relation generators and
 * uint8 column scans.  TEST INFRASTRUCTURE ONLY: imported/linked solely by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, as the checker.
 * The product path (libb200aqp.so) never links or calls anything in this directory.
 *
 * Parity pin: every function here is checked (tests/test_oracle_*.py) against
 *   (1) the UNMODIFIED reference compiled from /root/reference into oracle/_ref/ (oracle/Makefile),
 *   (2) the committed fixtures in tests/golden/ that were generated from (1) by
 *       tests/golden/make_golden.py, and
 *   (3) the closed-form known answers of SURVEY.md §4 (the reference ships no join/scan KATs).
 *
 * All paths below are relative to /root/reference/.
 */
#ifndef AQP_ORACLE_H
#define AQP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Join-Benchmarks/lib/SharedHeaders/include/data-types.h:44-47 */
typedef struct { uint32_t key; uint32_t payload; } oracle_row_t;
/* data-types.h:62-66 */
typedef struct { uint32_t key; uint32_t Rpayload; uint32_t Spayload; } oracle_triple_t;

/* ---- generators (Join-Benchmarks/lib/AppUtilities/src/generator.cpp) ------------------- */
/* glibc srandom_r/random_r TYPE_3 restated (what srand()/rand() of generator.cpp:75-80,:19 run) */
void     oracle_srand(uint32_t seed);
int32_t  oracle_rand(void);
/* generator.cpp:143-153 + :99-109: keys 1..n, Sattolo-style shuffle j=(int64)(rand()/(RAND_MAX+1.0)*i) */
void oracle_gen_pk(oracle_row_t *rel, uint64_t n);
/* generator.cpp:474-512: floor(n/maxid) shuffled copies of 1..maxid (+ remainder 1..rem) */
void oracle_gen_fk(oracle_row_t *rel, uint64_t n, int64_t maxid);
/* generator.cpp:515-553 + :156-169 (random_unique_gen_maxid): stride-jump keys for `-l sel` */
void oracle_gen_fk_sel(oracle_row_t *rel, uint64_t n, int64_t maxid);
/* genzipf.cpp:58-83: CDF lookup table of sum 1/i^z normalised */
void oracle_zipf_lut(double *lut, uint32_t alphabet_size, double z);
/* genzipf.cpp:113-137: binary search of r in lut -> position */
uint32_t oracle_zipf_pos(const double *lut, uint32_t alphabet_size, double r);

/* ---- RHO (Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp) --------------------------- */
uint32_t oracle_calc_num_radix_bits(uint64_t num_r, uint64_t nthreads);   /* :295-317 */
uint32_t oracle_calc_num_passes(uint32_t bits);                           /* :319-329 */
/* one partitioning pass on key bits [shift, shift+bits): histogram (:617-623), exclusive prefix
 * (:886-915 with one thread, no padding), scatter (:659-665). hist has 2^bits+1 entries and
 * receives partition start offsets (hist[2^bits] = n). Stable. */
void oracle_radix_partition(const oracle_row_t *in, uint64_t n, uint32_t shift, uint32_t bits,
                            oracle_row_t *out, uint64_t *offsets);
/* bucket chaining build + probe of one co-partition (:359-458). Returns matches; adds to
 * checksum (sum Rpayload+Spayload, CHT convention Joins/include/cht/CHTJoin.hpp:174) and keysum. */
int64_t oracle_bucket_chaining_join(const oracle_row_t *R, uint64_t nR, const oracle_row_t *S, uint64_t nS,
                                    uint32_t num_radix_bits, uint64_t *checksum, uint64_t *keysum,
                                    oracle_triple_t *out, uint64_t out_cap, uint64_t *out_n);
/* whole join (:1369-1638 orchestration with one thread): bits/passes as the reference derives
 * them for `nthreads` (force_2_passes mirrors -DFORCE_2_PHASES, :1387-1392). */
int64_t oracle_rho(const oracle_row_t *R, uint64_t nR, const oracle_row_t *S, uint64_t nS,
                   int nthreads, int force_2_passes, uint64_t *checksum, uint64_t *keysum,
                   oracle_triple_t *out, uint64_t out_cap);

/* ---- scans (Scan-Micro-Benchmarks/shared_libraries/SimdScan/src/SIMD512.cpp) ------------ */
uint64_t oracle_scan_count(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n);            /* :7-32   */
void     oracle_bitvector_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint64_t *out); /* :210-222 */
uint64_t oracle_index_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint64_t *out);     /* :225-249 */
/* microbenchmarks/SimdScanMulti/shared/ScalarScan.hpp:8-20 (processes all n, no /64 truncation) */
uint64_t oracle_scalar_index_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint64_t *out);
uint64_t oracle_scan_sum(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n);                                /* :34-86   */
uint64_t oracle_value_scan(uint8_t lo, uint8_t hi, const uint8_t *in, size_t n, uint32_t *out);               /* :89-150  */
void     oracle_dict_code_range(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, uint8_t *lo,
                                uint8_t *hi);                                                                    /* :297-305 */
uint64_t oracle_dict_scan_8_64(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint8_t *in,
                               size_t n, int64_t *out);
uint64_t oracle_explicit_index_scan(uint8_t lo, uint8_t hi, const uint64_t *index, const uint8_t *in, size_t n, uint64_t *out); /* :152-208 */
void     oracle_wide_code_range(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, size_t dict_size, uint32_t *lo,
                                uint32_t *hi);                                                                   /* :539-547 */
uint64_t oracle_dict_scan_16_64(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint16_t *in, size_t n,
                                int64_t *out);                                                                   /* :531-577 */
uint64_t oracle_dict_scan_32_64(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, size_t dict_size,
                                const uint32_t *in, size_t n, int64_t *out);                                     /* :579-622 */                                                          /* :289-336 */
/* shared_libraries/SharedHeaders/include/Allocator.hpp:95-109: v[i] = i mod 256 */
void     oracle_fill_tiled_column(uint8_t *data, size_t n);

/* ---- TPC-H-style pipelines (Join-Benchmarks/lib/TPCH-Queries/src/tpch.cpp) ---------------------- */
/* column tables, byte-compatible with Join-Benchmarks/lib/SharedHeaders/include/TpcHTypes.hpp:50-83 */
typedef struct { uint64_t n; oracle_row_t *l_orderkey; uint64_t *l_shipdate, *l_commitdate, *l_receiptdate;
                 uint8_t *l_shipmode; uint32_t *l_partkey; float *l_quantity; uint8_t *l_shipinstruct;
                 char *l_returnflag; } oracle_lineitem_t;
typedef struct { uint64_t n; oracle_row_t *o_orderkey; uint64_t *o_orderdate; uint32_t *o_custkey; } oracle_orders_t;
typedef struct { uint64_t n; oracle_row_t *c_custkey; uint8_t *c_mktsegment; uint32_t *c_nationkey; } oracle_customer_t;
typedef struct { uint64_t n; oracle_row_t *p_partkey; uint8_t *p_brand; uint32_t *p_size; uint8_t *p_container; } oracle_part_t;
/* row counts, like the reference; filtered[] receives the selection cardinalities (3 entries) */
int64_t oracle_tpch_q12(const oracle_lineitem_t *l, const oracle_orders_t *o, uint64_t *filtered);                 /* :219-253 */
int64_t oracle_tpch_q3(const oracle_customer_t *c, const oracle_orders_t *o, const oracle_lineitem_t *l,
                       uint64_t *filtered, uint64_t *join1_rows);                                                    /* :37-117  */
int64_t oracle_tpch_q19(const oracle_lineitem_t *l, const oracle_part_t *p, uint64_t *filtered, uint64_t *join1_rows); /* :255-309 */

#ifdef __cplusplus
}
#endif
#endif
