/*
 * aqp/b200_tpch.h — TPC-H-style Q3 / Q12 / Q19 filter->join pipelines on the B200.
 *
 * The callers of the join hot path in the reference: Join-Benchmarks/lib/TPCH-Queries/src/tpch.cpp
 * (tpch_q3 :37, tpch_q12 :219, tpch_q19 :255), with the column structs of
 * Join-Benchmarks/lib/SharedHeaders/include/TpcHTypes.hpp:50-88 and the predicates of
 * Q3Predicates.hpp:25-55, Q12Predicates.hpp:22-37, Q19Predicates.hpp:27-78. Like the reference the queries
 * return ROW COUNTS only (tpch.cpp:101,:241,:290-299), no aggregates.
 *
 * Two layers, as for the join itself:
 *   host drop-ins   tpch_q3 / tpch_q12 / tpch_q19 with the reference's signatures (tpch.hpp:7-21): the columns
 *                   they read are copied to HBM (inside the call), the pipeline runs on the device;
 *   device-resident b200_tpch_generate_device / b200_tpch_upload fill the library's device tables once,
 *                   b200_tpch_q*_device run on them — filters, joins and intermediate tables never leave HBM.
 */
#ifndef AQP_B200_TPCH_H
#define AQP_B200_TPCH_H

#include <stdint.h>

#include "data_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef TPCTYPES_HPP   /* the reference's TpcHTypes.hpp was not included: declare its POD tables */

/* dictionary codes written by the reference's loader (TpcHTypes.hpp:8-29, TpcHCommons.cpp:142-183) */
#define B200_L_SHIPMODE_MAIL 1
#define B200_L_SHIPMODE_SHIP 2
#define B200_L_SHIPMODE_AIR 3
#define B200_L_SHIPMODE_AIR_REG 4   /* never produced: the loader compares with "AIR REG", dbgen writes "REG AIR" */
#define B200_L_SHIPINSTRUCT_DELIVER_IN_PERSON 1
#define B200_MKT_BUILDING 1

/* TpcHTypes.hpp:50-61 */
struct LineItemTable {
    uint64_t numTuples;
    tuple_t *l_orderkey;      /* key = orderkey, payload = row id */
    uint64_t *l_shipdate;     /* epoch seconds */
    uint64_t *l_commitdate;
    uint64_t *l_receiptdate;
    uint8_t *l_shipmode;
    type_key *l_partkey;
    float *l_quantity;
    uint8_t *l_shipinstruct;
    char *l_returnflag;
};
/* TpcHTypes.hpp:63-68 */
struct OrdersTable {
    uint64_t numTuples;
    tuple_t *o_orderkey;      /* key = orderkey, payload = row id */
    uint64_t *o_orderdate;
    type_key *o_custkey;
};
/* TpcHTypes.hpp:70-75 */
struct CustomerTable {
    uint64_t numTuples;
    tuple_t *c_custkey;
    uint8_t *c_mktsegment;
    type_key *c_nationkey;
};
/* TpcHTypes.hpp:77-83 */
struct PartTable {
    uint64_t numTuples;
    tuple_t *p_partkey;
    uint8_t *p_brand;
    uint32_t *p_size;
    uint8_t *p_container;
};
#endif /* TPCTYPES_HPP */

/* ---- host drop-ins (tpch.hpp:7-21). `algorithm` must be "RHO". result->totalresults = the query's row
 *      count: Q3/Q12 the final join's matches (tpch.cpp:101,:241); Q19 the matches that survive
 *      q19FinalPredicate (the reference only logs that number, :299, and returns the un-filtered join result;
 *      here result->result is the empty table). ------------------------------------------------------------ */
void tpch_q3(struct result_t *result, const struct CustomerTable *c, const struct OrdersTable *o,
             const struct LineItemTable *l, const char *algorithm, struct joinconfig_t *config);
void tpch_q12(struct result_t *result, const struct LineItemTable *l, const struct OrdersTable *o,
              const char *algorithm, struct joinconfig_t *config);
void tpch_q19(struct result_t *result, const struct LineItemTable *l, const struct PartTable *p,
              const char *algorithm, struct joinconfig_t *config);

/* ---- device-resident tables -------------------------------------------------------------------------------- */
struct b200_tpch_stats_t {
    uint64_t result_rows;       /* the query's answer (see above) */
    uint64_t input_rows;        /* sum of the input table cardinalities (the reference's throughput basis) */
    uint64_t filtered[3];       /* rows surviving selection 1..3 (tpch.cpp's selection_1..3) */
    uint64_t join1_rows;        /* matches of the first join (Q3: customers x orders, Q19: part x lineitem) */
    float ms_total;             /* device time of the whole pipeline, CUDA events */
    float ms_filter;
    float ms_join;
    float ms_other;             /* transforms / post-filter */
    uint32_t kernel_launches;
    uint32_t reserved;
};

/* Synthesize the four tables straight into HBM. Cardinalities follow TPC-H: customer 150 000 x SF, orders
 * 1 500 000 x SF, lineitem 4 per order, part 200 000 x SF. Keys and encodings follow dbgen + the reference's
 * loader: sparse order keys (8 of every 32 values), customer keys not divisible by 3, dates as epoch seconds
 * (o_orderdate uniform in 1992-01-01..1998-08-02, l_shipdate = o_orderdate + 1..121 d, l_commitdate =
 * o_orderdate + 30..90 d, l_receiptdate = l_shipdate + 1..30 d), uniform categorical columns with the
 * reference's dictionary codes (7 ship modes, 4 ship instructions, 5 market segments, 25 brands, 40
 * containers, sizes 1..50, quantities 1..50). Values are counter-based hashes of (seed, row). */
int b200_tpch_generate_device(double scale_factor, uint64_t seed);
/* Copy host tables (any pointer may be NULL = table not needed) into the library's device tables. */
int b200_tpch_upload(const struct LineItemTable *l, const struct OrdersTable *o, const struct CustomerTable *c,
                     const struct PartTable *p);
/* Copy the device tables into freshly malloc'd host columns (release with b200_tpch_free_host). */
int b200_tpch_download(struct LineItemTable *l, struct OrdersTable *o, struct CustomerTable *c, struct PartTable *p);
void b200_tpch_free_host(struct LineItemTable *l, struct OrdersTable *o, struct CustomerTable *c, struct PartTable *p);
void b200_tpch_free_device(void);

/* The reference's binary column format (CSVConvert.cpp:16-190 writes it from dbgen's .tbl files, TpcHCommons.cpp:194-214,
 * :235-295,:423-451,:506-537,:594-623 reads it): <root>/scaleNNN/<table>.tbl.dir/{size, <column>.bin} with <table> in
 * lineitem / orders / customer / part. read: every column file that exists is loaded into 64-byte aligned host memory,
 * missing ones stay NULL (release with b200_tpch_free_host; hand to b200_tpch_upload or the host drop-ins). write: the
 * non-NULL columns of the given host tables. Host code, no GPU involved. */
int b200_tpch_read_binary(const char *root, int scale, struct LineItemTable *l, struct OrdersTable *o, struct CustomerTable *c,
                          struct PartTable *p);
int b200_tpch_write_binary(const char *root, int scale, const struct LineItemTable *l, const struct OrdersTable *o,
                           const struct CustomerTable *c, const struct PartTable *p);

/* Multi-GPU, one process per GPU: every rank generates rows [total * rank / world, total * (rank+1) / world)
 * of every table (values depend on the global row only: the shards together are the single-GPU tables), filters its own
 * line items and joins through the sharded join of b200_mg_* (include/aqp/b200_aqp.h). b200_tpch_mg_init wraps
 * b200_mg_init_caps with capacities derived from the shard sizes; stats->result_rows of b200_tpch_q*_mg is the GLOBAL
 * answer, the other fields describe this rank. Finish with b200_mg_finalize(). */
int b200_tpch_generate_shard_device(double scale_factor, uint64_t seed, uint32_t rank, uint32_t world);
int b200_tpch_mg_init(int rank, int world, const unsigned char *nccl_unique_id /* 128 bytes */);
int b200_tpch_q12_mg(struct b200_tpch_stats_t *stats);
/* Q3: the matches of join 1 stay sharded by customer key and feed join 2 as its build side (b200_mg_join_materialize);
 * Q19: the attributes of the final predicate travel packed in the payloads, the per-rank counts are all-reduced.
 * stats->join1_rows and stats->result_rows are GLOBAL. Both need b200_tpch_mg_init like Q12. (Q3: the matches of join 1
 * a rank ends up owning are its build-side shard of join 2 and must fit the capacity b200_tpch_mg_init derived from the
 * orders shard - they do unless almost all qualifying orders hash to one rank; the call fails with an error otherwise.) */
int b200_tpch_q3_mg(struct b200_tpch_stats_t *stats);
int b200_tpch_q19_mg(struct b200_tpch_stats_t *stats);

int b200_tpch_q3_device(struct b200_tpch_stats_t *stats);
int b200_tpch_q12_device(struct b200_tpch_stats_t *stats);
int b200_tpch_q19_device(struct b200_tpch_stats_t *stats);

#ifdef __cplusplus
}
#endif
#endif /* AQP_B200_TPCH_H */
