/*
 * aqp/data_types.h — the drop-in ABI of the join operator.
 *
 * Byte-compatible restatement of the reference's POD types so that callers compiled against
 * Join-Benchmarks/lib/SharedHeaders/include/data-types.h can link against libb200aqp.so unchanged.
 * Field order, widths and names follow data-types.h (cited per struct); layout is asserted at the
 * bottom of this file and again, against the reference header itself, by tests/test_abi.py.
 *
 * If the reference's own data-types.h has already been included (DATA_TYPES_H defined) this header
 * declares nothing, so both can be used in one translation unit.
 */
#ifndef AQP_DATA_TYPES_H
#define AQP_DATA_TYPES_H

#include <stdint.h>

#ifndef DATA_TYPES_H

typedef uint32_t type_key;    /* data-types.h:31 */
typedef uint32_t type_value;  /* data-types.h:32 */

typedef struct row_t tuple_t;
typedef struct table_t relation_t;
typedef struct result_t result_t;
typedef struct joinconfig_t joinconfig_t;

/* data-types.h:44-47 — 8-byte AoS tuple */
struct row_t {
    type_key key;
    type_value payload;
};

/* data-types.h:49-54 — relation handle; tuples are borrowed, never written by the join */
struct table_t {
    struct row_t *tuples;
    uint64_t num_tuples;
    int ratio_holes;
    int sorted;
};

/* data-types.h:68-72 — one materialised match */
struct output_triple_t {
    type_key key;
    type_value Rpayload;
    type_value Spayload;
};

#ifndef CSKB
#define CSKB 16               /* data-types.h:74-76 */
#endif
#define CHUNK_SIZE (1024 * CSKB)
#define TUPLES_PER_CHUNK ((CHUNK_SIZE - 8) / sizeof(struct output_triple_t))   /* = 1364 */

/* data-types.h:81-84 */
struct table_chunk_t {
    uint64_t num_tuples;
    struct output_triple_t tuples[TUPLES_PER_CHUNK];
};

/* data-types.h:86-92 */
struct chunked_table_t {
    struct table_chunk_t **chunks;
    uint64_t current_chunk;
    uint64_t num_chunks;
    uint64_t chunk_capacity;
    uint64_t num_tuples;
};

/* data-types.h:107-114. RHO sets result_type = 1 and result -> chunked_table_t
 * (radix_join.cpp:1466,:1556). There is no checksum field: see
 * b200_last_join_stats() in b200_aqp.h. */
struct result_t {
    int64_t totalresults;
    int nthreads;
    double throughput;
    int materialized;
    void *result;
    int result_type; /* 0 = threadresult_t*, 1 = chunked_table_t* */
};

/* data-types.h:159 */
enum numa_strategy_t { RANDOM, RING, NEXT };

/* data-types.h:162-176. The GPU path reads MATERIALIZE only; NTHREADS is echoed into
 * result_t.nthreads; ALLOC_CORE (thread pinning) has no GPU meaning and is ignored. */
struct joinconfig_t {
    int NTHREADS;
    int PARTFANOUT;
    int SCALARSORT;
    int SCALARMERGE;
    int MWAYMERGEBUFFERSIZE;
    enum numa_strategy_t NUMASTRATEGY;
    int RADIXBITS;
    int WRITETOFILE;
    int MATERIALIZE;
    int PRINT;
    int CRACKING_THRESHOLD;
    int ALLOC_CORE;
};

#endif /* DATA_TYPES_H */

#if defined(__cplusplus)
static_assert(sizeof(struct row_t) == 8, "row_t must be 8 bytes");
static_assert(sizeof(struct table_t) == 24, "table_t layout");
static_assert(sizeof(struct output_triple_t) == 12, "output_triple_t layout");
static_assert(sizeof(struct table_chunk_t) == 8 + 12 * TUPLES_PER_CHUNK, "table_chunk_t layout");
static_assert(sizeof(struct chunked_table_t) == 40, "chunked_table_t layout");
static_assert(sizeof(struct result_t) == 48, "result_t layout");
static_assert(sizeof(struct joinconfig_t) == 48, "joinconfig_t layout");
#endif

#endif /* AQP_DATA_TYPES_H */
