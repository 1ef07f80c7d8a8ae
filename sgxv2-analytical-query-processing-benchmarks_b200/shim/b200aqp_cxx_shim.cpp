// b200aqp_cxx_shim.cpp — C++-linkage faces of libb200aqp.so's C symbols, for UNMODIFIED reference callers.
//
// The reference is C++: its callers reach the join library through C++-mangled functions
//   run_join        Join-Benchmarks/lib/Joins/include/joins.hpp:4-6   (App/TEEBench/native.cpp:137, the seven call sites in
//                   lib/TPCH-Queries/src/tpch.cpp :68,:101,:141,:167,:202,:241,:282)
//   destroy_table   Join-Benchmarks/lib/Joins/include/ChunkedTable.hpp:20  (tpch.cpp:82 and result_transformers)
//   seed_generator / create_relation_* / delete_relation   lib/AppUtilities/include/generator.h:26-107
// libb200aqp.so exports the same names with C linkage (one ABI for C, C++, ctypes, cgo ...). A translation unit cannot
// declare both linkages of one name, so the C symbols are bound here under private names with assembler labels and
// the mangled functions forward to them. Compile this file into the reference's build in place of its `joins` and
// `app-utilities` libraries; nothing in the reference's sources changes.
// tests/test_gpu_dropin.py builds the reference's TPC-H pipelines exactly that way and runs them on the GPU.
#include <cstdint>

struct result_t;
struct table_t;
struct joinconfig_t;
struct chunked_table_t;

extern "C" {
void b200aqp_c_run_join(result_t *, const table_t *, const table_t *, const char *, const joinconfig_t *) __asm__("run_join");
void b200aqp_c_destroy_table(chunked_table_t *) __asm__("destroy_table");
void b200aqp_c_seed_generator(unsigned int) __asm__("seed_generator");
int b200aqp_c_create_relation_pk(table_t *, uint64_t, int) __asm__("create_relation_pk");
int b200aqp_c_create_relation_fk(table_t *, uint64_t, const int64_t, int) __asm__("create_relation_fk");
int b200aqp_c_create_relation_fk_sel(table_t *, uint64_t, const int64_t, int) __asm__("create_relation_fk_sel");
int b200aqp_c_create_relation_zipf(table_t *, uint64_t, const int64_t, const double, int) __asm__("create_relation_zipf");
void b200aqp_c_delete_relation(table_t *) __asm__("delete_relation");
}

void run_join(result_t *res, const table_t *relR, const table_t *relS, const char *algorithm_name, const joinconfig_t *config) {
    b200aqp_c_run_join(res, relR, relS, algorithm_name, config);
}
void destroy_table(chunked_table_t *table) { b200aqp_c_destroy_table(table); }
void seed_generator(unsigned int seed) { b200aqp_c_seed_generator(seed); }
int create_relation_pk(table_t *reln, uint64_t ntuples, int sorted) { return b200aqp_c_create_relation_pk(reln, ntuples, sorted); }
int create_relation_fk(table_t *reln, uint64_t ntuples, const int64_t maxid, int sorted) {
    return b200aqp_c_create_relation_fk(reln, ntuples, maxid, sorted);
}
int create_relation_fk_sel(table_t *reln, uint64_t ntuples, const int64_t maxid, int sorted) {
    return b200aqp_c_create_relation_fk_sel(reln, ntuples, maxid, sorted);
}
int create_relation_zipf(table_t *reln, uint64_t ntuples, const int64_t maxid, const double zipfparam, int sorted) {
    return b200aqp_c_create_relation_zipf(reln, ntuples, maxid, zipfparam, sorted);
}
void delete_relation(table_t *reln) { b200aqp_c_delete_relation(reln); }
