/*
 * TEST INFRASTRUCTURE ONLY — never linked into, or called from, the product path.
 *
 * extern "C" driver around the UNMODIFIED reference query pipelines
 * Join-Benchmarks/lib/TPCH-Queries/src/tpch.cpp (tpch_q3 :37, tpch_q12 :219, tpch_q19 :255), compiled where they
 * lie together with the reference RHO (see oracle/Makefile, target libref_tpch.so, flags -DSIMD -DFULL_QUERY
 * -DCHUNKED_TABLE -DUNROLL -DFORCE_2_PHASES = the paper's configuration). The reference's run_join dispatcher
 * (joins.cpp) does not compile at HEAD, so run_join is provided here and forwards to RHO().
 * Q19's answer ("Total matches") is only logged by the reference (tpch.cpp:299); it is parsed from the
 * captured log.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#include "data-types.h"
#include "TpcHTypes.hpp"
#include "tpch.hpp"
#include "joins.hpp"
#include "ChunkedTable.hpp"
#include "Logger.hpp"

#ifndef REF_TPCH_DROPIN
#include "radix/radix_join.h"
void run_join(result_t *res, const table_t *relR, const table_t *relS, const char *algorithm_name,
              const joinconfig_t *config) {
    if (strcmp(algorithm_name, "RHO") != 0) {
        fprintf(stderr, "ref_tpch_shim: only RHO is wired\n");
        exit(EXIT_FAILURE);
    }
    result_t *tmp = RHO(relR, relS, config);
    memcpy(res, tmp, sizeof(result_t));
    free(tmp);
}
#endif
#if defined(REF_TPCH_DROPIN) && defined(REF_TPCH_BACKTRACE)
#include <execinfo.h>
#include <signal.h>
static void segv_bt(int sig) {
    void *frames[64];
    int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}
__attribute__((constructor)) static void install_bt() { signal(SIGSEGV, segv_bt); }
#endif
/* -DREF_TPCH_DROPIN (target libdropin_tpch.so): no join code of the reference is compiled in; run_join and
 * destroy_table come from libb200aqp.so through shim/b200aqp_cxx_shim.cpp - the link-level drop-in test. */

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* run f with stdout captured into a temp file; returns the value following `key` in the log (or -1) */
template <typename F>
static long long captured(F f, const char *key, double *seconds) {
    cpu_set_t saved;
    pthread_getaffinity_np(pthread_self(), sizeof(saved), &saved);
    char path[] = "/tmp/ref_tpch_XXXXXX";
    int fd = mkstemp(path);
    fflush(stdout);
    int saved_out = dup(1);
    if (fd >= 0) dup2(fd, 1);
    double t0 = now_s();
    f();
    double t1 = now_s();
    fflush(stdout);
    dup2(saved_out, 1);
    close(saved_out);
    pthread_setaffinity_np(pthread_self(), sizeof(saved), &saved);
    if (seconds) *seconds = t1 - t0;
    long long val = -1;
    if (fd >= 0) {
        FILE *fp = fopen(path, "r");
        char line[4096];
        while (fp && fgets(line, sizeof line, fp)) {
            const char *p = key ? strstr(line, key) : nullptr;
            if (p) val = strtoll(p + strlen(key), nullptr, 10);
        }
        if (fp) fclose(fp);
        close(fd);
        unlink(path);
    }
    return val;
}

static void drop_result(result_t &r) {
    if (r.result_type == 1 && r.result) {
        auto *ct = static_cast<chunked_table_t *>(r.result);
        if (r.materialized) destroy_table(ct); else free(ct->chunks);
        free(ct);
    }
}

extern "C" {

long long ref_tpch_q3(const CustomerTable *c, const OrdersTable *o, const LineItemTable *l, int nthreads, double *seconds) {
    result_t r{};
    joinconfig_t cfg{};
    cfg.NTHREADS = nthreads;
    captured([&] { tpch_q3(&r, c, o, l, "RHO", &cfg); }, nullptr, seconds);
    long long m = r.totalresults;
    drop_result(r);
    return m;
}

long long ref_tpch_q12(const LineItemTable *l, const OrdersTable *o, int nthreads, double *seconds) {
    result_t r{};
    joinconfig_t cfg{};
    cfg.NTHREADS = nthreads;
    captured([&] { tpch_q12(&r, l, o, "RHO", &cfg); }, nullptr, seconds);
    long long m = r.totalresults;
    drop_result(r);
    return m;
}

/* returns the post-filter match count; *join_rows receives the un-filtered join result size */
long long ref_tpch_q19(const LineItemTable *l, const PartTable *p, int nthreads, long long *join_rows, double *seconds) {
    result_t r{};
    joinconfig_t cfg{};
    cfg.NTHREADS = nthreads;
    long long m = captured([&] { tpch_q19(&r, l, p, "RHO", &cfg); }, "Total matches = ", seconds);
    if (join_rows) *join_rows = r.totalresults;
    drop_result(r);
    return m;
}

}  /* extern "C" */
