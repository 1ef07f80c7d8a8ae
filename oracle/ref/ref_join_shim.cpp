/*
 * TEST INFRASTRUCTURE ONLY — never linked into, or called from, the product path.
 *
 * extern "C" driver around the UNMODIFIED reference sources compiled where they lie under
 * /root/reference (see oracle/Makefile).  It replaces the reference's own (non-compiling,
 * SURVEY.md §0.5) joins.cpp dispatcher by calling RHO() directly
 * (Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp:1640) and exposes the reference's
 * relation generators (Join-Benchmarks/lib/AppUtilities/src/generator.cpp:75,:352,:474,:638).
 *
 * Used by: tests/ (to pin the C restatement in oracle/), tests/golden/make_golden.py (to
 * generate committed fixtures) and bench.py's `--impl reference` / cpu_baseline leg.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <unistd.h>
#include <pthread.h>
#include <sched.h>

#include "data-types.h"
#include "generator.h"
#include "radix/radix_join.h"
#include "ChunkedTable.hpp"
#include "rdtscpWrapper.h"
#include "Logger.hpp"

static double g_tsc_hz = 0.0;

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* measured TSC rate: the reference hard-codes CYCLES_PER_MICROSECOND (CMakeLists.txt:17) */
static double tsc_hz() {
    if (g_tsc_hz > 0) return g_tsc_hz;
    double t0 = now_s();
    uint64_t c0 = rdtscp_s();
    while (now_s() - t0 < 0.05) {}
    double t1 = now_s();
    uint64_t c1 = rdtscp_s();
    g_tsc_hz = (double) (c1 - c0) / (t1 - t0);
    return g_tsc_hz;
}

extern "C" {

double ref_tsc_hz(void) { return tsc_hz(); }

void ref_seed_generator(unsigned seed) { seed_generator(seed); }

/* The generators malloc() the tuple array themselves (generator.cpp:357); the caller owns it
 * afterwards and releases it with ref_free(). Payload is never written by the reference
 *  so we zero it here to make the bytes deterministic. */
static void zero_payload(table_t *t) {
    for (uint64_t i = 0; i < t->num_tuples; ++i) t->tuples[i].payload = 0;
}

row_t *ref_create_relation_pk(uint64_t n) {
    table_t t{};
    if (create_relation_pk(&t, n, 0) != 0) return nullptr;
    zero_payload(&t);
    return t.tuples;
}

row_t *ref_create_relation_fk(uint64_t n, int64_t maxid) {
    table_t t{};
    if (create_relation_fk(&t, n, maxid, 0) != 0) return nullptr;
    zero_payload(&t);
    return t.tuples;
}

row_t *ref_create_relation_fk_sel(uint64_t n, int64_t maxid) {
    table_t t{};
    if (create_relation_fk_sel(&t, n, maxid, 0) != 0) return nullptr;
    zero_payload(&t);
    return t.tuples;
}

row_t *ref_create_relation_zipf(uint64_t n, int64_t maxid, double z) {
    table_t t{};
    if (create_relation_zipf(&t, n, maxid, z, 0) != 0) return nullptr;
    zero_payload(&t);
    return t.tuples;
}

void ref_free(void *p) { free(p); }

/*
 * Run the reference RHO(). stdout of the reference logger is captured into a temp file so
 * the per-phase cycle counters it prints (radix_join.cpp:265-292) can be returned:
 *   cycles[0] Total Join Time, [1] Partition Overall, [2] Pass One, [3] Pass Two,
 *   [4] Build+Join Overall, [5] Build, [6] Join(probe)
 * If `triples` != NULL and materialize != 0 the chunked result (data-types.h:68-92) is
 * flattened into it (up to `cap` entries). checksum/keysum are computed from the
 * materialised chunks:  checksum = sum(Rpayload + Spayload), keysum = sum(key).
 */
int ref_rho(const row_t *R, uint64_t nR, const row_t *S, uint64_t nS, int nthreads, int materialize,
            int64_t *matches, uint64_t *checksum, uint64_t *keysum,
            output_triple_t *triples, uint64_t cap, uint64_t *cycles, double *wall_seconds) {
    table_t tr{const_cast<row_t *>(R), nR, 0, 0};
    table_t ts{const_cast<row_t *>(S), nS, 0, 0};
    joinconfig_t cfg{};
    cfg.NTHREADS = nthreads;
    cfg.MATERIALIZE = materialize;
    cfg.ALLOC_CORE = 0;

    /* save affinity: join_init_run re-pins the calling thread (radix_join.cpp:1378,:1483) */
    cpu_set_t saved;
    pthread_getaffinity_np(pthread_self(), sizeof(saved), &saved);

    char path[] = "/tmp/ref_rho_XXXXXX";
    int fd = mkstemp(path);
    fflush(stdout);
    int saved_out = dup(1);
    if (fd >= 0) dup2(fd, 1);

    double t0 = now_s();
    result_t *res = RHO(&tr, &ts, &cfg);
    double t1 = now_s();

    fflush(stdout);
    dup2(saved_out, 1);
    close(saved_out);
    pthread_setaffinity_np(pthread_self(), sizeof(saved), &saved);

    if (wall_seconds) *wall_seconds = t1 - t0;
    if (cycles) {
        static const char *keys[7] = {"Total Join Time (cycles)", "Partition Overall (cycles)",
                                      "Partition Pass One (cycles)", "Partition Pass Two (cycles)",
                                      "Build+Join Overall (cycles)", "Build (cycles)", "Join (cycles)"};
        for (int i = 0; i < 7; ++i) cycles[i] = 0;
        if (fd >= 0) {
            FILE *f = fopen(path, "r");
            char line[4096];
            while (f && fgets(line, sizeof line, f)) {
                for (int i = 0; i < 7; ++i) {
                    const char *p = strstr(line, keys[i]);
                    if (p) {
                        const char *c = strchr(p + strlen(keys[i]), ':');
                        if (c) cycles[i] = strtoull(c + 1, nullptr, 10);
                    }
                }
            }
            if (f) fclose(f);
        }
    }
    if (fd >= 0) { close(fd); unlink(path); }
    if (!res) return -1;

    *matches = res->totalresults;
    uint64_t cs = 0, ks = 0, n = 0;
    if (res->result_type == 1 && res->result) {
        auto *ct = static_cast<chunked_table_t *>(res->result);
        if (materialize) {
            for (uint64_t c = 0; c < ct->num_chunks; ++c) {
                const table_chunk_t *ch = ct->chunks[c];
                for (uint64_t i = 0; i < ch->num_tuples; ++i) {
                    const output_triple_t &t = ch->tuples[i];
                    cs += (uint64_t) t.Rpayload + (uint64_t) t.Spayload;
                    ks += t.key;
                    if (triples && n < cap) triples[n] = t;
                    ++n;
                }
            }
            destroy_table(ct);
        } else {
            free(ct->chunks);
        }
        free(ct);
    }
    if (checksum) *checksum = cs;
    if (keysum) *keysum = ks;
    free(res);
    return 0;
}

}  /* extern "C" */
