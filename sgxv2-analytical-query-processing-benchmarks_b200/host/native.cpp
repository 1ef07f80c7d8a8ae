// native.cpp — the join benchmark driver, GPU edition.
//
// Host-side mirror of Join-Benchmarks/App/TEEBench/native.cpp:20-147 with the command line of
// Join-Benchmarks/lib/AppUtilities/src/commons.cpp:37 (-a -r -s -n -m -z -l -x -y -d; thread
// pinning / SSB-mitigation flags are accepted and ignored). It generates R (seed 11111) and S
// (seed 22222) with the reference-identical host generators, calls run_join() from libb200aqp.so
// and prints the log lines SGXv2Scripts/scripts/helpers/runner.py scrapes ("Throughput (M rec/sec)",
// "Total Join Time (cycles)", ...; cycles are a nominal 1 GHz counter, i.e. CPMS = 1000).
// Extra flag: -g generates the relations on the device instead (b200_gen_*_device) and joins them
// through the preload-style device path, so very large inputs do not pay the sequential glibc shuffle.
#include <getopt.h>
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>

#include "aqp/b200_aqp.h"

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static double g_t0;
static void info(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    printf("\x1b[32m[%8.4f][ INFO] %s\x1b[0m\n", now_s() - g_t0, buf);   // Logger.cpp:71-75 (green INFO, reset)
}

int main(int argc, char **argv) {
    g_t0 = now_s();
    info("Welcome from native (B200)!");
    uint64_t r_size = 2097152, s_size = 2097152;   // native.cpp:33-34 defaults
    unsigned r_seed = 11111, s_seed = 22222;
    int nthreads = 2, selectivity = 100, materialize = 0, on_device = 0, reps = 1;
    double skew = 0;
    char alg[128] = "RHO";
    static option long_opts[] = {{"sort-r", no_argument, nullptr, 1}, {"sort-s", no_argument, nullptr, 1},
                                 {"mitigation", no_argument, nullptr, 1}, {"reps", required_argument, nullptr, 'R'},
                                 {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "a:c:d:e:l:n:mr:s:t:u:x:y:z:hvg", long_opts, nullptr)) != -1) {
        switch (c) {
            case 'a': strncpy(alg, optarg, sizeof alg - 1); break;
            case 'd':
                if (!strcmp(optarg, "cache-fit")) { r_size = 10u * 1024 * 1024 / 8; s_size = 40u * 1024 * 1024 / 8; }
                else if (!strcmp(optarg, "cache-exceed")) { r_size = 100u * 1024 * 1024 / 8; s_size = 400u * 1024 * 1024 / 8; }
                else if (!strcmp(optarg, "L")) { r_size = 50000000; s_size = 200000000; }
                else { fprintf(stderr, "Unrecognized dataset: %s\n", optarg); return EXIT_FAILURE; }
                break;
            case 'l': selectivity = atoi(optarg); break;
            case 'm': materialize = 1; break;
            case 'n': nthreads = atoi(optarg); break;
            case 'r': r_size = strtoull(optarg, nullptr, 10); break;
            case 's': s_size = strtoull(optarg, nullptr, 10); break;
            case 'x': r_seed = (unsigned) strtoul(optarg, nullptr, 10); break;
            case 'y': s_seed = (unsigned) strtoul(optarg, nullptr, 10); break;
            case 'z': skew = atof(optarg); break;
            case 'g': on_device = 1; break;
            case 'R': reps = atoi(optarg); break;
            case 'h': printf("usage: native [-a RHO] [-r |R|] [-s |S|] [-n threads] [-m] [-z skew] [-l sel] [-x seedR] [-y seedS] [-g] [--reps N]\n"); return 0;
            default: break;   // -c -e -t -u -v, --sort-*, --mitigation: accepted, no GPU meaning
        }
    }
    if (b200_init(-1) != 0) {
        fprintf(stderr, "[ERROR] %s\n", b200_last_error());
        return EXIT_FAILURE;
    }
    b200_set_verbose(1);
    joinconfig_t cfg{};
    cfg.NTHREADS = nthreads;
    cfg.MATERIALIZE = materialize;
    result_t res{};

    if (on_device) {
        info("Build relation R on device with size = %.2lf MB (%lu tuples)", 8.0 * r_size / pow(2, 20), (unsigned long) r_size);
        row_t *dR = (row_t *) b200_device_alloc(r_size * 8), *dS = (row_t *) b200_device_alloc(s_size * 8);
        if (!dR || !dS || b200_gen_pk_device(dR, r_size, 0, r_size, r_seed, nullptr) ||
            b200_gen_fk_device(dS, s_size, r_size, 0, s_size, s_seed, nullptr)) {
            fprintf(stderr, "[ERROR] %s\n", b200_last_error());
            return EXIT_FAILURE;
        }
        b200_join_stats_t st{};
        for (int i = 0; i < reps; ++i) {
            if (b200_join_device(dR, r_size, dS, s_size, nullptr, 0, &st, nullptr)) {
                fprintf(stderr, "[ERROR] %s\n", b200_last_error());
                return EXIT_FAILURE;
            }
            info("Total Join Time (cycles)    : %lu", (unsigned long) (st.ms_total * 1e6));
            info("Throughput (M rec/sec) : %.2lf", (double) (r_size + s_size) / (st.ms_total * 1e3));
        }
        info("Matches = %lu", (unsigned long) st.matches);
        info("Checksum = %lu", (unsigned long) st.checksum);
        b200_device_free(dR);
        b200_device_free(dS);
        return 0;
    }

    table_t R{}, S{};
    seed_generator(r_seed);
    info("Build relation R with size = %.2lf MB (%lu tuples)", 8.0 * r_size / pow(2, 20), (unsigned long) r_size);
    if (create_relation_pk(&R, r_size, 0)) return EXIT_FAILURE;
    seed_generator(s_seed);
    info("Build relation S with size = %.2lf MB (%lu tuples)", 8.0 * s_size / pow(2, 20), (unsigned long) s_size);
    int rc;
    if (skew > 0) {
        rc = create_relation_zipf(&S, s_size, (int64_t) r_size, skew, 0);             // native.cpp:90-91
    } else if (selectivity != 100) {
        info("Table S selectivity = %d", selectivity);
        uint32_t maxid = selectivity != 0 ? (uint32_t) (100 * r_size / selectivity) : 0;   // native.cpp:93-97
        rc = create_relation_fk_sel(&S, s_size, maxid, 0);
    } else {
        rc = create_relation_fk(&S, s_size, (int64_t) r_size, 0);                     // native.cpp:100
    }
    if (rc) return EXIT_FAILURE;
    for (uint64_t i = 0; i < r_size; ++i) R.tuples[i].payload = (uint32_t) i;   // payload = row id (TpcHCommons.cpp:332)
    for (uint64_t i = 0; i < s_size; ++i) S.tuples[i].payload = (uint32_t) i;

    info("Running algorithm %s", alg);
    for (int i = 0; i < reps; ++i) {
        double t = now_s();
        run_join(&res, &R, &S, alg, &cfg);
        double dt = now_s() - t;
        info("Total join runtime: %.4fs", dt);
        info("throughput = %.2lf [M rec / s]", (double) (r_size + s_size) / dt / 1e6);
        info("Matches = %lu", (unsigned long) res.totalresults);
        if (res.result) {
            destroy_table((chunked_table_t *) res.result);
            free(res.result);
        }
    }
    delete_relation(&R);
    delete_relation(&S);
    return 0;
}
