// native_mg.cpp — the join benchmark driver for G GPUs of one box: the multi-GPU edition of host/native.cpp
// (Join-Benchmarks/App/TEEBench/native.cpp:20-147), in C++ against the extern "C" API only.
//
//   native_mg -g <gpus> [-r |R|] [-s |S|] [-z skew] [-x seedR] [-y seedS] [--reps N]
//
// The parent forks one process per GPU (before anything touches CUDA). Rank 0 obtains the NCCL unique id and
// publishes it through an anonymous shared mapping; every rank binds its GPU, generates ITS row range of R and S
// straight into HBM (the generators are defined on the global row index, so the G shards together are exactly the
// relations host/native.cpp -g joins on one GPU) and calls b200_mg_join(). Rank 0 prints the reference's log lines
// ("Total Join Time (cycles)", "Throughput (M rec/sec)", SGXv2Scripts/scripts/helpers/runner.py:19-53; cycles are a
// nominal 1 GHz counter) with the time of the slowest rank.
#include <getopt.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
    fflush(stdout);
}

struct Shared {
    std::atomic<int> id_ready;
    unsigned char id[128];
    float ms[8][64];   // [rank][rep]
    std::atomic<int> failed;
};

static int rank_main(int rank, int world, Shared *sh, uint64_t nR, uint64_t nS, unsigned r_seed, unsigned s_seed, double skew,
                     int reps) {
    auto fail = [&](const char *what) {
        fprintf(stderr, "[rank %d][ERROR] %s: %s\n", rank, what, b200_last_error());
        sh->failed = 1;
        sh->id_ready = 1;   // do not leave the others waiting for the id
        return EXIT_FAILURE;
    };
    if (b200_init(rank)) return fail("b200_init");
    if (rank == 0) {
        if (b200_mg_unique_id(sh->id)) return fail("b200_mg_unique_id");
        sh->id_ready = 1;
    } else {
        while (!sh->id_ready.load()) usleep(1000);
        if (sh->failed) return EXIT_FAILURE;
    }
    if (b200_mg_init(rank, world, sh->id, nR, nS)) return fail("b200_mg_init");
    // this rank's row range
    const uint64_t r0 = nR / world * rank, r1 = rank == world - 1 ? nR : nR / world * (rank + 1);
    const uint64_t s0 = nS / world * rank, s1 = rank == world - 1 ? nS : nS / world * (rank + 1);
    row_t *dR = (row_t *) b200_device_alloc((r1 - r0) * 8 + 64), *dS = (row_t *) b200_device_alloc((s1 - s0) * 8 + 64);
    if (!dR || !dS) return fail("b200_device_alloc");
    if (b200_gen_pk_device(dR, nR, r0, r1 - r0, r_seed, nullptr)) return fail("b200_gen_pk_device");
    if (skew > 0 ? b200_gen_zipf_device(dS, nR, skew, s0, s1 - s0, s_seed, nullptr)
                 : b200_gen_fk_device(dS, nS, nR, s0, s1 - s0, s_seed, nullptr))
        return fail("generate S");
    b200_device_sync();
    b200_mg_result_t res{};
    for (int i = 0; i < reps + 1; ++i) {   // one untimed warm-up
        if (b200_mg_join(dR, r1 - r0, dS, s1 - s0, &res)) return fail("b200_mg_join");
        if (i > 0 && i - 1 < 64) sh->ms[rank][i - 1] = res.ms_total;
    }
    if (rank == 0) {
        info("Running RHO (B200 x%d) with 2 passes and %u radix bits (%u + %u)", world, res.radix_bits, res.bits_pass1,
             res.bits_pass2);
        info("Phases on rank 0 (ms): histogram %.3f, scatter+exchange %.3f, barrier %.3f, pass 2 %.3f, build+probe %.3f, "
             "result all-reduce %.3f",
             res.ms_hist, res.ms_scatter, res.ms_barrier, res.ms_pass2, res.ms_join, res.ms_reduce);
    }
    if (b200_mg_finalize()) return fail("b200_mg_finalize");   // collective: every rank has stored its times
    if (rank == 0) {
        for (int i = 0; i < reps && i < 64; ++i) {
            float ms = 0;
            for (int g = 0; g < world; ++g) ms = sh->ms[g][i] > ms ? sh->ms[g][i] : ms;   // the slowest rank
            info("Total Join Time (cycles)    : %lu", (unsigned long) (ms * 1e6));
            info("Throughput (M rec/sec) : %.2lf", (double) (nR + nS) / (ms * 1e3));
        }
        info("Total input tuples : %lu", (unsigned long) (nR + nS));
        info("Result tuples : %lu", (unsigned long) res.matches);
        info("Matches = %lu", (unsigned long) res.matches);
        info("Checksum = %lu", (unsigned long) res.checksum);
        info("Keysum = %lu", (unsigned long) res.keysum);
    }
    b200_device_free(dR);
    b200_device_free(dS);
    return 0;
}

int main(int argc, char **argv) {
    g_t0 = now_s();
    uint64_t r_size = 1ull << 27, s_size = 1ull << 29;   // BASELINE config 3
    unsigned r_seed = 11111, s_seed = 22222;
    int gpus = 1, reps = 5;
    double skew = 0;
    static option long_opts[] = {{"reps", required_argument, nullptr, 'R'}, {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "a:g:n:r:s:x:y:z:h", long_opts, nullptr)) != -1) {
        switch (c) {
            case 'g': gpus = atoi(optarg); break;
            case 'r': r_size = strtoull(optarg, nullptr, 10); break;
            case 's': s_size = strtoull(optarg, nullptr, 10); break;
            case 'x': r_seed = (unsigned) strtoul(optarg, nullptr, 10); break;
            case 'y': s_seed = (unsigned) strtoul(optarg, nullptr, 10); break;
            case 'z': skew = atof(optarg); break;
            case 'R': reps = atoi(optarg); break;
            case 'h': printf("usage: native_mg -g gpus [-r |R|] [-s |S|] [-z skew] [-x seedR] [-y seedS] [--reps N]\n"); return 0;
            default: break;   // -a RHO, -n threads: accepted
        }
    }
    if (gpus != 1 && gpus != 2 && gpus != 4 && gpus != 8) {
        fprintf(stderr, "[ERROR] -g must be 1, 2, 4 or 8\n");
        return EXIT_FAILURE;
    }
    info("Welcome from native_mg (B200 x%d)!", gpus);
    info("Build relation R on %d device(s) with size = %.2lf MB (%lu tuples)", gpus, 8.0 * r_size / pow(2, 20), (unsigned long) r_size);
    info("Build relation S on %d device(s) with size = %.2lf MB (%lu tuples)", gpus, 8.0 * s_size / pow(2, 20), (unsigned long) s_size);
    Shared *sh = static_cast<Shared *>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (sh == MAP_FAILED) {
        perror("mmap");
        return EXIT_FAILURE;
    }
    // anonymous shared mappings start zeroed
    fflush(stdout);
    pid_t pid[8];
    for (int g = 0; g < gpus; ++g) {
        pid[g] = fork();   // before any CUDA call: every process creates its own context
        if (pid[g] == 0) _exit(rank_main(g, gpus, sh, r_size, s_size, r_seed, s_seed, skew, reps));
        if (pid[g] < 0) {
            perror("fork");
            return EXIT_FAILURE;
        }
    }
    int rc = 0;
    for (int g = 0; g < gpus; ++g) {
        int st = 0;
        waitpid(pid[g], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = EXIT_FAILURE;
    }
    return rc;
}
