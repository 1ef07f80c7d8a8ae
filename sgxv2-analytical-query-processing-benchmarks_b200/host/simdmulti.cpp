// simdmulti.cpp — the scan benchmark driver, GPU edition.
//
// Host-side mirror of Scan-Micro-Benchmarks/microbenchmarks/SimdScanMulti/App/App.cpp:40-230 and
// multithreadedscan.cpp:14-117 for the two modes on the hot path (bitvector, noIndex = implicit row
// ids). Flags follow App/flags.hpp:8-40 in spirit (gflags is not available here):
//   --mode=bitvector|noIndex --num_entries=N --selectivity=PCT --num_runs=K --warmup=W --unique
// Data = 0..255 tiled (Allocator.hpp:95-109), predicate [0, round(sel/100*255)] (types.hpp:125).
// Prints one CSV row like PerfEventBlock does, with GB/s computed as in results/plot.py:22-23.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "aqp/b200_aqp.h"

int main(int argc, char **argv) {
    std::string mode = "bitvector";
    size_t n = 1ull << 28, runs = 10, warmup = 1;
    int sel = 10, unique = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&](const char *k) -> const char * {
            size_t l = strlen(k);
            return a.compare(0, l, k) == 0 && a.size() > l && a[l] == '=' ? a.c_str() + l + 1 : nullptr;
        };
        if (const char *v = val("--mode")) mode = v;
        else if (const char *v = val("--num_entries")) n = strtoull(v, nullptr, 10);
        else if (const char *v = val("--selectivity")) sel = atoi(v);
        else if (const char *v = val("--num_runs")) runs = strtoull(v, nullptr, 10);
        else if (const char *v = val("--warmup")) warmup = strtoull(v, nullptr, 10);
        else if (a == "--unique") unique = 1;
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    if (n % 64) { fprintf(stderr, "num_entries must be a multiple of 64 (flags.hpp:42-46)\n"); return 2; }
    if (b200_init(-1)) { fprintf(stderr, "%s\n", b200_last_error()); return 1; }
    const uint8_t lo = 0, hi = (uint8_t) std::round(sel / 100.0 * 255.0);

    std::vector<uint8_t> data(n);
    for (size_t i = 0; i < n; ++i) data[i] = (uint8_t) (i & 255);
    uint64_t ns = 0;
    size_t count = 0;
    if (mode == "bitvector") {
        std::vector<uint64_t> out(n / 64);
        b200_bitvector_scan_user(lo, hi, data.data(), n, out.data(), &ns, runs, warmup, unique);
        for (uint64_t w : out) count += __builtin_popcountll(w);
    } else if (mode == "noIndex") {
        size_t cap = n / 256 * (hi + 1) + 64;   // pre_alloc_per_thread: count()+64 (ResultAllocators.hpp:17)
        std::vector<uint64_t> out(cap);
        b200_index_scan_user(lo, hi, data.data(), n, out.data(), cap, &count, &ns, runs, warmup, unique);
    } else if (mode == "scalar") {   // ScalarScan.hpp:8-20: the scalar twin, all n values
        size_t cap = n / 256 * (hi + 1) + 320;
        std::vector<uint64_t> out(cap);
        double t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        for (size_t r = 0; r < (unique ? 1 : runs); ++r) count = b200_scalar_index_scan(lo, hi, data.data(), n, out.data(), cap);
        ns = (uint64_t) ((std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0) * 1e9);
    } else {
        fprintf(stderr, "mode must be bitvector, noIndex or scalar\n");
        return 2;
    }
    size_t eff_runs = unique ? 1 : runs;
    double gbs = (double) n * eff_runs / (ns * 1e-9) / 1e9;
    printf("mode,num_entries,selectivity,predicate_low,predicate_high,num_runs,matches,timeMicroSec,GBs,copyMicroSec\n");
    printf("%s,%zu,%d,%u,%u,%zu,%zu,%.1f,%.2f,%.1f\n", mode.c_str(), n, sel, lo, hi, eff_runs, count, ns * 1e-3, gbs,
           b200_scan_last_copy_ns() * 1e-3);
    return 0;
}
