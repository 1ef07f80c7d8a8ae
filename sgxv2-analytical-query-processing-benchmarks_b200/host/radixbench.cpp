// radixbench.cpp — the histogram / partitioning micro-benchmarks, GPU edition.
//
// Host-side mirror of Scan-Micro-Benchmarks/microbenchmarks/RadixPartitioning (App/Histogram.cpp:20-30 flags,
// Shared/histogram_algorithms.hpp:10-100, Shared/partitioning_algorithms.hpp:12-28, results/0_histogram.sh):
// one histogram pass and one scatter pass over `data_size` 8-byte tuples for every radix width in
// [min_radix_bits, max_radix_bits], through the stage-level C ABI (b200_radix_hist_device,
// b200_exclusive_scan_u32_device, b200_radix_scatter_device). The CPU benchmark's knobs "mode" and "unrolling factor"
// have no GPU counterpart (the kernels keep 8 resp. 16 tuples in flight per thread); what is swept is the fan-out,
// which decides where the histogram lives (shared memory up to 2^15 bins, global REDs above) and how long the
// scatter's per-partition runs are.
//   --data_size=N --min_radix_bits=a --max_radix_bits=b --repeat=K --num_keys_exp=e (keys uniform in [0, 2^e))
// Prints one CSV row per radix width: bits,fanout,hist_ms,hist_GBps,scatter_ms,scatter_GBps (GB/s of algorithmic
// bytes: 8 B/tuple for the histogram, 16 B/tuple for the scatter).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <chrono>

#include "aqp/b200_aqp.h"

#define CK(x)                                                                      \
    do {                                                                           \
        if ((x) != 0) {                                                            \
            fprintf(stderr, "%s failed: %s\n", #x, b200_last_error());             \
            return 1;                                                              \
        }                                                                          \
    } while (0)

int main(int argc, char **argv) {
    size_t n = 1ull << 27;
    int min_bits = 0, max_bits = 15, repeat = 5, keys_exp = 30;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&](const char *k) -> const char * {
            size_t l = strlen(k);
            return a.compare(0, l, k) == 0 && a.size() > l && a[l] == '=' ? a.c_str() + l + 1 : nullptr;
        };
        if (const char *v = val("--data_size")) n = strtoull(v, nullptr, 10);
        else if (const char *v = val("--min_radix_bits")) min_bits = atoi(v);
        else if (const char *v = val("--max_radix_bits")) max_bits = atoi(v);
        else if (const char *v = val("--repeat")) repeat = atoi(v);
        else if (const char *v = val("--num_keys_exp")) keys_exp = atoi(v);
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    if (max_bits > 24 || min_bits < 0 || min_bits > max_bits || keys_exp < 1 || keys_exp > 32) {
        fprintf(stderr, "need 0 <= min_radix_bits <= max_radix_bits <= 24, 1 <= num_keys_exp <= 32\n");
        return 2;
    }
    CK(b200_init(-1));
    row_t *d_in = static_cast<row_t *>(b200_device_alloc(n * sizeof(row_t)));
    row_t *d_out = static_cast<row_t *>(b200_device_alloc(n * sizeof(row_t) + 64));
    const size_t max_fan = (size_t) 1 << max_bits;
    uint32_t *d_hist = static_cast<uint32_t *>(b200_device_alloc((max_fan + 1) * 4));
    uint32_t *d_off = static_cast<uint32_t *>(b200_device_alloc((max_fan + 1) * 4));
    uint32_t *d_cur = static_cast<uint32_t *>(b200_device_alloc((max_fan + 1) * 4));
    if (!d_in || !d_out || !d_hist || !d_off || !d_cur) { fprintf(stderr, "%s\n", b200_last_error()); return 1; }
    // uniform random keys (RNG.cpp of the reference draws them from a xorshift generator): a keyed permutation of
    // 1..2^keys_exp repeated, generated in HBM
    CK(b200_gen_fk_device(d_in, n, (uint64_t) 1 << keys_exp, 0, n, 12345, nullptr));
    CK(b200_device_sync());
    // timed with the host clock between two device synchronisations (the library owns its stream; a launch costs
    // ~10 us, the passes take 0.2-1 ms)
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::vector<uint32_t> zeros(max_fan, 0u);
    printf("bits,fanout,hist_ms,hist_GBps,scatter_ms,scatter_GBps\n");
    for (int bits = min_bits; bits <= max_bits; ++bits) {
        const uint32_t fan = 1u << bits;
        float best_h = 1e30f, best_s = 1e30f;
        for (int r = 0; r < repeat + 1; ++r) {
            CK(b200_memcpy_h2d(d_hist, zeros.data(), (size_t) fan * 4));   // the histogram call accumulates
            CK(b200_device_sync());
            double t0 = now_ms();
            CK(b200_radix_hist_device(d_in, n, 0, (uint32_t) bits, d_hist, nullptr));
            CK(b200_device_sync());
            float ms = (float) (now_ms() - t0);
            if (r && ms < best_h) best_h = ms;
            if (bits >= 1 && bits <= 8) {   // one scatter pass handles up to 2^8 partitions
                CK(b200_exclusive_scan_u32_device(d_hist, fan, d_off, nullptr));
                CK(b200_device_sync());
                t0 = now_ms();
                CK(b200_radix_scatter_device(d_in, n, 0, (uint32_t) bits, d_off, d_cur, d_out, nullptr));
                CK(b200_device_sync());
                ms = (float) (now_ms() - t0);
                if (r && ms < best_s) best_s = ms;
            }
        }
        if (bits >= 1 && bits <= 8)
            printf("%d,%u,%.4f,%.1f,%.4f,%.1f\n", bits, fan, best_h, 8.0 * n / best_h / 1e6, best_s, 16.0 * n / best_s / 1e6);
        else
            printf("%d,%u,%.4f,%.1f,,\n", bits, fan, best_h, 8.0 * n / best_h / 1e6);
    }
    b200_device_free(d_in);
    b200_device_free(d_out);
    b200_device_free(d_hist);
    b200_device_free(d_off);
    b200_device_free(d_cur);
    return 0;
}
