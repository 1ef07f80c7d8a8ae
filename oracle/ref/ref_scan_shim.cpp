/*
 * TEST INFRASTRUCTURE ONLY — never linked into, or called from, the product path.
 *
 * extern "C" driver around the UNMODIFIED reference scan kernels
 * Scan-Micro-Benchmarks/shared_libraries/SimdScan/src/SIMD512.cpp compiled where it lies
 * (see oracle/Makefile).  Mirrors the ECALL-shaped entry points of
 * Scan-Micro-Benchmarks/microbenchmarks/SimdScanMulti/Enclave/Enclave.cpp:100-133,:270-299
 * with a std::thread row-range fan-out like scan_wrapper (App/multithreadedscan.cpp:227-258).
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <thread>
#include <vector>

#include "SIMD512.hpp"

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

extern "C" {

uint64_t ref_scan_count(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n) {
    return SIMD512::count(lo, hi, reinterpret_cast<const __m512i *>(data), n);
}

void ref_bitvector_scan(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint64_t *out) {
    SIMD512::bitvector_scan(lo, hi, reinterpret_cast<const __m512i *>(data), n,
                            reinterpret_cast<__mmask64 *>(out));
}

/* out must hold count+64 entries (shared/ResultAllocators.hpp:17). Returns nothing, like the
 * reference; the caller learns the count from ref_scan_count. */
void ref_index_scan(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint64_t *out) {
    SIMD512::implicit_index_scan(lo, hi, reinterpret_cast<const __m512i *>(data), n,
                                 reinterpret_cast<size_t *>(out));
}

/* self_alloc variant as used by run_index_scan (multithreadedscan.cpp:97-109): returns count. */
uint64_t ref_index_scan_self_alloc(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint64_t *out,
                                   uint64_t cap) {
    CacheAlignedVector<size_t> v;
    SIMD512::implicit_index_scan_self_alloc(lo, hi, reinterpret_cast<const __m512i *>(data), n, v, true);
    uint64_t c = v.size();
    if (out) memcpy(out, v.data(), sizeof(uint64_t) * (c < cap ? c : cap));
    return c;
}

uint64_t ref_scan_sum(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n) {
    return SIMD512::sum(lo, hi, reinterpret_cast<const __m512i *>(data), n);
}

/* out must hold count + 16 entries: the last compress-store of a block may be followed by nothing, but the
 * reference gives no slack contract, so the caller is generous. Returns the count. */
uint64_t ref_value_scan(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint32_t *out) {
    return SIMD512::scan(lo, hi, reinterpret_cast<const __m512i *>(data), n, out);
}

/* which: 0 = dict_scan_8bit_64bit, 1 = ..._scalar_gather_scatter, 2 = ..._scalar_unroll, 3 = ..._opt_write
 * (all with cut = true). Returns the count; copies min(count, cap) values to out. */
uint64_t ref_dict_scan_8_64(int which, int64_t lo, int64_t hi, const int64_t *dict, const uint8_t *data, size_t n,
                            int64_t *out, uint64_t cap) {
    CacheAlignedVector<int64_t> v;
    const __m512i *in = reinterpret_cast<const __m512i *>(data);
    switch (which) {
    case 0: SIMD512::dict_scan_8bit_64bit(lo, hi, dict, in, n, v, true); break;
    case 1: SIMD512::dict_scan_8bit_64bit_scalar_gather_scatter(lo, hi, dict, in, n, v, true); break;
    case 2: SIMD512::dict_scan_8bit_64bit_scalar_unroll(lo, hi, dict, in, n, v, true); break;
    default: SIMD512::dict_scan_8bit_64bit_opt_write(lo, hi, dict, in, n, v, true); break;
    }
    uint64_t c = v.size();
    if (out) memcpy(out, v.data(), sizeof(int64_t) * (c < cap ? c : cap));
    return c;
}

uint64_t ref_explicit_index_scan(uint8_t lo, uint8_t hi, const uint64_t *index, const uint8_t *data, size_t n, uint64_t *out) {
    /* the function returns nothing: the count is SIMD512::count's (same predicate, same blocks) */
    SIMD512::explicit_index_scan(lo, hi, reinterpret_cast<const __m512i *>(index), reinterpret_cast<const __m512i *>(data), n,
                                 reinterpret_cast<size_t *>(out));
    return SIMD512::count(lo, hi, reinterpret_cast<const __m512i *>(data), n);
}

uint64_t ref_dict_scan_16_64(int64_t lo, int64_t hi, const int64_t *dict, const uint16_t *data, size_t n, int64_t *out,
                             uint64_t cap) {
    CacheAlignedVector<int64_t> v;
    SIMD512::dict_scan_16bit_64bit(lo, hi, dict, reinterpret_cast<const __m512i *>(data), n, v);
    uint64_t c = v.size();
    if (out) memcpy(out, v.data(), sizeof(int64_t) * (c < cap ? c : cap));
    return c;
}

uint64_t ref_dict_scan_32_64(int64_t lo, int64_t hi, const int64_t *dict, size_t dict_size, const uint32_t *data, size_t n,
                             int64_t *out, uint64_t cap) {
    CacheAlignedVector<int64_t> v;
    SIMD512::dict_scan_32bit_64bit(lo, hi, dict, dict_size, reinterpret_cast<const __m512i *>(data), n, v);
    uint64_t c = v.size();
    if (out) memcpy(out, v.data(), sizeof(int64_t) * (c < cap ? c : cap));
    return c;
}

/*
 * Multi-threaded timed runs, one row range per thread (multithreadedscan.cpp:231-235):
 * mode 0 = bitvector, 1 = row-id list (pre-allocated per thread with count()+64 as
 * pre_alloc_per_thread does, ResultAllocators.hpp:8-19). Returns seconds for `runs` passes.
 */
double ref_scan_mt(int mode, uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, int nthreads,
                   int warmup, int runs, uint64_t *bitvector_out) {
    size_t per = (n / nthreads) / 64 * 64;
    std::vector<CacheAlignedVector<size_t>> idx(nthreads);
    if (mode == 1) {
        for (int t = 0; t < nthreads; ++t) {
            size_t c = SIMD512::count(lo, hi, reinterpret_cast<const __m512i *>(data + t * per), per);
            idx[t].resize(c + 64);
        }
    }
    auto body = [&](int t) {
        const __m512i *in = reinterpret_cast<const __m512i *>(data + t * per);
        if (mode == 0)
            SIMD512::bitvector_scan(lo, hi, in, per, reinterpret_cast<__mmask64 *>(bitvector_out + t * per / 64));
        else
            SIMD512::implicit_index_scan_self_alloc(lo, hi, in, per, idx[t], false);
    };
    auto pass = [&](int reps) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t] { for (int r = 0; r < reps; ++r) body(t); });
        for (auto &x : th) x.join();
    };
    if (warmup > 0) pass(warmup);
    double t0 = now_s();
    pass(runs);
    return now_s() - t0;
}

}  /* extern "C" */
