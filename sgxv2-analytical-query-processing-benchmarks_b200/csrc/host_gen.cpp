// host_gen.cpp — host relation generators with the reference's entry points.
//
// Mirrors Join-Benchmarks/lib/AppUtilities/include/generator.h:26-107 /
// src/generator.cpp (seed_generator :75-80, create_relation_pk :352-377, create_relation_fk
// :474-512, create_relation_fk_sel :515-553, create_relation_zipf :638-660, delete_relation
// :663-668). The uniform generators draw from libc srand()/rand() in exactly the reference's
// order, so for the same seed the key sequence is bit-identical (tests/test_abi.py::test_host_generators_reproduce_reference checks
// this against the compiled reference and the committed fixtures). Differences, on purpose:
//   * payload is zeroed (the reference leaves malloc garbage);
//   * `sorted` != 0 is rejected (sorting helpers are out of scope of the hot path);
//   * Zipf uses the seed_generator() value to seed std::mt19937_64 (the reference seeds from
//     std::random_device and is not repeatable, genzipf.cpp:44-45,:104-105).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <random>
#include <algorithm>
#include <vector>

#include "aqp/b200_aqp.h"

namespace aqp {   // csrc/hostcopy.cpp
void *host_alloc_prefer_pinned(size_t bytes);
void host_free_any(void *p);
}

static int g_seeded = 0;
static unsigned int g_seed_value = 0;

extern "C" void seed_generator(unsigned int seed) {
    srand(seed);
    g_seed_value = seed;
    g_seeded = 1;
}

static void check_seed() {
    if (!g_seeded) {
        g_seed_value = (unsigned int) time(nullptr);
        srand(g_seed_value);
        g_seeded = 1;
    }
}

// generator.cpp:19
static inline double rand_range(double n) { return (double) rand() / ((double) RAND_MAX + 1) * n; }

static void shuffle_keys(row_t *t, uint64_t n) {   // generator.cpp:99-109
    if (n == 0) return;
    for (uint64_t i = n - 1; i > 0; i--) {
        int64_t j = (int64_t) rand_range((double) i);
        type_key tmp = t[i].key;
        t[i].key = t[j].key;
        t[j].key = tmp;
    }
}

static void unique_gen(row_t *t, uint64_t n) {   // generator.cpp:143-153
    for (uint64_t i = 0; i < n; i++) {
        t[i].key = (type_key) (i + 1);
        t[i].payload = 0;
    }
    shuffle_keys(t, n);
}

static void unique_gen_maxid(row_t *t, uint64_t n, uint32_t maxid) {   // generator.cpp:156-169
    double jump = (double) (maxid / n);
    double id = maxid == 0 ? 0 : 1;
    for (uint32_t i = 0; i < n; i++) {
        t[i].key = (uint32_t) id;
        t[i].payload = 0;
        id += jump;
    }
    shuffle_keys(t, n);
}

static int alloc_relation(table_t *rel, uint64_t n, int sorted) {
    if (sorted) {
        fprintf(stderr, "b200aqp: sorted relations are not supported by this generator\n");
        return -1;
    }
    rel->num_tuples = n;
    // pinned when a device is present: run_join() then copies the relation with one DMA at the PCIe rate
    rel->tuples = (row_t *) aqp::host_alloc_prefer_pinned((n ? n : 1) * sizeof(row_t));
    rel->sorted = 0;
    rel->ratio_holes = 0;
    if (!rel->tuples) {
        perror("out of memory");
        return -1;
    }
    return 0;
}

extern "C" int create_relation_pk(table_t *rel, uint64_t n, int sorted) {
    check_seed();
    if (alloc_relation(rel, n, sorted)) return -1;
    unique_gen(rel->tuples, n);
    return 0;
}

extern "C" int create_relation_fk(table_t *rel, uint64_t n, const int64_t maxid, int sorted) {
    check_seed();
    if (maxid <= 0) return -1;
    if (alloc_relation(rel, n, sorted)) return -1;
    uint64_t iters = n / (uint64_t) maxid;
    for (uint64_t i = 0; i < iters; i++) unique_gen(rel->tuples + (uint64_t) maxid * i, (uint64_t) maxid);
    uint64_t rem = n % (uint64_t) maxid;
    if (rem > 0) unique_gen(rel->tuples + (uint64_t) maxid * iters, rem);
    return 0;
}

extern "C" int create_relation_fk_sel(table_t *rel, uint64_t n, const int64_t maxid, int sorted) {
    check_seed();
    if (alloc_relation(rel, n, sorted)) return -1;
    uint64_t iters = maxid != 0 ? n / (uint64_t) maxid : 0;
    for (uint64_t i = 0; i < iters; i++)
        unique_gen_maxid(rel->tuples + (uint64_t) maxid * i, (uint64_t) maxid, (uint32_t) maxid);
    uint64_t rem = maxid != 0 ? n % (uint64_t) maxid : n;
    if (rem > 0) unique_gen_maxid(rel->tuples + (uint64_t) maxid * iters, rem, (uint32_t) maxid);
    return 0;
}

extern "C" int create_relation_zipf(table_t *rel, uint64_t n, const int64_t maxid, const double z, int sorted) {
    check_seed();
    if (maxid <= 0) return -1;
    if (alloc_relation(rel, n, sorted)) return -1;
    const uint32_t asz = (uint32_t) maxid;
    // genzipf.cpp:33-51: alphabet = random permutation of 1..size
    std::vector<uint32_t> alphabet(asz);
    for (uint32_t i = 0; i < asz; i++) alphabet[i] = i + 1;
    std::mt19937_64 gen_a{(uint64_t) g_seed_value * 2654435761ull + 1};
    std::shuffle(alphabet.begin(), alphabet.end(), gen_a);
    // genzipf.cpp:58-83: cumulative distribution lookup table
    std::vector<double> lut(asz);
    double scaling = 0.0;
    for (uint32_t i = 1; i <= asz; i++) scaling += 1.0 / pow((double) i, z);
    double sum = 0.0;
    for (uint32_t i = 1; i <= asz; i++) {
        sum += 1.0 / pow((double) i, z);
        lut[i - 1] = sum / scaling;
    }
    // genzipf.cpp:104-137
    std::mt19937_64 gen{(uint64_t) g_seed_value * 40503ull + 7};
    std::uniform_real_distribution<double> dist{0.0, 1.0};
    for (uint64_t i = 0; i < n; i++) {
        const double r = dist(gen);
        uint32_t left = 0, right = asz - 1, pos;
        if (lut[0] >= r) {
            pos = 0;
        } else {
            while (right - left > 1) {
                uint32_t m = (left + right) / 2;
                if (lut[m] < r) left = m; else right = m;
            }
            pos = right;
        }
        rel->tuples[i].key = alphabet[pos];
        rel->tuples[i].payload = 0;
    }
    return 0;
}

extern "C" void delete_relation(table_t *rel) {
    aqp::host_free_any(rel->tuples);
    rel->tuples = nullptr;
}
