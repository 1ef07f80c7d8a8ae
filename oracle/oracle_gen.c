/*
 * oracle_gen.c — restatement of the reference relation generators.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h).  Reference: Join-Benchmarks/lib/AppUtilities/src/generator.cpp and genzipf.cpp.
 *
 * The reference draws from libc srand()/rand() (generator.cpp:19,:75-80).  On glibc that is the
 * TYPE_3 additive-feedback generator of random_r.c: r[i] = r[i-3] + r[i-31] (mod 2^32), output
 * r[i] >> 1, state seeded by the Park-Miller LCG 16807 and warmed up by 310 discarded draws.
 * glibc is not vendored in /root/reference, so that published algorithm is restated here and
 * pinned against libc rand() itself and against the compiled reference generator
 * (tests/test_oracle_gen.py: seed 11111 -> 533741 233869 796176 204571 at |R|=2^20).
 */
#include "oracle.h"
#include <math.h>

#define ORACLE_RAND_MAX 2147483647

static int32_t g_r[34];
static int g_f, g_b; /* front / rear indices into the 31-word ring */

void oracle_srand(uint32_t seed) {
    int32_t st[34];
    if (seed == 0) seed = 1;
    st[0] = (int32_t) seed;
    for (int i = 1; i < 31; ++i) {
        /* 16807 * st[i-1] % 2147483647 without overflow (Schrage), as random_r.c does */
        long hi = st[i - 1] / 127773;
        long lo = st[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        st[i] = (int32_t) word;
    }
    for (int i = 0; i < 31; ++i) g_r[i] = st[i];
    g_f = 3;  /* fptr = &state[rand_sep] */
    g_b = 0;  /* rptr = &state[0] */
    for (int i = 0; i < 310; ++i) (void) oracle_rand();
}

int32_t oracle_rand(void) {
    uint32_t v = (uint32_t) g_r[g_f] + (uint32_t) g_r[g_b];
    g_r[g_f] = (int32_t) v;
    int32_t result = (int32_t) (v >> 1);
    if (++g_f >= 31) g_f = 0;
    if (++g_b >= 31) g_b = 0;
    return result;
}

/* generator.cpp:19 RAND_RANGE(N) = (double)rand() / ((double)RAND_MAX + 1) * N */
static double rand_range(double n) {
    return (double) oracle_rand() / ((double) ORACLE_RAND_MAX + 1) * n;
}

/* generator.cpp:99-109 */
static void knuth_shuffle(oracle_row_t *t, uint64_t n) {
    if (n == 0) return;
    for (uint64_t i = n - 1; i > 0; i--) {
        int64_t j = (int64_t) rand_range((double) i);
        uint32_t tmp = t[i].key;
        t[i].key = t[j].key;
        t[j].key = tmp;
    }
}

void oracle_gen_pk(oracle_row_t *rel, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) { rel[i].key = (uint32_t) (i + 1); rel[i].payload = 0; }
    knuth_shuffle(rel, n);
}

void oracle_gen_fk(oracle_row_t *rel, uint64_t n, int64_t maxid) {
    uint64_t iters = n / (uint64_t) maxid;
    for (uint64_t i = 0; i < iters; ++i) oracle_gen_pk(rel + (uint64_t) maxid * i, (uint64_t) maxid);
    uint64_t rem = n % (uint64_t) maxid;
    if (rem > 0) oracle_gen_pk(rel + (uint64_t) maxid * iters, rem);
}

/* generator.cpp:156-169 */
static void unique_gen_maxid(oracle_row_t *rel, uint64_t n, uint32_t maxid) {
    double jump = (double) (maxid / n);   /* integer division first, exactly as the reference */
    double id = maxid == 0 ? 0 : 1;
    for (uint32_t i = 0; i < n; ++i) {
        rel[i].key = (uint32_t) id;
        rel[i].payload = 0;
        id += jump;
    }
    knuth_shuffle(rel, n);
}

void oracle_gen_fk_sel(oracle_row_t *rel, uint64_t n, int64_t maxid) {
    uint64_t iters = maxid != 0 ? n / (uint64_t) maxid : 0;
    for (uint64_t i = 0; i < iters; ++i)
        unique_gen_maxid(rel + (uint64_t) maxid * i, (uint64_t) maxid, (uint32_t) maxid);
    uint64_t rem = maxid != 0 ? n % (uint64_t) maxid : n;
    if (rem > 0) unique_gen_maxid(rel + (uint64_t) maxid * iters, rem, (uint32_t) maxid);
}

void oracle_zipf_lut(double *lut, uint32_t alphabet_size, double z) {
    double scaling = 0.0;
    for (uint32_t i = 1; i <= alphabet_size; ++i) scaling += 1.0 / pow((double) i, z);
    double sum = 0.0;
    for (uint32_t i = 1; i <= alphabet_size; ++i) {
        sum += 1.0 / pow((double) i, z);
        lut[i - 1] = sum / scaling;
    }
}

uint32_t oracle_zipf_pos(const double *lut, uint32_t alphabet_size, double r) {
    uint32_t left = 0, right = alphabet_size - 1, m;
    if (lut[0] >= r) return 0;
    while (right - left > 1) {
        m = (left + right) / 2;
        if (lut[m] < r) left = m; else right = m;
    }
    return right;
}
