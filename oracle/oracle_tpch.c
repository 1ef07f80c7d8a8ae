/*
 * oracle_tpch.c — scalar restatement of the reference's TPC-H-style Q3 / Q12 / Q19 pipelines.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Reference: Join-Benchmarks/lib/TPCH-Queries/src/tpch.cpp (tpch_q3 :37-117, tpch_q12 :219-253, tpch_q19
 * :255-309) with the scalar predicates / copy functions of Q3Predicates.hpp:25-55, Q12Predicates.hpp:22-37,
 * Q19Predicates.hpp:27-78 and the transformer copy_Sp_Sp (result_transformers.hpp:51-54). Joins go through
 * oracle_rho(). The queries return row counts only, as the reference does.
 */
#include "oracle.h"
#include <stdlib.h>

#define TS_1994_01_01 757382400ull   /* TpcHTypes.hpp:31-41 */
#define TS_1995_01_01 788918400ull
#define TS_1995_03_15 795225600ull
#define TS_1995_03_16 795312000ull

static oracle_row_t *alloc_rows(uint64_t n) { return (oracle_row_t *) malloc(sizeof(oracle_row_t) * (n ? n : 1)); }

int64_t oracle_tpch_q12(const oracle_lineitem_t *l, const oracle_orders_t *o, uint64_t *filtered) {
    oracle_row_t *f = alloc_rows(l->n);
    uint64_t m = 0;
    for (uint64_t i = 0; i < l->n; ++i) {   /* q12Predicate + q12Copy */
        uint8_t sm = l->l_shipmode[i];
        uint64_t c = l->l_commitdate[i], s = l->l_shipdate[i], r = l->l_receiptdate[i];
        if ((sm == 1 || sm == 2) && c < r && s < c && r >= TS_1994_01_01 && r < TS_1995_01_01) f[m++] = l->l_orderkey[i];
    }
    if (filtered) filtered[0] = m;
    int64_t res = oracle_rho(o->o_orderkey, o->n, f, m, 1, 1, NULL, NULL, NULL, 0);   /* tpch.cpp:240-241 */
    free(f);
    return res;
}

int64_t oracle_tpch_q3(const oracle_customer_t *c, const oracle_orders_t *o, const oracle_lineitem_t *l,
                       uint64_t *filtered, uint64_t *join1_rows) {
    oracle_row_t *fc = alloc_rows(c->n), *fo = alloc_rows(o->n);
    uint64_t nc = 0, no = 0, nl = 0;
    for (uint64_t i = 0; i < c->n; ++i)
        if (c->c_mktsegment[i] == 1) fc[nc++] = c->c_custkey[i];                      /* q3CustomerPredicate/Copy */
    for (uint64_t i = 0; i < o->n; ++i)
        if (o->o_orderdate[i] < TS_1995_03_15) {                                       /* q3OrdersPredicate/Copy   */
            fo[no].key = o->o_custkey[i];
            fo[no].payload = o->o_orderkey[i].key;
            ++no;
        }
    /* join 1, materialised; the output can be larger than |orders'| only if customer keys repeat */
    int64_t cnt = oracle_rho(fc, nc, fo, no, 1, 1, NULL, NULL, NULL, 0);
    oracle_triple_t *t = (oracle_triple_t *) malloc(sizeof(oracle_triple_t) * (cnt ? cnt : 1));
    uint64_t cs = 0, ks = 0;
    oracle_rho(fc, nc, fo, no, 1, 1, &cs, &ks, t, (uint64_t) cnt);
    oracle_row_t *u = alloc_rows((uint64_t) cnt);
    for (int64_t i = 0; i < cnt; ++i) {                                                /* copy_Sp_Sp */
        u[i].key = t[i].Spayload;
        u[i].payload = t[i].Spayload;
    }
    oracle_row_t *fl = alloc_rows(l->n);
    for (uint64_t i = 0; i < l->n; ++i)
        if (l->l_shipdate[i] >= TS_1995_03_16) fl[nl++] = l->l_orderkey[i];          /* q3LineitemPredicate/Copy */
    if (filtered) { filtered[0] = nc; filtered[1] = no; filtered[2] = nl; }
    if (join1_rows) *join1_rows = (uint64_t) cnt;
    int64_t res = oracle_rho(u, (uint64_t) cnt, fl, nl, 1, 1, NULL, NULL, NULL, 0);  /* tpch.cpp:100-101 */
    free(fc); free(fo); free(t); free(u); free(fl);
    return res;
}

static int q19_final(const oracle_part_t *p, uint32_t rp, const oracle_lineitem_t *l, uint32_t rl) {
    uint8_t b = p->p_brand[rp], k = p->p_container[rp];
    uint32_t s = p->p_size[rp];
    float q = l->l_quantity[rl];
    int p1 = b == 1 && (k >= 1 && k <= 4) && (s >= 1 && s <= 5) && (q >= 1 && q <= 1 + 10);
    int p2 = b == 2 && (k >= 5 && k <= 8) && (s >= 1 && s <= 10) && (q >= 10 && q <= 10 + 10);
    int p3 = b == 3 && (k >= 9 && k <= 12) && (s >= 1 && s <= 15) && (q >= 20 && q <= 20 + 10);
    return p1 || p2 || p3;
}

int64_t oracle_tpch_q19(const oracle_lineitem_t *l, const oracle_part_t *p, uint64_t *filtered, uint64_t *join1_rows) {
    oracle_row_t *fp = alloc_rows(p->n), *fl = alloc_rows(l->n);
    uint64_t np = 0, nl = 0;
    for (uint64_t i = 0; i < p->n; ++i) {                                             /* q19PartPredicate/Copy */
        uint8_t b = p->p_brand[i], k = p->p_container[i];
        uint32_t s = p->p_size[i];
        if ((b == 1 || b == 2 || b == 3) && (k >= 1 && k <= 12) && (s >= 1 && s <= 15)) fp[np++] = p->p_partkey[i];
    }
    for (uint64_t i = 0; i < l->n; ++i) {                                             /* q19LineItemPredicate/Copy */
        float q = l->l_quantity[i];
        uint8_t sm = l->l_shipmode[i];
        if ((q >= 1 && q <= 20 + 10) && (sm == 3 || sm == 4) && l->l_shipinstruct[i] == 1) {
            fl[nl].key = l->l_partkey[i];
            fl[nl].payload = l->l_orderkey[i].payload;
            ++nl;
        }
    }
    int64_t cnt = oracle_rho(fp, np, fl, nl, 1, 1, NULL, NULL, NULL, 0);
    oracle_triple_t *t = (oracle_triple_t *) malloc(sizeof(oracle_triple_t) * (cnt ? cnt : 1));
    uint64_t cs = 0, ks = 0;
    oracle_rho(fp, np, fl, nl, 1, 1, &cs, &ks, t, (uint64_t) cnt);
    int64_t matches = 0;
    for (int64_t i = 0; i < cnt; ++i) matches += q19_final(p, t[i].Rpayload, l, t[i].Spayload);   /* tpch.cpp:288-299 */
    if (filtered) { filtered[0] = np; filtered[1] = nl; filtered[2] = (uint64_t) matches; }
    if (join1_rows) *join1_rows = (uint64_t) cnt;
    free(fp); free(fl); free(t);
    return matches;
}
