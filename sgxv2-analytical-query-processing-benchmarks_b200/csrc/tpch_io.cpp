// tpch_io.cpp — the reference's binary column format for TPC-H tables, reader and writer.
//
// Join-Benchmarks/App/TpcH/CSVConvert.cpp:16-190 writes, TpcHCommons.cpp:194-214,:235-295,:423-451,:506-537,:594-623 reads:
//   <root>/scaleNNN/<table>.tbl.dir/size            decimal row count (ASCII)
//   <root>/scaleNNN/<table>.tbl.dir/<column>.bin    the column as a raw little-endian array
// with <table> in lineitem, orders, customer, part and the column types of TpcHTypes.hpp:50-83 (key columns are
// tuple_t = {key, row id}, dates uint64 epoch seconds, dictionary codes uint8, part keys / sizes / customer keys
// uint32, l_quantity float). The reference reads a query-specific subset of the files; here every column file that
// exists is read, a missing one leaves its pointer NULL (b200_tpch_upload skips NULL columns).
// Host code only: real dbgen data converted by the reference's CSVConvert loads straight into the GPU pipelines with
//   b200_tpch_read_binary(root, scale, &l, &o, &c, &p); b200_tpch_upload(&l, &o, &c, &p); b200_tpch_q12_device(&stats);
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "aqp/b200_aqp.h"
#include "aqp/b200_tpch.h"
#include "common.cuh"

namespace {

std::string table_dir(const char *root, int scale, const char *table) {
    char buf[32];
    snprintf(buf, sizeof buf, "scale%03d", scale);   // getPath(): setw(3) setfill('0')
    return std::string(root) + "/" + buf + "/" + table + ".tbl.dir";
}

bool read_size(const std::string &dir, uint64_t *n) {
    FILE *f = fopen((dir + "/size").c_str(), "r");
    if (!f) return false;
    unsigned long long v = 0;
    const bool ok = fscanf(f, "%llu", &v) == 1;
    fclose(f);
    *n = v;
    return ok;
}

// NULL if the file does not exist; aborts the whole read (returns false) if it exists but is short
template <typename X>
bool read_col(const std::string &dir, const char *name, uint64_t n, X **out) {
    *out = nullptr;
    FILE *f = fopen((dir + "/" + name + ".bin").c_str(), "rb");
    if (!f) return true;
    void *p = nullptr;
    if (posix_memalign(&p, 64, (n ? n : 1) * sizeof(X)) != 0) {   // 64-byte aligned like the reference's loader
        fclose(f);
        aqp::set_error("tpch binary read: out of host memory");
        return false;
    }
    const size_t got = fread(p, sizeof(X), n, f);
    fclose(f);
    if (got != n) {
        free(p);
        aqp::set_error(dir + "/" + name + ".bin holds fewer rows than the size file says");
        return false;
    }
    *out = static_cast<X *>(p);
    return true;
}

template <typename X>
bool write_col(const std::string &dir, const char *name, uint64_t n, const X *col) {
    if (!col) return true;
    FILE *f = fopen((dir + "/" + name + ".bin").c_str(), "wb");
    if (!f) {
        aqp::set_error("tpch binary write: cannot create " + dir + "/" + name + ".bin");
        return false;
    }
    const bool ok = fwrite(col, sizeof(X), n, f) == n;
    fclose(f);
    if (!ok) aqp::set_error("tpch binary write: short write to " + dir + "/" + name + ".bin");
    return ok;
}

bool make_dirs(const char *root, int scale, const std::string &dir) {
    char buf[32];
    snprintf(buf, sizeof buf, "scale%03d", scale);
    mkdir(root, 0777);
    mkdir((std::string(root) + "/" + buf).c_str(), 0777);
    if (mkdir(dir.c_str(), 0777) != 0) {
        struct stat st;
        if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) {
            aqp::set_error("tpch binary write: cannot create directory " + dir);
            return false;
        }
    }
    FILE *f = fopen((dir + "/size").c_str(), "w");
    return f != nullptr && (fclose(f), true);
}

bool write_size(const std::string &dir, uint64_t n) {
    FILE *f = fopen((dir + "/size").c_str(), "w");
    if (!f) return false;
    fprintf(f, "%llu", (unsigned long long) n);
    fclose(f);
    return true;
}

}  // namespace

extern "C" {

int b200_tpch_read_binary(const char *root, int scale, struct LineItemTable *l, struct OrdersTable *o, struct CustomerTable *c,
                          struct PartTable *p) {
    bool ok = true;
    if (l) {
        memset(l, 0, sizeof *l);
        const std::string d = table_dir(root, scale, "lineitem");
        if (read_size(d, &l->numTuples)) {
            const uint64_t n = l->numTuples;
            ok = ok && read_col(d, "l_orderkey", n, &l->l_orderkey) && read_col(d, "l_shipdate", n, &l->l_shipdate) &&
                 read_col(d, "l_commitdate", n, &l->l_commitdate) && read_col(d, "l_receiptdate", n, &l->l_receiptdate) &&
                 read_col(d, "l_shipmode", n, &l->l_shipmode) && read_col(d, "l_partkey", n, &l->l_partkey) &&
                 read_col(d, "l_quantity", n, &l->l_quantity) && read_col(d, "l_shipinstruct", n, &l->l_shipinstruct) &&
                 read_col(d, "l_returnflag", n, &l->l_returnflag);
        }
    }
    if (o) {
        memset(o, 0, sizeof *o);
        const std::string d = table_dir(root, scale, "orders");
        if (read_size(d, &o->numTuples)) {
            const uint64_t n = o->numTuples;
            ok = ok && read_col(d, "o_orderkey", n, &o->o_orderkey) && read_col(d, "o_orderdate", n, &o->o_orderdate) &&
                 read_col(d, "o_custkey", n, &o->o_custkey);
        }
    }
    if (c) {
        memset(c, 0, sizeof *c);
        const std::string d = table_dir(root, scale, "customer");
        if (read_size(d, &c->numTuples)) {
            const uint64_t n = c->numTuples;
            ok = ok && read_col(d, "c_custkey", n, &c->c_custkey) && read_col(d, "c_mktsegment", n, &c->c_mktsegment) &&
                 read_col(d, "c_nationkey", n, &c->c_nationkey);
        }
    }
    if (p) {
        memset(p, 0, sizeof *p);
        const std::string d = table_dir(root, scale, "part");
        if (read_size(d, &p->numTuples)) {
            const uint64_t n = p->numTuples;
            ok = ok && read_col(d, "p_partkey", n, &p->p_partkey) && read_col(d, "p_brand", n, &p->p_brand) &&
                 read_col(d, "p_size", n, &p->p_size) && read_col(d, "p_container", n, &p->p_container);
        }
    }
    return ok ? 0 : -1;
}

int b200_tpch_write_binary(const char *root, int scale, const struct LineItemTable *l, const struct OrdersTable *o,
                           const struct CustomerTable *c, const struct PartTable *p) {
    bool ok = true;
    if (l && l->numTuples) {
        const std::string d = table_dir(root, scale, "lineitem");
        const uint64_t n = l->numTuples;
        ok = ok && make_dirs(root, scale, d) && write_size(d, n) && write_col(d, "l_orderkey", n, l->l_orderkey) &&
             write_col(d, "l_shipdate", n, l->l_shipdate) && write_col(d, "l_commitdate", n, l->l_commitdate) &&
             write_col(d, "l_receiptdate", n, l->l_receiptdate) && write_col(d, "l_shipmode", n, l->l_shipmode) &&
             write_col(d, "l_partkey", n, l->l_partkey) && write_col(d, "l_quantity", n, l->l_quantity) &&
             write_col(d, "l_shipinstruct", n, l->l_shipinstruct) && write_col(d, "l_returnflag", n, l->l_returnflag);
    }
    if (o && o->numTuples) {
        const std::string d = table_dir(root, scale, "orders");
        const uint64_t n = o->numTuples;
        ok = ok && make_dirs(root, scale, d) && write_size(d, n) && write_col(d, "o_orderkey", n, o->o_orderkey) &&
             write_col(d, "o_custkey", n, o->o_custkey) && write_col(d, "o_orderdate", n, o->o_orderdate);
    }
    if (c && c->numTuples) {
        const std::string d = table_dir(root, scale, "customer");
        const uint64_t n = c->numTuples;
        ok = ok && make_dirs(root, scale, d) && write_size(d, n) && write_col(d, "c_custkey", n, c->c_custkey) &&
             write_col(d, "c_mktsegment", n, c->c_mktsegment) && write_col(d, "c_nationkey", n, c->c_nationkey);
    }
    if (p && p->numTuples) {
        const std::string d = table_dir(root, scale, "part");
        const uint64_t n = p->numTuples;
        ok = ok && make_dirs(root, scale, d) && write_size(d, n) && write_col(d, "p_partkey", n, p->p_partkey) &&
             write_col(d, "p_brand", n, p->p_brand) && write_col(d, "p_container", n, p->p_container) &&
             write_col(d, "p_size", n, p->p_size);
    }
    return ok ? 0 : -1;
}

}  // extern "C"
