"""Link-level proof of the drop-in boundary: the reference's UNMODIFIED query pipelines
(Join-Benchmarks/lib/TPCH-Queries/src/tpch.cpp, result_transformers.cpp, time_print.cpp and its logger), compiled from
/root/reference by oracle/Makefile's `dropin` target WITHOUT the reference's join library and linked against
libb200aqp.so through shim/b200aqp_cxx_shim.cpp (C++-mangled run_join / destroy_table -> the C symbols).
Every run_join call of tpch.cpp (:68,:101,:141,:167,:202,:241,:282) then executes on the GPU, and the materialised
chunked_table_t results go back through the reference's own transformers (result_transformers.hpp:77-127,
Q19Predicates.hpp:144-164) — which only works if chunk layout, counts and ownership are exactly the reference's."""
import pytest

pytestmark = pytest.mark.gpu


def test_reference_tpch_pipelines_on_gpu_join(gpu, oracle, golden):
    if not oracle.have_dropin():
        pytest.skip("oracle/_ref/libdropin_tpch.so not built (needs /root/reference at build time) or no AVX-512 host")
    for c in golden["tpch"]:
        t = oracle.synth_tpch(c["sf"], c["seed"])
        for q in (3, 12, 19):
            got = oracle.dropin_tpch_query(q, t, nthreads=4)
            assert got["result_rows"] == c[f"q{q}"]["result_rows"], (c["sf"], q, got, c[f"q{q}"])
            if q == 19:
                assert got["join1_rows"] == c["q19"]["join1_rows"]


def test_dropin_matches_oracle_on_fresh_tables(gpu, oracle):
    if not oracle.have_dropin():
        pytest.skip("oracle/_ref/libdropin_tpch.so not built")
    t = oracle.synth_tpch(0.2, 1234)
    for q in (3, 12, 19):
        exp = oracle.tpch_query(q, t)
        for _ in range(2):           # twice: the second run reuses the library's cached result slab
            got = oracle.dropin_tpch_query(q, t, nthreads=2)
            assert got["result_rows"] == exp["result_rows"], (q, got, exp)
