"""Pins oracle_tpch.c (scalar restatement of the reference's Q3 / Q12 / Q19 pipelines) against the committed
reference fixtures and, when present, the compiled reference pipelines themselves. CPU only."""
import numpy as np
import pytest


def test_tpch_matches_golden(oracle, golden):
    for c in golden["tpch"]:
        t = oracle.synth_tpch(c["sf"], c["seed"])
        for q in (3, 12, 19):
            r = oracle.tpch_query(q, t)
            assert r["result_rows"] == c[f"q{q}"]["result_rows"], (c["sf"], c["seed"], q)
            if q == 19:
                assert r["join1_rows"] == c["q19"]["join1_rows"]


def test_selection_cardinalities_are_plausible(oracle):
    t = oracle.synth_tpch(0.05, 9)
    nl, no, nc, npart = (len(t[k][c]) for k, c in (("lineitem", "l_shipmode"), ("orders", "o_custkey"),
                                                    ("customer", "c_mktsegment"), ("part", "p_brand")))
    q3 = oracle.tpch_query(3, t)
    assert abs(q3["filtered"][0] / nc - 0.2) < 0.03              # BUILDING is 1 of 5 segments
    assert abs(q3["filtered"][1] / no - 1168 / 2406) < 0.03      # o_orderdate < 1995-03-15
    assert q3["join1_rows"] <= q3["filtered"][1]                 # every order has exactly one customer
    q19 = oracle.tpch_query(19, t)
    assert abs(q19["filtered"][0] / npart - (3 / 25) * (12 / 40) * (15 / 50)) < 0.005
    assert abs(q19["filtered"][1] / nl - (30 / 50) * (1 / 7) * (1 / 4)) < 0.005
    assert q19["result_rows"] <= q19["join1_rows"] <= q19["filtered"][1]
    q12 = oracle.tpch_query(12, t)
    assert q12["result_rows"] == q12["filtered"][0]              # every line item has its order


def test_against_compiled_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this host")
    t = oracle.synth_tpch(0.03, 17)
    for q in (3, 12, 19):
        a, b = oracle.tpch_query(q, t), oracle.ref_tpch_query(q, t, nthreads=3)
        assert a["result_rows"] == b["result_rows"], q
        if q == 19:
            assert a["join1_rows"] == b["join1_rows"]
