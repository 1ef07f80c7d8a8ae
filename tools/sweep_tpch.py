"""Tuning helper: time the device-resident TPC-H-style pipelines. usage: python tools/sweep_tpch.py [scale factor] [query ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200"))
import b200aqp as A

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
A.init(0)
A.tpch_generate_device(sf, 1)
for q in ([int(a) for a in sys.argv[2:]] or (3, 12, 19)):
    for _ in range(2):
        r = A.tpch_query_device(q)
    runs = [A.tpch_query_device(q) for _ in range(3)]
    ms = sum(x["ms_total"] for x in runs) / len(runs)
    print(f"SF{sf:g} Q{q}: {ms:.3f} ms (filter {runs[-1]['ms_filter']:.3f}, join {runs[-1]['ms_join']:.3f}) rows={r['result_rows']}")
