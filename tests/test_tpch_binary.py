"""The reference's binary column format for TPC-H tables (Join-Benchmarks/App/TpcH/CSVConvert.cpp:16-190 writes it,
TpcHCommons.cpp:194-214,:235-295,:423-451,:506-537,:594-623 reads it), through b200_tpch_write_binary /
b200_tpch_read_binary. Host code only: runs without a GPU."""
import os

import numpy as np


def test_layout_matches_the_reference_format(aqp, oracle, tmp_path):
    t = oracle.synth_tpch(0.02, 3)
    aqp.tpch_write_binary(str(tmp_path), 7, t)
    # getPath(): "scale" + setw(3) setfill('0'); CSVConvert: <table>.tbl.dir/size + <column>.bin, raw little-endian arrays
    for table, cols in t.items():
        d = tmp_path / "scale007" / f"{table}.tbl.dir"
        n = len(next(iter(cols.values())))
        assert (d / "size").read_text().strip() == str(n)
        for col, a in cols.items():
            raw = np.fromfile(d / f"{col}.bin", dtype=a.dtype)
            assert raw.shape == a.shape and (raw == a).all(), (table, col)
    back = aqp.tpch_read_binary(str(tmp_path), 7)
    for table, cols in t.items():
        assert set(back[table]) == set(cols)
        for col, a in cols.items():
            assert back[table][col].dtype == a.dtype and (back[table][col] == a).all(), (table, col)


def test_reads_files_written_the_reference_way(aqp, oracle, tmp_path):
    """files laid down as CSVConvert.cpp does (here with numpy), with the per-query column subsets the reference's loaders
    expect (TpcHCommons.cpp:246-263: Q12 needs l_orderkey, three dates and l_shipmode) - absent columns stay absent"""
    t = oracle.synth_tpch(0.01, 11)
    d = tmp_path / "scale001" / "lineitem.tbl.dir"
    os.makedirs(d)
    n = len(t["lineitem"]["l_shipmode"])
    (d / "size").write_text(str(n))
    for col in ("l_orderkey", "l_shipdate", "l_commitdate", "l_receiptdate", "l_shipmode"):
        t["lineitem"][col].tofile(d / f"{col}.bin")
    back = aqp.tpch_read_binary(str(tmp_path), 1)
    assert set(back["lineitem"]) == {"l_orderkey", "l_shipdate", "l_commitdate", "l_receiptdate", "l_shipmode"}
    assert all((back["lineitem"][c] == t["lineitem"][c]).all() for c in back["lineitem"])
    assert back["orders"] == {} and back["part"] == {} and back["customer"] == {}
    # a column file shorter than the size file is an error, not a silent truncation
    t["lineitem"]["l_shipmode"][: n // 2].tofile(d / "l_shipmode.bin")
    try:
        aqp.tpch_read_binary(str(tmp_path), 1)
        raise AssertionError("short column file accepted")
    except aqp.AqpError as e:
        assert "fewer rows" in str(e)
