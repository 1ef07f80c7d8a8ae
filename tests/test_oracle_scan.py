"""Pins oracle_scan.c against closed forms, the committed reference fixtures and the compiled
SIMD512 kernels. CPU only."""
import numpy as np
import pytest

from helpers import sha


def _column(oracle, name, n):
    if name == "tiled":
        return oracle.tiled_column(n)
    return np.random.default_rng(7).integers(0, 256, n, dtype=np.uint8)


def test_closed_form_counts(oracle):
    n = 1 << 18
    col = oracle.tiled_column(n)
    for hi in (0, 26, 128, 255):                      # count = n/256 * (hi+1)
        assert oracle.scan_count(0, hi, col) == n // 256 * (hi + 1)
    assert oracle.scan_count(100, 50, col) == 0       # lo > hi selects nothing


def test_scan_matches_golden(oracle, golden):
    cols = {}
    for c in golden["scan"]:
        col = cols.setdefault(c["column"], _column(oracle, c["column"], c["n"]))
        assert oracle.scan_count(c["lo"], c["hi"], col) == c["count"], c
        assert sha(oracle.bitvector_scan(c["lo"], c["hi"], col)) == c["sha256_bitvector"], c
        assert sha(oracle.index_scan(c["lo"], c["hi"], col)) == c["sha256_rowids"], c


def test_bit_layout_and_tail(oracle):
    col = np.zeros(64 * 3 + 17, dtype=np.uint8)       # 17-value tail is ignored (SIMD512.cpp:216)
    col[[0, 63, 64 + 5, 64 * 3 + 2]] = 9
    bv = oracle.bitvector_scan(9, 9, col)
    assert list(bv) == [(1 << 0) | (1 << 63), 1 << 5, 0]
    assert list(oracle.index_scan(9, 9, col)) == [0, 63, 69]
    assert list(oracle.scalar_index_scan(9, 9, col)) == [0, 63, 69, 64 * 3 + 2]   # ScalarScan.hpp scans all n


def test_against_compiled_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this host")
    n = 1 << 16
    d = oracle.aligned_u8(n)
    d[:] = np.random.default_rng(1).integers(0, 256, n, dtype=np.uint8)
    for lo, hi in ((0, 0), (3, 77), (0, 255), (200, 100), (128, 128)):
        assert np.array_equal(oracle.bitvector_scan(lo, hi, d), oracle.ref_bitvector_scan(lo, hi, d))
        assert np.array_equal(oracle.index_scan(lo, hi, d), oracle.ref_index_scan(lo, hi, d))
        assert np.array_equal(oracle.index_scan(lo, hi, d), oracle.ref_index_scan_self_alloc(lo, hi, d))
        assert oracle.scan_count(lo, hi, d) == oracle.ref_scan_count(lo, hi, d)
