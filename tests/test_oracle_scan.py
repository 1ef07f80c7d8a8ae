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


# ---- the remaining SIMD512 variants (sum, value list, dictionary scan): the reference's OWN known-answer tests ----
def dict_kats(O):
    """The [main] dictionary-scan cases of Scan-Micro-Benchmarks/shared_libraries/SimdScan/tests/testsimdscan.cpp
    (8-bit column; :8-165 and :217-245), restated: (name, column, dictionary, low, high, expected size, spot checks)."""
    n = 1 << 20
    tiled = O.tiled_column(n)                                       # allocate_data_array_aligned<uint8_t>, Allocator.hpp:95-109
    mod4 = (np.arange(n) & 3).astype(np.uint8)                       # test 4: gen = in & 3
    ident = np.arange(256, dtype=np.int64)                           # allocate_data_array_aligned<int64_t>(256): dict[i] = i
    return [
        ("test 1 :8-28", tiled, ident, 0, 100, n // 256 * 101, {}),
        ("test 2 :30-54", tiled, ident, 1, 100, n // 256 * 100, {**{i: i + 1 for i in range(100)}, 100: 1}),
        ("test 3 :56-84", tiled, ident * 2, 0, 98, n // 256 * 50, {**{i: 2 * i for i in range(50)}, 99: 98, 100: 0}),
        ("test 4 :86-113", mod4, ident, 0, 2, n // 4 * 3, {0: 0, 1: 1, 2: 2, 3: 0}),
        ("test 5 :115-138", tiled, ident, 100, 199, n // 256 * 100, {0: 100, 99: 199}),
        ("test 6 :140-165", tiled, ident - 128, -10, 0, n // 256 * 11, {0: -10, 10: 0}),
        ("scalar gather scatter :217-245", tiled, ident * 2, 0, 98, n // 256 * 50, {**{i: 2 * i for i in range(50)}, 99: 98, 100: 0}),
    ]


def test_dict_scan_reference_known_answers(oracle):
    for name, col, d, lo, hi, size, spots in dict_kats(oracle):
        r = oracle.dict_scan_8_64(lo, hi, d, col)
        assert len(r) == size, name
        for i, v in spots.items():
            assert r[i] == v, (name, i)


def test_sum_value_dict_against_compiled_reference(oracle):
    if not (oracle.have_ref() and oracle.host_has_avx512()):
        pytest.skip("compiled reference needs AVX-512 and oracle/_ref")
    rng = np.random.default_rng(5)
    n = 64 * 1000 + 37
    col = oracle.aligned_u8(n)
    col[:] = rng.integers(0, 256, n, dtype=np.uint8)
    for lo, hi in [(0, 0), (0, 26), (5, 5), (17, 200), (100, 50), (0, 255), (255, 255)]:
        assert oracle.scan_sum(lo, hi, col) == oracle.ref_scan_sum(lo, hi, col), (lo, hi)
        assert np.array_equal(oracle.value_scan(lo, hi, col), oracle.ref_value_scan(lo, hi, col)), (lo, hi)
    dicts = [np.arange(256, dtype=np.int64), np.arange(256, dtype=np.int64) * 3 - 300,
             np.sort(rng.integers(-10**12, 10**12, 256))]
    for d in dicts:
        preds = [(int(d[10]), int(d[90])), (int(d[0]) - 5, int(d[0]) - 1), (int(d[255]) + 1, int(d[255]) + 9),
                 (int(d[200]), int(d[100])), (int(d[17]), int(d[17])), (int(d[0]), int(d[255]))]
        for lo, hi in preds:
            exp = oracle.ref_dict_scan_8_64(lo, hi, d, col)
            assert np.array_equal(oracle.dict_scan_8_64(lo, hi, d, col), exp), (lo, hi)
            for which in (1, 2, 3):   # the reference's other 8-bit implementations agree with each other
                assert np.array_equal(oracle.ref_dict_scan_8_64(lo, hi, d, col, which), exp), (which, lo, hi)


# ---- 16- / 32-bit dictionary scans and the explicit-index scan ------------------------------------------------------
def wide_dict_kats():
    """The reference's [main] cases for the wider codes (testsimdscan.cpp:167-215 and :478-528), restated:
    (name, code bits, column, dictionary, low, high, expected size, value check). allocate_data_array_aligned<T>(n)
    fills element i with (T) i; the 32-bit test 1 uses gen = in & 255."""
    n = 1 << 20
    c16 = (np.arange(n) & 0xffff).astype(np.uint16)
    d16 = np.arange(1 << 16, dtype=np.int64)
    c32_mod = (np.arange(n) & 255).astype(np.uint32)
    c32_id = np.arange(n, dtype=np.uint32)
    d32 = np.arange(1 << 20, dtype=np.int64)
    return [
        ("Dict 16_64 scan test 1 :167-190", 16, c16, d16, 0, 299, n // (1 << 16) * 300, lambda r: (r == np.arange(len(r)) % 300).all()),
        ("Dict 32_64 scan test 1 :192-215", 32, c32_mod, d32, 0, 299, n, None),
        ("Dict 32_64 scalar gather scatter test 1 :478-502", 32, c32_mod, d32, 0, 299, n, None),
        ("Dict 32_64 scalar gather scatter test 2 :504-528", 32, c32_id, d32, 0, 99, 100, lambda r: r[0] == 0 and r[10] == 10),
    ]


def test_wide_dict_scan_reference_known_answers(oracle):
    for name, bits, col, d, lo, hi, size, check in wide_dict_kats():
        r = oracle.dict_scan_wide(bits, lo, hi, d, col)
        assert len(r) == size, name
        assert check is None or check(r), name


def test_wide_dict_and_explicit_against_compiled_reference(oracle):
    if not (oracle.have_ref() and oracle.host_has_avx512()):
        pytest.skip("compiled reference needs AVX-512 and oracle/_ref")
    rng = np.random.default_rng(9)
    n = 64 * 500 + 21
    # explicit index scan: block i's byte group j reads index register i + j (SIMD512.cpp:176) - as written
    col = oracle.aligned_u8(n)
    col[:] = rng.integers(0, 256, n, dtype=np.uint8)
    index = oracle.aligned((n // 64 + 7) * 8, np.uint64)
    index[:] = rng.integers(0, 1 << 62, len(index), dtype=np.uint64)
    for lo, hi in [(0, 26), (7, 7), (100, 50), (0, 255), (200, 255)]:
        assert np.array_equal(oracle.explicit_index_scan(lo, hi, index, col), oracle.ref_explicit_index_scan(lo, hi, index, col)), (lo, hi)
    # 16-bit codes, sorted 65536-entry dictionary; predicates inside, below, above the dictionary and inverted
    d16 = np.sort(rng.integers(-10**12, 10**12, 1 << 16))
    c16 = oracle.aligned(n, np.uint16)
    c16[:] = rng.integers(0, 1 << 16, n, dtype=np.uint16)
    for lo, hi in [(int(d16[100]), int(d16[9000])), (int(d16[0]) - 9, int(d16[0]) - 1), (int(d16[-1]) + 1, int(d16[-1]) + 5),
                   (int(d16[500]), int(d16[400])), (int(d16[77]), int(d16[77])), (int(d16[0]), int(d16[-1]))]:
        assert np.array_equal(oracle.dict_scan_wide(16, lo, hi, d16, c16), oracle.ref_dict_scan_wide(16, lo, hi, d16, c16)), (lo, hi)
    # 32-bit codes: the reference narrows the code range through uint16_t (SIMD512.cpp:592-593) - dictionaries up to and
    # beyond 65536 entries behave as written there
    for dsize in (1000, 1 << 16, 100000):
        d32 = np.sort(rng.integers(-10**12, 10**12, dsize))
        c32 = oracle.aligned(n, np.uint32)
        c32[:] = rng.integers(0, dsize, n, dtype=np.uint32)
        for lo, hi in [(int(d32[10]), int(d32[dsize // 2])), (int(d32[0]) - 3, int(d32[0]) - 1), (int(d32[-1]) + 1, int(d32[-1]) + 2),
                       (int(d32[dsize - 5]), int(d32[dsize - 1])), (int(d32[3]), int(d32[3]))]:
            assert np.array_equal(oracle.dict_scan_wide(32, lo, hi, d32, c32), oracle.ref_dict_scan_wide(32, lo, hi, d32, c32)), (dsize, lo, hi)
