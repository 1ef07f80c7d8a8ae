"""Pins oracle_gen.c: against libc rand(), the committed reference fixtures and (when present) the
compiled reference generator itself. CPU only."""
import ctypes

import numpy as np
import pytest

from helpers import sha


def test_oracle_rand_is_glibc_rand(oracle):
    libc = ctypes.CDLL("libc.so.6")
    L = oracle.lib()
    for seed in (0, 1, 11111, 22222, 12345, 2 ** 32 - 1):
        libc.srand(seed)
        L.oracle_srand(seed)
        for _ in range(2000):
            assert libc.rand() == L.oracle_rand()


def test_survey_known_answers(oracle):
    # SURVEY.md §4 probed prefixes
    assert list(oracle.gen_pk(1 << 20, 11111)["key"][:4]) == [533741, 233869, 796176, 204571]


def _gen(oracle, c):
    if c["kind"] == "pk":
        return oracle.gen_pk(c["n"], c["seed"])
    if c["kind"] == "fk":
        return oracle.gen_fk(c["n"], c["maxid"], c["seed"])
    return oracle.gen_fk_sel(c["n"], c["maxid"], c["seed"])


def test_generators_match_golden(oracle, golden):
    for c in golden["generator"]:
        rel = _gen(oracle, c)
        assert [int(x) for x in rel["key"][:8]] == c["first8"], c
        assert sha(rel["key"]) == c["sha256_keys"], c
        assert not rel["payload"].any()


def test_pk_is_permutation_fk_is_blocks_of_permutations(oracle):
    R = oracle.gen_pk(5000, 3)
    assert np.array_equal(np.sort(R["key"]), np.arange(1, 5001, dtype=np.uint32))
    S = oracle.gen_fk(5000 * 3 + 77, 5000, 4)
    for b in range(3):
        assert np.array_equal(np.sort(S["key"][b * 5000:(b + 1) * 5000]), np.arange(1, 5001, dtype=np.uint32))
    # remainder block holds keys 1..rem only (generator.cpp:499-504)
    assert np.array_equal(np.sort(S["key"][15000:]), np.arange(1, 78, dtype=np.uint32))


def test_against_compiled_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this host")
    for n, seed in ((1 << 12, 99), (70001, 11111)):
        assert np.array_equal(oracle.gen_pk(n, seed), oracle.ref_gen_pk(n, seed))
    assert np.array_equal(oracle.gen_fk(200000, 70001, 5), oracle.ref_gen_fk(200000, 70001, 5))
    assert np.array_equal(oracle.gen_fk_sel(1 << 15, 100 * (1 << 13) // 25, 6),
                          oracle.ref_gen_fk_sel(1 << 15, 100 * (1 << 13) // 25, 6))


def test_zipf_lut_and_search(oracle):
    n = 1000
    lut = np.empty(n)
    oracle.lib().oracle_zipf_lut(lut.ctypes.data, n, 1.0)
    w = 1.0 / np.arange(1, n + 1)
    assert np.allclose(lut, np.cumsum(w) / w.sum(), rtol=1e-12)   # tolerance: fp64 summation order
    assert abs(lut[-1] - 1.0) < 1e-12
    L = oracle.lib()
    assert L.oracle_zipf_pos(lut.ctypes.data, n, 0.0) == 0
    assert L.oracle_zipf_pos(lut.ctypes.data, n, float(lut[0])) == 0
    assert L.oracle_zipf_pos(lut.ctypes.data, n, float(lut[0]) + 1e-9) == 1
    assert L.oracle_zipf_pos(lut.ctypes.data, n, 0.999999999) == n - 1
