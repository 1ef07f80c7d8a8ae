"""Pins oracle_join.c against closed forms, the committed reference fixtures and the compiled
reference. CPU only."""
import numpy as np
import pytest

from helpers import expected_pkfk, sha, sorted_triples


def test_radix_bits_and_passes(oracle):
    L = oracle.lib()
    # C1 |R|=2^24 -> 9 bits, C3 |R|=2^27 -> 12 bits; 1 pass unless forced
    assert L.oracle_calc_num_radix_bits(1 << 24, 1) == 9
    assert L.oracle_calc_num_radix_bits(1 << 27, 1) == 12
    assert L.oracle_calc_num_radix_bits(1000, 16) == 4     # at least log2(nthreads)
    assert L.oracle_calc_num_passes(13) == 1 and L.oracle_calc_num_passes(14) == 2


def test_closed_form_pkfk(oracle):
    nR, nS = 1 << 16, 1 << 18
    R = oracle.set_rowid_payload(oracle.gen_pk(nR, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(nS, nR, 22222))
    for force2 in (True, False):
        r = oracle.rho(R, S, nthreads=8, force_2_passes=force2)
        assert r["matches"] == nS
        assert r["keysum"] == (nS // nR) * nR * (nR + 1) // 2
        assert (r["matches"], r["checksum"], r["keysum"]) == expected_pkfk(R, S)


def _inputs(oracle, c, zipf_inputs):
    R = oracle.set_rowid_payload(oracle.gen_pk(c["nR"], 11111))
    if c["kind"] == "fk":
        S = oracle.gen_fk(c["nS"], c["nR"], 22222)
    elif c["kind"] == "fk_sel":
        S = oracle.gen_fk_sel(c["nS"], 100 * c["nR"] // c["sel"], 22222)
    else:
        S = np.zeros(c["nS"], dtype=oracle.ROW)
        S["key"] = zipf_inputs[f"S_z{c['z']}"]
    return R, oracle.set_rowid_payload(S)


def test_join_matches_golden(oracle, golden, zipf_inputs):
    for c in golden["join"]:
        R, S = _inputs(oracle, c, zipf_inputs)
        r = oracle.rho(R, S, nthreads=c["nthreads"], force_2_passes=c["force_2_passes"], materialize=True)
        assert (r["matches"], r["checksum"], r["keysum"]) == (c["matches"], c["checksum"], c["keysum"]), c
        assert sha(sorted_triples(r["triples"])) == c["sha256_sorted_triples"], c


def test_duplicate_build_keys_and_misses(oracle):
    rng = np.random.default_rng(5)
    R = np.zeros(3000, dtype=oracle.ROW)
    R["key"] = rng.integers(1, 500, 3000)          # many duplicates -> chains longer than 1
    S = np.zeros(10000, dtype=oracle.ROW)
    S["key"] = rng.integers(1, 1000, 10000)        # half of the probes miss
    oracle.set_rowid_payload(R)
    oracle.set_rowid_payload(S)
    r = oracle.rho(R, S, nthreads=4, materialize=True)
    cnt = np.bincount(R["key"], minlength=1001)
    assert r["matches"] == int(cnt[S["key"]].sum())
    assert len(r["triples"]) == r["matches"]


def test_empty_inputs(oracle):
    E = np.zeros(0, dtype=oracle.ROW)
    R = oracle.gen_pk(100, 1)
    assert oracle.rho(E, R)["matches"] == 0
    assert oracle.rho(R, E)["matches"] == 0


def test_per_pass_partition(oracle):
    R = oracle.gen_pk(10007, 11111)
    out, offs = oracle.radix_partition(R, 3, 5)
    assert offs[-1] == 10007
    for p in range(32):
        part = out[int(offs[p]):int(offs[p + 1])]
        assert ((part["key"] >> 3) & 31 == p).all()
    assert np.array_equal(np.sort(out["key"]), np.sort(R["key"]))


def test_against_compiled_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this host")
    R = oracle.set_rowid_payload(oracle.gen_pk(30011, 1))
    S = oracle.set_rowid_payload(oracle.gen_fk(200003, 30011, 2))
    for force2 in (True, False):
        a = oracle.rho(R, S, nthreads=3, force_2_passes=force2, materialize=True)
        b = oracle.ref_rho(R, S, nthreads=3, materialize=True, force_2_passes=force2)
        assert (a["matches"], a["checksum"], a["keysum"]) == (b["matches"], b["checksum"], b["keysum"])
        assert np.array_equal(sorted_triples(a["triples"]), sorted_triples(b["triples"]))
