"""GPU parity: the RHO join through the C ABI vs the CPU oracle — match count, checksum and keysum
bit-exact, materialised triples identical as a set — on the reference-seeded inputs, the committed
reference fixtures, edge cases, and closed forms at the full benchmark sizes."""
import os

import numpy as np
import pytest

from helpers import expected_pkfk, sha, sorted_triples

pytestmark = pytest.mark.gpu


def _check(gpu, oracle, R, S, materialize=True, expect=None):
    o = oracle.rho(R, S, nthreads=4, materialize=materialize) if expect is None else expect
    g = gpu.run_join(R, S, materialize=materialize, nthreads=4)
    assert (g["matches"], g["checksum"], g["keysum"]) == (o["matches"], o["checksum"], o["keysum"])
    assert g["result_type"] == 1 and g["nthreads"] == 4
    if materialize:
        assert g["table_num_tuples"] == o["matches"]
        assert np.array_equal(sorted_triples(g["triples"]), sorted_triples(o["triples"]))
    return g


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


def test_join_matches_golden(gpu, oracle, golden, zipf_inputs):
    for c in golden["join"]:
        if not c["force_2_passes"]:
            continue   # same inputs, same expected values
        R, S = _inputs(oracle, c, zipf_inputs)
        g = gpu.run_join(R, S, materialize=True)
        assert (g["matches"], g["checksum"], g["keysum"]) == (c["matches"], c["checksum"], c["keysum"]), c
        assert sha(sorted_triples(g["triples"])) == c["sha256_sorted_triples"], c


@pytest.mark.parametrize("nR,nS", [(1, 1), (1, 1000), (7, 3), (8192, 8192), (8193, 50000), (100003, 250007),
                                   (1 << 16, 1 << 18), (1 << 20, 1 << 22), (3_000_017, 5_000_011)])
def test_pkfk_vs_oracle(gpu, oracle, nR, nS):
    R = oracle.set_rowid_payload(oracle.gen_pk(nR, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(nS, nR, 22222))
    g = _check(gpu, oracle, R, S, materialize=nS <= (1 << 22))
    assert g["matches"] == nS


def test_config1_reference_seeds(gpu, oracle):
    """BASELINE config 1: |R|=2^24, |S|=2^26, seeds 11111/22222 (native.cpp:35-36); expected values are
    the survey's known answers and the oracle."""
    nR, nS = 1 << 24, 1 << 26
    R = oracle.set_rowid_payload(oracle.gen_pk(nR, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(nS, nR, 22222))
    assert list(R["key"][:4]) == [14386915, 8878978, 11266848, 11599710]
    assert list(S["key"][:4]) == [4221318, 11571670, 11656028, 6314437]
    g = gpu.run_join(R, S, materialize=False)
    assert g["matches"] == 67108864 and g["keysum"] == 562949986975744
    o = oracle.rho(R, S, nthreads=8)
    assert (g["matches"], g["checksum"], g["keysum"]) == (o["matches"], o["checksum"], o["keysum"])
    assert (g["radix_bits"], g["num_passes"]) == (11, 2)


def test_duplicates_misses_and_key_zero(gpu, oracle):
    rng = np.random.default_rng(5)
    R = np.zeros(30000, dtype=oracle.ROW)
    R["key"] = rng.integers(0, 5000, 30000)           # duplicates (chains > 1) and key 0
    S = np.zeros(100000, dtype=oracle.ROW)
    S["key"] = rng.integers(0, 10000, 100000)         # half miss
    oracle.set_rowid_payload(R)
    oracle.set_rowid_payload(S)
    g = _check(gpu, oracle, R, S)
    assert g["matches"] > len(S)                      # output larger than |S|: exercises the capacity re-run


def test_empty_relations(gpu, oracle):
    E = np.zeros(0, dtype=oracle.ROW)
    R = oracle.set_rowid_payload(oracle.gen_pk(1000, 1))
    for a, b in ((E, R), (R, E), (E, E)):
        g = gpu.run_join(a, b, materialize=True)
        # a materialised result has one (empty) chunk per thread, like the reference's concatenated per-thread lists
        assert g["matches"] == 0 and g["checksum"] == 0 and len(g["triples"]) == 0 and g["num_chunks"] == 1
        g = gpu.run_join(a, b, materialize=True, nthreads=6)
        assert g["matches"] == 0 and g["num_chunks"] == 6 and g["table_num_tuples"] == 0


def test_sparse_keys_unbalanced_partitions(gpu, oracle):
    """TPC-H-style sparse keys (dbgen orderkeys use 8 of every 32 values) and keys that agree on all
    radix bits: partitions are unbalanced and exceed the shared-memory table, forcing several build rounds."""
    n = 200000
    i = np.arange(n, dtype=np.uint64)
    R = np.zeros(n, dtype=oracle.ROW)
    R["key"] = ((i // 8) * 32 + (i % 8) + 1).astype(np.uint32)
    S = np.zeros(4 * n, dtype=oracle.ROW)
    S["key"] = np.random.default_rng(2).choice(R["key"], 4 * n)
    oracle.set_rowid_payload(R)
    oracle.set_rowid_payload(S)
    _check(gpu, oracle, R, S)
    R2 = R.copy()
    R2["key"] = (np.arange(n, dtype=np.uint32) << 12) | 5      # every key lands in ONE partition
    S2 = S.copy()
    S2["key"] = np.random.default_rng(3).choice(R2["key"], 4 * n)
    _check(gpu, oracle, R2, S2)


def test_zipf_skew(gpu, oracle):
    nR, nS = 1 << 18, 1 << 21
    R = oracle.set_rowid_payload(oracle.gen_pk(nR, 11111))
    for z in (0.5, 1.0, 1.5):
        S = oracle.set_rowid_payload(oracle.gen_zipf(nS, nR, z, seed=9))
        g = _check(gpu, oracle, R, S, materialize=False)
        assert g["matches"] == nS


@pytest.mark.parametrize("bits", [0, 3, 8, 9, 16])
def test_forced_radix_bits(gpu, oracle, bits):
    """Every pass configuration (none / one pass / two passes) gives the same result."""
    R = oracle.set_rowid_payload(oracle.gen_pk(300007, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(1000003, 300007, 22222))
    os.environ["B200_AQP_RADIX_BITS"] = str(bits)
    try:
        g = _check(gpu, oracle, R, S, materialize=False)
        assert g["radix_bits"] == bits
    finally:
        del os.environ["B200_AQP_RADIX_BITS"]


@pytest.mark.parametrize("bits", [14, 15])
def test_histogram_free_plan(gpu, oracle, bits, monkeypatch):
    """Count joins with two passes and >= 128 pass-1 partitions run without a histogram pass: every partition gets a region
    of fixed capacity (plan_flags 1). A sample of the inputs that shows skew - Zipf - declines the plan (4); a region that
    overflows anyway sends the join back to exact offsets (2).
    Same results either way and with the plan switched off. Forced at a small size through the radix-bit override."""
    nR, nS = 300007, 1000003
    R = oracle.set_rowid_payload(oracle.gen_pk(nR, 11111))
    fk = oracle.set_rowid_payload(oracle.gen_fk(nS, nR, 22222))
    miss = fk.copy()
    miss["key"][::3] += nR
    dup = R.copy()
    dup["key"] = dup["key"] // 4 + 1
    zipf = oracle.set_rowid_payload(oracle.gen_zipf(nS, nR, 1.0, seed=9))
    monkeypatch.setenv("B200_AQP_RADIX_BITS", str(bits))
    monkeypatch.setenv("B200_AQP_HISTFREE", "1")
    for name, r, s, flags in (("fk", R, fk, 1), ("miss", R, miss, 1), ("dup", dup, fk, None), ("zipf", R, zipf, 4)):
        g = _check(gpu, oracle, r, s, materialize=False)
        assert g["radix_bits"] == bits and g["num_passes"] == 2, name
        assert flags is None or g["plan_flags"] == flags, (name, g["plan_flags"])
    monkeypatch.setenv("B200_AQP_HISTFREE_NOSAMPLE", "1")    # without the sampled test the regions themselves catch it
    g = _check(gpu, oracle, R, zipf, materialize=False)
    assert g["plan_flags"] == 2
    monkeypatch.delenv("B200_AQP_HISTFREE_NOSAMPLE")
    monkeypatch.setenv("B200_AQP_HISTFREE", "0")
    g = _check(gpu, oracle, R, fk, materialize=False)
    assert g["plan_flags"] == 0
    g = _check(gpu, oracle, R, fk, materialize=True)     # materialising joins keep exact offsets
    assert g["plan_flags"] == 0


def test_preload_split(gpu, oracle):
    R = oracle.set_rowid_payload(oracle.gen_pk(1 << 16, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(1 << 18, 1 << 16, 22222))
    o = oracle.rho(R, S)
    gpu.preload_relations(R, S)
    for _ in range(3):
        g = gpu.join_preload()
        assert (g["matches"], g["checksum"], g["keysum"]) == (o["matches"], o["checksum"], o["keysum"])
        assert g["ms_h2d"] == 0.0
    gpu.lib().b200_free_preload()
    with pytest.raises(gpu.AqpError):
        gpu.join_preload()


def test_host_join_overlapped_with_its_copy(gpu, oracle, monkeypatch):
    """B200_AQP_E2E_CHUNKS: the probe side travels in chunks and R is joined with each as it lands - same result as one join,
    for pageable and pinned relations, duplicates and misses included; materialising joins ignore the switch"""
    nR, nS = 1 << 16, (1 << 18) + 12345
    R = oracle.set_rowid_payload(oracle.gen_pk(nR, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(nS, nR, 22222))
    S["key"][::5] += nR            # misses
    R["key"][1::7] = R["key"][::7][:len(R["key"][1::7])]   # duplicate build keys
    o = oracle.rho(R, S, nthreads=2, materialize=True)
    for k in ("1", "4", "7"):
        monkeypatch.setenv("B200_AQP_E2E_CHUNKS", k)
        g = gpu.run_join(R, S, materialize=False)
        assert (g["matches"], g["checksum"], g["keysum"]) == (o["matches"], o["checksum"], o["keysum"]), k
        g = gpu.run_join(R, S, materialize=True)
        assert np.array_equal(sorted_triples(g["triples"]), sorted_triples(o["triples"])), k
    monkeypatch.delenv("B200_AQP_E2E_CHUNKS")


def test_chunked_table_layout(gpu, oracle):
    R = oracle.set_rowid_payload(oracle.gen_pk(5000, 1))
    S = oracle.set_rowid_payload(oracle.gen_fk(20000, 5000, 2))
    g = gpu.run_join(R, S, materialize=True)
    assert g["num_chunks"] == -(-20000 // gpu.TUPLES_PER_CHUNK)
    assert len(g["triples"]) == 20000
    # never fewer chunks than threads: thread t of the reference's chunk consumers starts at chunks[t]
    # (Q19Predicates.hpp:147-151); the surplus chunks are empty
    g32 = gpu.run_join(R, S, materialize=True, nthreads=32)
    assert g32["num_chunks"] == 32 and len(g32["triples"]) == 20000
    m, cs, ks = expected_pkfk(R, S)
    t = g["triples"]
    assert int(t["Rpayload"].astype(np.uint64).sum() + t["Spayload"].astype(np.uint64).sum()) == cs
    assert int(t["key"].astype(np.uint64).sum()) == ks


@pytest.mark.parametrize("logR,logS", [(24, 26), (27, 29)])
def test_full_size_device_generated(gpu, logR, logS):
    """BASELINE configs 1 and 3 with the on-device generators: matches = |S|, keysum and checksum
    have closed forms because every block of |R| S-tuples is a permutation of the R keys and
    payload = row id."""
    import torch
    nR, nS = 1 << logR, 1 << logS
    dev = torch.device("cuda:0")
    R = torch.empty(nR * 2, dtype=torch.int32, device=dev)
    S = torch.empty(nS * 2, dtype=torch.int32, device=dev)
    gpu.gen_pk_device(R.data_ptr(), nR, seed=11111)
    gpu.gen_fk_device(S.data_ptr(), nS, nR, seed=22222)
    torch.cuda.synchronize()
    s = gpu.join_device(R.data_ptr(), nR, S.data_ptr(), nS)
    rep = nS // nR
    assert s["matches"] == nS
    assert s["keysum"] == rep * nR * (nR + 1) // 2
    assert s["checksum"] == rep * (nR * (nR - 1) // 2) + nS * (nS - 1) // 2
    s2 = gpu.join_device(R.data_ptr(), nR, S.data_ptr(), nS)   # idempotent, inputs untouched
    assert (s2["matches"], s2["checksum"], s2["keysum"]) == (s["matches"], s["checksum"], s["keysum"])


@pytest.mark.parametrize("z", [0.5, 1.0])
def test_full_size_zipf(gpu, z):
    """BASELINE config 4: |R|=2^27, Zipf-skewed |S|=2^29 generated on the device. R is a primary key, so
    every S tuple matches exactly once; checksum and keysum are re-derived with torch from the relations."""
    import torch
    nR, nS = 1 << 27, 1 << 29
    dev = torch.device("cuda:0")
    R = torch.empty(nR * 2, dtype=torch.int32, device=dev)
    S = torch.empty(nS * 2, dtype=torch.int32, device=dev)
    gpu.gen_pk_device(R.data_ptr(), nR, seed=11111)
    gpu.gen_zipf_device(S.data_ptr(), nS, nR, z, seed=22222)
    gpu.lib().b200_device_sync()
    s = gpu.join_device(R.data_ptr(), nR, S.data_ptr(), nS)
    assert s["matches"] == nS
    Rk, Rp = R.view(nR, 2)[:, 0].long(), R.view(nR, 2)[:, 1].long()
    inv = torch.empty(nR + 1, dtype=torch.int64, device=dev)
    inv[Rk] = Rp
    Sk, Sp = S.view(nS, 2)[:, 0].long(), S.view(nS, 2)[:, 1].long()
    assert int(Sk.min()) >= 1 and int(Sk.max()) <= nR
    assert s["keysum"] == int(Sk.sum())
    assert s["checksum"] == int(inv[Sk].sum() + Sp.sum())
    hot = int(torch.bincount(Sk[: 1 << 24].int()).max())
    assert hot > (1 << 24) * (0.03 if z == 1.0 else 1e-5)      # the skew is really there


STAGED_SCRIPT = r'''
import os, sys
sys.path.insert(0, os.environ["AQP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["AQP_ROOT"], "sgxv2-analytical-query-processing-benchmarks_b200"))
import numpy as np
import b200aqp as A, oracle as O
A.init(0)
for nR, nS, zipf in ((300007, 1000003, 0.0), (1 << 16, 1 << 20, 1.0), (5000, 70001, 0.0)):
    R = O.set_rowid_payload(O.gen_pk(nR, 11111))
    S = O.set_rowid_payload(O.gen_zipf(nS, nR, zipf, 22222) if zipf else O.gen_fk(nS, nR, 22222))
    exp = O.rho(R, S, nthreads=1)
    got = A.run_join(R, S)
    assert (got["matches"], got["checksum"], got["keysum"]) == (exp["matches"], exp["checksum"], exp["keysum"]), (nR, nS, got)
print("STAGED OK", os.environ.get("B200_AQP_SCATTER"), os.environ.get("B200_AQP_SCATTER_BULK"))
'''


@pytest.mark.parametrize("bulk", ["1", "0"])
def test_staged_scatter_kernel_still_correct(tmp_path, bulk):
    """radix_scatter_kernel (scan + look-ups; the path for destinations that are not 16-byte aligned) is selected with
    an environment switch that the library reads once per process, so it runs in a child process: with and without
    its TMA bulk-store write-out, uniform and Zipf-skewed."""
    import subprocess
    import sys
    from conftest import ROOT
    script = tmp_path / "staged.py"
    script.write_text(STAGED_SCRIPT)
    env = dict(os.environ, AQP_ROOT=ROOT, B200_AQP_SCATTER="staged", B200_AQP_SCATTER_BULK=bulk)
    p = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "STAGED OK staged" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
