"""GPU parity: scans through the C ABI vs the CPU oracle (bit-exact), the committed reference
fixtures, and size-independent properties at the full 2^30 benchmark size."""
import numpy as np
import pytest

from helpers import sha

pytestmark = pytest.mark.gpu

PREDS = [(0, 0), (0, 26), (0, 128), (0, 255), (5, 5), (17, 200), (100, 50), (255, 255), (128, 255), (1, 254), (127, 128)]


def _column(oracle, name, n):
    if name == "tiled":
        return oracle.tiled_column(n)
    return np.random.default_rng(7).integers(0, 256, n, dtype=np.uint8)


def test_scan_matches_golden(gpu, oracle, golden):
    cols = {}
    for c in golden["scan"]:
        col = cols.setdefault(c["column"], _column(oracle, c["column"], c["n"]))
        bv, _ = gpu.bitvector_scan_user(c["lo"], c["hi"], col)
        ids, cnt, _ = gpu.index_scan_user(c["lo"], c["hi"], col)
        assert cnt == c["count"], c
        assert sha(bv) == c["sha256_bitvector"], c
        assert sha(ids) == c["sha256_rowids"], c


@pytest.mark.parametrize("n", [0, 64, 128, 64 * 3 + 17, 16384, 16384 - 64, 16384 + 64, 16384 * 5 + 64 * 7 + 13,
                               (1 << 20) + 64, 3_000_000])
def test_scan_vs_oracle_sizes(gpu, oracle, n):
    rng = np.random.default_rng(n)
    col = rng.integers(0, 256, n, dtype=np.uint8)
    for lo, hi in PREDS:
        bv, _ = gpu.bitvector_scan_user(lo, hi, col)
        assert np.array_equal(bv, oracle.bitvector_scan(lo, hi, col)), (n, lo, hi)
        ids, cnt, _ = gpu.index_scan_user(lo, hi, col)
        exp = oracle.index_scan(lo, hi, col)
        assert cnt == len(exp) == oracle.scan_count(lo, hi, col)
        assert np.array_equal(ids, exp), (n, lo, hi)


def test_all_256_equality_predicates(gpu, oracle):
    col = np.random.default_rng(3).integers(0, 256, 1 << 16, dtype=np.uint8)
    for v in range(256):
        bv, _ = gpu.bitvector_scan_user(v, v, col)
        assert np.array_equal(bv, oracle.bitvector_scan(v, v, col)), v


def test_random_ranges_and_sparse_columns(gpu, oracle):
    rng = np.random.default_rng(11)
    n = 1 << 18
    sparse = np.where(rng.random(n) < 0.001, 0, rng.integers(1, 256, n)).astype(np.uint8)   # ~0.1 % selectivity
    dense = rng.integers(0, 256, n, dtype=np.uint8)
    for col in (sparse, dense):
        for _ in range(12):
            lo, hi = (int(x) for x in rng.integers(0, 256, 2))
            bv, _ = gpu.bitvector_scan_user(lo, hi, col)
            assert np.array_equal(bv, oracle.bitvector_scan(lo, hi, col)), (lo, hi)
            ids, cnt, _ = gpu.index_scan_user(lo, hi, col)
            assert np.array_equal(ids, oracle.index_scan(lo, hi, col)), (lo, hi)
    ids, cnt, _ = gpu.index_scan_user(0, 0, sparse)
    assert 100 < cnt < 500


def test_mixed_density_tiles(gpu, oracle):
    """Row-id expansion picks a kernel per 65536-value tile (empty: none, < 3/4 full: shared-memory window,
    otherwise whole-warp): one column with every kind side by side, in an order that makes the lists ragged,
    plus a capacity that cuts through a dense tile."""
    rng = np.random.default_rng(23)
    T = 1 << 16
    kinds = [0.0, 1.0, 0.8, 0.3, 0.0, 1.0, 1e-5, 0.74, 0.76, 0.5, 1.0, 0.0, 0.02, 0.999, 0.25, 1.0, 0.6]
    parts = []
    for dens in kinds:
        hit = rng.random(T) < dens
        parts.append(np.where(hit, rng.integers(10, 20, T), rng.integers(100, 256, T)).astype(np.uint8))
    col = np.concatenate(parts + [parts[2][:64 * 37]])           # ragged tail tile
    exp = oracle.index_scan(10, 19, col)
    ids, cnt, _ = gpu.index_scan_user(10, 19, col)
    assert cnt == len(exp) and np.array_equal(ids, exp)
    bv, _ = gpu.bitvector_scan_user(10, 19, col)
    assert np.array_equal(bv, oracle.bitvector_scan(10, 19, col))
    cut = int(np.searchsorted(exp, 5 * T + 1000))                  # inside the second full tile
    ids, cnt, _ = gpu.index_scan_user(10, 19, col, capacity=cut)
    assert cnt == len(exp) and np.array_equal(ids, exp[:cut])


def test_index_scan_capacity_clamp(gpu, oracle):
    col = oracle.tiled_column(1 << 16)
    ids, cnt, _ = gpu.index_scan_user(0, 127, col, capacity=1000)
    assert cnt == (1 << 15) and len(ids) == 1000
    assert np.array_equal(ids, oracle.index_scan(0, 127, col)[:1000])


def test_timed_runs_accumulate(gpu, oracle):
    col = oracle.tiled_column(1 << 22)
    _, t1 = gpu.bitvector_scan_user(0, 26, col, num_runs=1, warmup_runs=1, unique_data=False)
    _, t8 = gpu.bitvector_scan_user(0, 26, col, num_runs=8, warmup_runs=1, unique_data=False)
    _, tu = gpu.bitvector_scan_user(0, 26, col, num_runs=8, warmup_runs=3, unique_data=True)   # forced to 1 run
    assert t1 > 0 and t8 > 2 * t1 and tu < t8


def test_full_size_properties(gpu):
    """2^30 values (BASELINE config 2) on the device-resident API; verified with closed forms on the
    tiled column: count = n/256*(hi+1), bitvector period = 4 words, row ids = {i : i mod 256 <= hi}."""
    import torch
    n = 1 << 30
    dev = torch.device("cuda:0")
    data = torch.empty(n, dtype=torch.uint8, device=dev)
    assert gpu.lib().b200_fill_tiled_column_device(data.data_ptr(), n, 0, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(data[:512].cpu(), (torch.arange(512) % 256).to(torch.uint8))
    bv = torch.empty(n // 64, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    for hi in (0, 26, 128, 255):
        gpu.bitvector_scan_device(0, hi, data.data_ptr(), n, bv.data_ptr())
        gpu.scan_count_device(0, hi, data.data_ptr(), n, cnt.data_ptr())
        torch.cuda.synchronize()
        exp_count = n // 256 * (hi + 1)
        assert int(cnt.item()) == exp_count
        words = [sum(1 << k for k in range(64) if (64 * w + k) <= hi) for w in range(4)]
        words = [x - (1 << 64) if x >= (1 << 63) else x for x in words]
        exp = torch.tensor(words, dtype=torch.int64, device=dev)
        assert torch.equal(bv.view(-1, 4), exp.expand(n // 256, 4))
        ids = torch.empty(exp_count, dtype=torch.int64, device=dev)
        gpu.index_scan_device(0, hi, data.data_ptr(), n, ids.data_ptr(), exp_count, cnt.data_ptr())
        torch.cuda.synchronize()
        assert int(cnt.item()) == exp_count
        v = ids.view(n // 256, hi + 1)
        assert torch.equal(v[:, 0], torch.arange(0, n, 256, device=dev))            # ascending, complete
        assert torch.equal(v - v[:, :1], torch.arange(hi + 1, device=dev).expand(n // 256, hi + 1))
        del ids, v
    # id_base shifts every id (multi-GPU row-range shards)
    ids = torch.empty(n // 256, dtype=torch.int64, device=dev)
    gpu.index_scan_device(0, 0, data.data_ptr(), n, ids.data_ptr(), n // 256, cnt.data_ptr(), id_base=1 << 40)
    torch.cuda.synchronize()
    assert torch.equal(ids, torch.arange(0, n, 256, device=dev) + (1 << 40))


def test_dict_scan_reference_known_answers_on_gpu(gpu, oracle):
    """The reference's own [main] dictionary-scan cases (testsimdscan.cpp:8-165,:217-245) through the C ABI."""
    from test_oracle_scan import dict_kats
    for name, col, d, lo, hi, size, spots in dict_kats(oracle):
        r, cnt = gpu.dict_scan_8bit_64bit(lo, hi, d, col)
        assert cnt == size == len(r), name
        for i, v in spots.items():
            assert r[i] == v, (name, i)
        assert np.array_equal(r, oracle.dict_scan_8_64(lo, hi, d, col)), name


@pytest.mark.parametrize("n", [0, 64, 64 * 3 + 17, 65536 + 64, 3_000_000])
def test_sum_value_dict_vs_oracle(gpu, oracle, n):
    rng = np.random.default_rng(n + 1)
    col = rng.integers(0, 256, n, dtype=np.uint8)
    for lo, hi in PREDS:
        assert gpu.scan_sum(lo, hi, col) == oracle.scan_sum(lo, hi, col), (n, lo, hi)
        vals, cnt = gpu.value_scan(lo, hi, col)
        exp = oracle.value_scan(lo, hi, col)
        assert cnt == len(exp) and np.array_equal(vals, exp), (n, lo, hi)
    d = np.sort(rng.integers(-10**12, 10**12, 256))
    for lo, hi in [(int(d[10]), int(d[90])), (int(d[0]) - 5, int(d[0]) - 1), (int(d[255]) + 1, int(d[255]) + 9),
                   (int(d[200]), int(d[100])), (int(d[17]), int(d[17])), (int(d[0]), int(d[255]))]:
        r, cnt = gpu.dict_scan_8bit_64bit(lo, hi, d, col)
        exp = oracle.dict_scan_8_64(lo, hi, d, col)
        assert cnt == len(exp) and np.array_equal(r, exp), (n, lo, hi)
    # dense tiles (whole-warp expansion kernel) and a capacity that cuts the output
    if n >= 65536:
        dense = np.where(rng.random(n) < 0.9, 7, 200).astype(np.uint8)
        vals, cnt = gpu.value_scan(0, 100, dense, capacity=1000)
        exp = oracle.value_scan(0, 100, dense)
        assert cnt == len(exp) and np.array_equal(vals, exp[:1000])
        r, cnt = gpu.dict_scan_8bit_64bit(0, 100, np.arange(256), dense)
        assert cnt == len(exp) and np.array_equal(r, exp.astype(np.int64))


def test_wide_dict_scan_reference_known_answers_on_gpu(gpu, oracle):
    """testsimdscan.cpp:167-215,:478-528 through the C ABI"""
    from test_oracle_scan import wide_dict_kats
    for name, bits, col, d, lo, hi, size, check in wide_dict_kats():
        r, cnt = gpu.dict_scan_wide(bits, lo, hi, d, col)
        assert cnt == size == len(r), name
        assert check is None or check(r), name
        assert np.array_equal(r, oracle.dict_scan_wide(bits, lo, hi, d, col)), name


@pytest.mark.parametrize("n", [0, 31, 64 * 7 + 5, 4096 + 32, 1_000_003])
def test_wide_dict_explicit_scalar_vs_oracle(gpu, oracle, n):
    rng = np.random.default_rng(n + 3)
    col = rng.integers(0, 256, n, dtype=np.uint8)
    index = rng.integers(0, 1 << 62, (n // 64 + 7) * 8, dtype=np.uint64)
    for lo, hi in PREDS:
        ids, cnt = gpu.explicit_index_scan(lo, hi, index, col)
        exp = oracle.explicit_index_scan(lo, hi, index, col)
        assert cnt == len(exp) and np.array_equal(ids, exp), (n, lo, hi)
        ids, cnt = gpu.scalar_index_scan(lo, hi, col)            # all n values, incl. the < 64 behind the last block
        exp = oracle.scalar_index_scan(lo, hi, col)
        assert cnt == len(exp) and np.array_equal(ids, exp), (n, lo, hi)
    d16 = np.sort(rng.integers(-10**9, 10**9, 1 << 16))
    c16 = rng.integers(0, 1 << 16, n, dtype=np.uint16)
    d32 = np.sort(rng.integers(-10**9, 10**9, 70000))
    c32 = rng.integers(0, 70000, n, dtype=np.uint32)
    for bits, d, c in ((16, d16, c16), (32, d32, c32)):
        for lo, hi in [(int(d[50]), int(d[5000])), (int(d[0]) - 2, int(d[0]) - 1), (int(d[-1]) + 1, int(d[-1]) + 2),
                       (int(d[9]), int(d[9])), (int(d[600]), int(d[500]))]:
            r, cnt = gpu.dict_scan_wide(bits, lo, hi, d, c)
            exp = oracle.dict_scan_wide(bits, lo, hi, d, c)
            assert cnt == len(exp) and np.array_equal(r, exp), (bits, n, lo, hi)
    if n > 4096:
        r, cnt = gpu.dict_scan_wide(16, int(d16[0]), int(d16[-1]), d16, c16, capacity=100)
        assert cnt == n // 32 * 32 and np.array_equal(r, d16[c16[:100]])
