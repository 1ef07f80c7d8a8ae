"""On-device relation generators: same key distribution as the reference generators
(generator.cpp:352,:474) — PK a permutation of 1..n, FK blocks of permutations — and row-range
generation (multi-GPU shards) consistent with whole-relation generation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gen_pk(gpu, n, seed, begin=0, cnt=None):
    cnt = n if cnt is None else cnt
    buf = gpu.DeviceBuffer(max(8 * cnt, 16))
    gpu.gen_pk_device(buf.ptr, n, seed, begin, cnt)
    gpu.lib().b200_device_sync()
    return buf.download(gpu.ROW, cnt)


def _gen_fk(gpu, n, maxid, seed, begin=0, cnt=None):
    cnt = n if cnt is None else cnt
    buf = gpu.DeviceBuffer(max(8 * cnt, 16))
    gpu.gen_fk_device(buf.ptr, n, maxid, seed, begin, cnt)
    gpu.lib().b200_device_sync()
    return buf.download(gpu.ROW, cnt)


@pytest.mark.parametrize("n", [1, 2, 3, 1000, 4096, 100003, 1 << 20])
def test_pk_is_permutation(gpu, n):
    R = _gen_pk(gpu, n, 11111)
    assert np.array_equal(np.sort(R["key"]), np.arange(1, n + 1, dtype=np.uint32))
    assert np.array_equal(R["payload"], np.arange(n, dtype=np.uint32))
    if n >= 1000:
        assert not np.array_equal(R["key"], np.arange(1, n + 1, dtype=np.uint32))
        assert not np.array_equal(R["key"], _gen_pk(gpu, n, 22222)["key"])     # seed matters
    if n >= 100000:
        # low radix bits of the first half look uniform (what partitioning sees)
        h = np.bincount(R["key"][: n // 2] & 15, minlength=16)
        assert h.min() > 0.9 * (n // 2) / 16 and h.max() < 1.1 * (n // 2) / 16


def test_fk_blocks_are_permutations(gpu):
    maxid, n = 5000, 5000 * 3 + 77
    S = _gen_fk(gpu, n, maxid, 22222)
    for b in range(3):
        assert np.array_equal(np.sort(S["key"][b * maxid:(b + 1) * maxid]), np.arange(1, maxid + 1, dtype=np.uint32))
    assert np.array_equal(np.sort(S["key"][3 * maxid:]), np.arange(1, 78, dtype=np.uint32))   # generator.cpp:499-504
    assert not np.array_equal(S["key"][:maxid], S["key"][maxid:2 * maxid])
    assert np.array_equal(S["payload"], np.arange(n, dtype=np.uint32))


def test_row_ranges_compose(gpu):
    n = 100003
    whole = _gen_pk(gpu, n, 5)
    parts = [_gen_pk(gpu, n, 5, b, min(25000, n - b)) for b in range(0, n, 25000)]
    assert np.array_equal(np.concatenate(parts), whole)
    whole = _gen_fk(gpu, n, 30011, 6)
    parts = [_gen_fk(gpu, n, 30011, 6, b, min(40000, n - b)) for b in range(0, n, 40000)]
    assert np.array_equal(np.concatenate(parts), whole)


def test_zipf_device_distribution_and_join(gpu, oracle):
    maxid, n = 1 << 12, 1 << 20
    for z in (0.5, 1.0):
        buf = gpu.DeviceBuffer(8 * n)
        gpu.gen_zipf_device(buf.ptr, n, maxid, z, seed=7)
        gpu.lib().b200_device_sync()
        S = buf.download(gpu.ROW, n)
        assert S["key"].min() >= 1 and S["key"].max() <= maxid
        assert np.array_equal(S["payload"], np.arange(n, dtype=np.uint32))
        # frequencies follow rank^-z: compare the sorted empirical frequencies with the reference's table
        lut = np.empty(maxid)
        oracle.lib().oracle_zipf_lut(lut.ctypes.data, maxid, z)
        p = np.diff(np.concatenate([[0.0], lut]))
        freq = np.sort(np.bincount(S["key"], minlength=maxid + 1)[1:])[::-1] / n
        assert abs(freq[0] - p[0]) < 0.1 * p[0] + 3e-4            # tolerance: sampling noise of n = 2^20 draws
        assert abs(freq[:16].sum() - p[:16].sum()) < 0.05 * p[:16].sum() + 1e-3
        # row ranges compose and the relation joins like the oracle says
        half = gpu.DeviceBuffer(8 * (n // 2))
        gpu.gen_zipf_device(half.ptr, n // 2, maxid, z, seed=7, row_begin=n // 2)
        gpu.lib().b200_device_sync()
        assert np.array_equal(half.download(gpu.ROW, n // 2), S[n // 2:])
        R = oracle.set_rowid_payload(oracle.gen_pk(maxid, 11111))
        o = oracle.rho(R, S)
        g = gpu.run_join(R, S)
        assert (g["matches"], g["checksum"], g["keysum"]) == (o["matches"], o["checksum"], o["keysum"])
        assert g["matches"] == n
