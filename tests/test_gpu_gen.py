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
