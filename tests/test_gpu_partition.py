"""GPU parity of the partitioning stages (histogram, prefix sum, scatter) through the C ABI vs the
oracle's restatement of (key & MASK) >> R partitioning: identical histograms and offsets, and every
partition holds the same multiset of tuples (order inside a partition is unspecified)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(gpu, rel, shift, bits):
    n = len(rel)
    fan = 1 << bits
    d_in = gpu.to_device(rel)
    d_hist = gpu.to_device(np.zeros(fan, dtype=np.uint32))
    d_off = gpu.DeviceBuffer(4 * (fan + 1))
    d_cur = gpu.DeviceBuffer(4 * fan)
    d_out = gpu.DeviceBuffer(max(8 * n, 16))
    L = gpu.lib()
    assert L.b200_radix_hist_device(d_in.ptr, n, shift, bits, d_hist.ptr, None) == 0
    assert L.b200_exclusive_scan_u32_device(d_hist.ptr, fan, d_off.ptr, None) == 0
    assert L.b200_radix_scatter_device(d_in.ptr, n, shift, bits, d_off.ptr, d_cur.ptr, d_out.ptr, None) == 0
    assert L.b200_device_sync() == 0
    return d_hist.download(np.uint32, fan), d_off.download(np.uint32, fan + 1), d_out.download(gpu.ROW, n)


@pytest.mark.parametrize("n,shift,bits", [(1, 0, 4), (4095, 0, 7), (4096, 0, 7), (4097, 3, 5), (100003, 0, 8),
                                          (1 << 20, 7, 7), (1 << 20, 0, 1), (3_000_017, 5, 6), (50000, 0, 0)])
def test_partition_stages_vs_oracle(gpu, oracle, n, shift, bits):
    rel = oracle.set_rowid_payload(oracle.gen_pk(n, 11111))
    hist, off, out = _run(gpu, rel, shift, bits)
    exp_out, exp_off = oracle.radix_partition(rel, shift, bits)
    assert np.array_equal(off.astype(np.uint64), exp_off)
    assert np.array_equal(hist.astype(np.uint64), np.diff(exp_off))
    for p in range(1 << bits):
        a, b = int(exp_off[p]), int(exp_off[p + 1])
        got = np.sort(out[a:b], order=["key", "payload"])
        exp = np.sort(exp_out[a:b], order=["key", "payload"])
        assert np.array_equal(got, exp), p


def test_wide_histograms(gpu, oracle):
    rel = oracle.gen_pk(1 << 20, 3)
    for bits in (11, 14, 15, 16):     # 16 > shared-memory histogram limit: global-atomic fallback
        d_in = gpu.to_device(rel)
        d_hist = gpu.to_device(np.zeros(1 << bits, dtype=np.uint32))
        assert gpu.lib().b200_radix_hist_device(d_in.ptr, len(rel), 0, bits, d_hist.ptr, None) == 0
        assert gpu.lib().b200_device_sync() == 0
        hist = d_hist.download(np.uint32, 1 << bits)
        assert np.array_equal(hist, np.bincount(rel["key"] & ((1 << bits) - 1), minlength=1 << bits))


def test_skewed_digits(gpu, oracle):
    rel = np.zeros(500000, dtype=oracle.ROW)
    rel["key"] = np.where(np.random.default_rng(1).random(500000) < 0.9, 77, np.arange(500000)).astype(np.uint32)
    oracle.set_rowid_payload(rel)
    hist, off, out = _run(gpu, rel, 0, 8)
    exp_out, exp_off = oracle.radix_partition(rel, 0, 8)
    assert np.array_equal(off.astype(np.uint64), exp_off)
    assert np.array_equal(np.sort(out, order=["key", "payload"]), np.sort(exp_out, order=["key", "payload"]))
    assert ((out["key"] & 255) == np.repeat(np.arange(256), np.diff(exp_off).astype(np.int64))).all()
