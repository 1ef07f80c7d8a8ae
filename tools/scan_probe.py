"""Tuning helper: run the device-resident scans once per predicate (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200"))
import torch
import b200aqp as A

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
A.init(0)
dev = torch.device("cuda:0")
data = torch.empty(n, dtype=torch.uint8, device=dev)
assert A.lib().b200_fill_tiled_column_device(data.data_ptr(), n, 0, None) == 0
bv = torch.empty(n // 64, dtype=torch.int64, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
ids = torch.empty(n, dtype=torch.int64, device=dev)
for _ in range(reps):
    A.bitvector_scan_device(0, 26, data.data_ptr(), n, bv.data_ptr())
    for hi in (0, 26, 128, 255):
        A.index_scan_device(0, hi, data.data_ptr(), n, ids.data_ptr(), n, cnt.data_ptr())
        A.lib().b200_device_sync()
        assert int(cnt.item()) == n // 256 * (hi + 1)
print("ok")
