"""Tuning helper: run the device-resident scans once per case (for ncu launch lists / captures).
usage: python tools/scan_probe.py [log2 n] [reps] [cases]   cases: comma list of skew,0,26,128,255,bv (default all)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200"))
import torch
import b200aqp as A

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cases = sys.argv[3].split(",") if len(sys.argv) > 3 else ["bv", "skew", "0", "26", "128", "255"]
A.init(0)
dev = torch.device("cuda:0")
data = torch.empty(n, dtype=torch.uint8, device=dev)
assert A.lib().b200_fill_tiled_column_device(data.data_ptr(), n, 0, None) == 0
skew = None
if "skew" in cases:
    skew = torch.empty(n, dtype=torch.uint8, device=dev)
    assert A.lib().b200_fill_skewed_column_device(skew.data_ptr(), n, 0, 1000, 42, None) == 0
bv = torch.empty(n // 64, dtype=torch.int64, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
ids = torch.empty(n if "255" in cases else n // 2 + n // 64, dtype=torch.int64, device=dev)
A.lib().b200_device_sync()
for _ in range(reps):
    for c in cases:
        if c == "bv":
            A.bitvector_scan_device(0, 26, data.data_ptr(), n, bv.data_ptr())
        elif c == "skew":
            A.index_scan_device(0, 0, skew.data_ptr(), n, ids.data_ptr(), ids.numel(), cnt.data_ptr())
        else:
            hi = int(c)
            A.index_scan_device(0, hi, data.data_ptr(), n, ids.data_ptr(), ids.numel(), cnt.data_ptr())
            A.lib().b200_device_sync()
            assert int(cnt.item()) == n // 256 * (hi + 1)
    A.lib().b200_device_sync()
print("ok")
