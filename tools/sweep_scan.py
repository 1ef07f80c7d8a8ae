"""Tuning helper (not part of the product): time the device-resident scans for one build of the library.
usage: B200_AQP_LIB=<lib.so> python tools/sweep_scan.py [log2 n] [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200"))
import torch
import b200aqp as A

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 30)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
A.init(0)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
data = torch.empty(n, dtype=torch.uint8, device=dev)
assert A.lib().b200_fill_tiled_column_device(data.data_ptr(), n, 0, A._st(st)) == 0
bv = torch.empty(n // 64, dtype=torch.int64, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ms = timeit(lambda: A.bitvector_scan_device(0, 26, data.data_ptr(), n, bv.data_ptr(), st))
line = [f"bitvector {ms:.4f} ms frac={1.125 * n / ms / 1e6 / peak:.3f}"]
for hi in (0, 26, 128, 255):
    k = n // 256 * (hi + 1)
    ids = torch.empty(k, dtype=torch.int64, device=dev)
    ms = timeit(lambda: A.index_scan_device(0, hi, data.data_ptr(), n, ids.data_ptr(), k, cnt.data_ptr(), stream=st))
    assert int(cnt.item()) == k
    line.append(f"rowid[0,{hi}] {ms:.4f} ms frac={(n + 8 * k) / ms / 1e6 / peak:.3f}")
    del ids
print(os.path.basename(A.LIB_PATH), " | ".join(line))
