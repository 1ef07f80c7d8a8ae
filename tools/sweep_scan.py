"""Tuning helper (not part of the product): time the device-resident scans for one build of the library.
usage: B200_AQP_LIB=<lib.so> python tools/sweep_scan.py [log2 n] [reps] [geom,geom,...]
Row-id scans are timed for every geometry of the single-pass kernel named on the command line (B200_AQP_SCAN_GEOM,
read per call); run with B200_AQP_INDEX_SCAN=twopass for the two-pass design. Two columns: the reference's tiled
column (0.39 / 10.5 / 50.4 / 100 %) and the seeded skewed column (0.1 %, predicate [0,0])."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200"))
import torch
import b200aqp as A

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 30)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
geoms = sys.argv[3].split(",") if len(sys.argv) > 3 else ["default"]
A.init(0)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
data = torch.empty(n, dtype=torch.uint8, device=dev)
skew = torch.empty(n, dtype=torch.uint8, device=dev)
assert A.lib().b200_fill_tiled_column_device(data.data_ptr(), n, 0, A._st(st)) == 0
assert A.lib().b200_fill_skewed_column_device(skew.data_ptr(), n, 0, 1000, 42, A._st(st)) == 0
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
print(f"{os.path.basename(A.LIB_PATH)} n=2^{n.bit_length() - 1} bitvector {ms:.4f} ms frac={1.125 * n / ms / 1e6 / peak:.3f}", flush=True)
k_skew = int((skew == 0).sum().item())
cases = [("skew0.1%", skew, 0, 0, k_skew)] + [(f"tiled[0,{hi}]", data, 0, hi, n // 256 * (hi + 1)) for hi in (0, 26, 128, 255)]
for geom in geoms:
    if geom != "default":
        os.environ["B200_AQP_SCAN_GEOM"] = geom
    line = []
    for name, col, lo, hi, k in cases:
        ids = torch.empty(k, dtype=torch.int64, device=dev)
        ms = timeit(lambda: A.index_scan_device(lo, hi, col.data_ptr(), n, ids.data_ptr(), k, cnt.data_ptr(), stream=st))
        assert int(cnt.item()) == k, (name, int(cnt.item()), k)
        # ascending and exactly the matching positions
        if col is data:
            v = ids.view(n // 256, hi + 1)
            ok = torch.equal(v[:, 0], torch.arange(0, n, 256, device=dev)) and \
                torch.equal(v - v[:, :1], torch.arange(hi + 1, device=dev).expand(n // 256, hi + 1))
        else:
            ok = torch.equal(ids, torch.nonzero(col == 0).view(-1))
        line.append(f"{name} {ms:.4f} ms frac={(n + 8 * k) / ms / 1e6 / peak:.3f}{'' if ok else ' WRONG'}")
        del ids
    print(f"geom={geom} " + " | ".join(line), flush=True)
