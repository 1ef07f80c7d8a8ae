"""Tuning helper (not part of the product): time the device-resident join phases for one build of the library.
usage: B200_AQP_LIB=<path/to/lib.so> python tools/sweep_join.py [logR logS] [reps] [zipf exponent of S]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200"))
import torch
import b200aqp as A

logR = int(sys.argv[1]) if len(sys.argv) > 1 else 27
logS = int(sys.argv[2]) if len(sys.argv) > 2 else 29
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
A.init(0)
nR, nS = 1 << logR, 1 << logS
dev = torch.device("cuda:0")
R = torch.empty(nR * 2, dtype=torch.int32, device=dev)
S = torch.empty(nS * 2, dtype=torch.int32, device=dev)
A.gen_pk_device(R.data_ptr(), nR, 11111)
zipf = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
if zipf > 0:
    A.gen_zipf_device(S.data_ptr(), nS, nR, zipf, 22222)
else:
    A.gen_fk_device(S.data_ptr(), nS, nR, 22222)
A.lib().b200_device_sync()
best = None
for i in range(reps + 2):
    s = A.join_device(R.data_ptr(), nR, S.data_ptr(), nS)
    assert s["matches"] == nS or zipf
    if i >= 2 and (best is None or s["ms_total"] < best["ms_total"]):
        best = s
print(os.path.basename(A.LIB_PATH), f"2^{logR}x2^{logS}" + (f" zipf {zipf}" if zipf else ""),
      " ".join(f"{k}={best[k]:.3f}" for k in ("ms_total", "ms_hist", "ms_pass1", "ms_pass2", "ms_join")),
      f"Gtuples/s={(nR + nS) / best['ms_total'] / 1e6:.1f}")
