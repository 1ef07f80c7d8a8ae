"""Instruction histogram of the hot kernels of libb200aqp.so (cuobjdump -sass), the evidence that the path uses the
Blackwell copy / barrier hardware it claims: UBLKCP = cp.async.bulk (TMA, non-tensor), SYNCS = mbarrier, LDG.E.*.256 =
256-bit global loads, IDP.4A = dp4a, REDUX = warp reductions, ATOMS / RED = shared / global atomics.
usage: python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200", "libb200aqp.so")
HOT = ["radix_hist_smem_kernel", "radix_scatter_bins_kernel", "radix_scatter_peer_kernel", "build_probe_kernel",
       "bitvector_scan_kernel", "rowid_scan_fused_kernel<0, 16, 16, true>", "scan_count_kernel", "filter_compact_kernel",
       "filter_compact_bytes_kernel", "region_sample_kernel"]
COLS = ["total", "UBLKCP", "SYNCS", "LDG.256", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG/RED", "REDUX", "IDP", "LOP3", "IMAD", "SHFL",
        "BAR"]


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        kernels[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = kernels[cur]
        c["total"] += 1
        base = op.split(".")[0]
        if base == "UBLKCP": c["UBLKCP"] += 1
        elif base == "SYNCS": c["SYNCS"] += 1
        elif base == "LDG":
            c["LDG.256" if ".256" in op else "LDG"] += 1
        elif base in ("STG", "LDS", "STS", "ATOMS", "REDUX", "IDP", "LOP3", "IMAD", "SHFL", "BAR"): c[base] += 1
        elif base in ("ATOMG", "RED", "ATOM"): c["ATOMG/RED"] += 1

print("# r02 SASS summary of the hot kernels (cuobjdump -sass libb200aqp.so, sm_100a)\n")
print("Static instruction counts per kernel instantiation (not executed counts). UBLKCP = `cp.async.bulk` (TMA bulk copy,")
print("global<->shared), SYNCS = mbarrier arrive / try_wait, LDG.256 = `ld.global.v8.u32`, IDP = `dp4a`.\n")
print("| kernel | " + " | ".join(COLS) + " |")
print("|---|" + "---|" * len(COLS))
for name in sorted(kernels):
    short = re.sub(r"^void aqp::", "", name)
    short = re.sub(r"\(.*$", "", short)
    if not any(h.split("<")[0] in short for h in HOT):
        continue
    if "rowid_scan_fused_kernel" in short and "16, 16, true" not in short.replace("(int)", "").replace("(bool)", ""):
        continue
    c = kernels[name]
    print(f"| `{short}` | " + " | ".join(str(c[k]) for k in COLS) + " |")
