"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file <csv>` launch list per kernel.
usage: python tools/ncu_launches.py <launches.csv> "<title>" > profiles/<name>.md"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(r for r in rows if "Kernel Name" in r)
body = rows[rows.index(hdr) + 1:]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in body:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = r[ik].split("(")[0].strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {n} | {us:.1f} | {us / n:.1f} | {100 * us / tot:.1f} % |")
