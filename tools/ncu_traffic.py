"""Per-kernel DRAM traffic from an `ncu --set full` report -> JSON (bench.py reads profiles/traffic.json for
roofline.traffic). usage: python tools/ncu_traffic.py <file.ncu-rep> [more.ncu-rep ...] > profiles/traffic.json
For every kernel name: launches captured, mean dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes),
mean gpu__time_duration (ns, under ncu: cold-cache, serialised)."""
import csv, io, json, subprocess, sys

out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    def to_ns(v, u):
        return float(v) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ns": 0.0, "per_launch": [], "source": []})
        b = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
            to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        e["launches"] += 1
        e["dram_bytes"] += b
        e["ns"] += to_ns(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        e["per_launch"].append(round(b))
        if rep not in e["source"]:
            e["source"].append(rep)
for e in out.values():
    e["dram_bytes_per_launch"] = e.pop("dram_bytes") / e["launches"]
    e["ns_per_launch"] = e.pop("ns") / e["launches"]
print(json.dumps(out, indent=1))
