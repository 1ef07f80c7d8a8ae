"""Per-region hot spots of one captured launch: buckets the SASS of the source page (needs -lineinfo / --import-source)
into basic-block-ish runs and prints executed warp instructions and stall samples per run.
usage: python tools/ncu_hot.py <file.ncu-rep> [launch index, default 1] [min share %, default 2]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
k = int(kid) - 1
rows = allrows[starts[k]:starts[k + 1]]
print(rows[0][1][:120])
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_ex = sum(int(r[iex] or 0) for r in body)
tot_smp = sum(int(r[ismp] or 0) for r in body)
print(f"total warp instructions {tot_ex}, samples {tot_smp}")
run = []
def flush():
    if not run:
        return
    ex = sum(int(r[iex] or 0) for r in run)
    smp = sum(int(r[ismp] or 0) for r in run)
    if 100.0 * ex / max(tot_ex, 1) >= minshare or 100.0 * smp / max(tot_smp, 1) >= minshare:
        st = {}
        for i, h in stall_cols:
            st[h] = sum(int(r[i] or 0) for r in run)
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        ops = {}
        for r in run:
            op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        topops = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
        print(f"{run[0][ia]}..{run[-1][ia]} n={len(run):4d} exec {100.0 * ex / tot_ex:5.1f}% samples {100.0 * smp / max(tot_smp, 1):5.1f}%  "
              + " ".join(f"{k[6:]}={v}" for k, v in top if v) + "  | " + " ".join(f"{k}:{v}" for k, v in topops))
    run.clear()
for r in body:
    run.append(r)
    op = r[isrc]
    if any(t in op for t in ("BRA", "BAR.", "EXIT", "BSYNC", "WARPSYNC", "RET", "CALL")):
        flush()
flush()
