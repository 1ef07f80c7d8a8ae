"""Hot SASS lines of one kernel of an .ncu-rep (ncu --set full --import-source on).
usage: python tools/ncu_hot.py <file.ncu-rep> <kernel index in the report> [min percent]"""
import csv, io, subprocess, sys
rep, which = sys.argv[1], int(sys.argv[2])
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
k = kernels[which]
idx = {h: i for i, h in enumerate(k["hdr"])}
I = idx["Instructions Executed"]
tot = sum(int(r[I]) for r in k["rows"])
samples = sum(int(r[idx["# Samples"]]) for r in k["rows"])
print(k["name"][:100], "total warp-inst", tot, "samples", samples)
print("line  sass".ljust(72), " inst%  stall%  smem-wavefronts  ideal")
for n, r in enumerate(k["rows"]):
    ie = int(r[I])
    sp = int(r[idx["# Samples"]])
    if ie < tot * minpct / 100 and sp < samples * minpct / 100:
        continue
    print(str(n).rjust(4), r[idx["Source"]].strip()[:64].ljust(64), "%5.1f" % (ie * 100 / tot), "%6.1f" % (sp * 100 / max(samples, 1)),
          r[idx["L1 Wavefronts Shared"]].rjust(11), r[idx["L1 Wavefronts Shared Ideal"]].rjust(11))
