"""Per-source-line stall samples of the kernels in an `ncu --set full --import-source on` report.
usage: python tools/source_hot.py <report.ncu-rep> [top=25] [kernel-name substring]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
flt = sys.argv[3] if len(sys.argv) > 3 else None  # only kernels whose name contains this
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# blocks: "File Path",path / "Function Name",.. / header / lines: (line no, source, "-", ...) followed by SASS rows with empty line no
agg = defaultdict(lambda: [0, ""])
fpath = None
hdr = None
kernels = 0
active = True
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kernels += 1
        active = flt is None or flt in r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        continue
    if hdr is None or kernels > len(set([1])) and False:
        continue
    if not active:
        continue
    if r[0] != "" and len(r) > si:
        try:
            n = int(r[si] or 0)
        except ValueError:
            continue
        key = (fpath, int(r[0]))
        agg[key][0] += n
        agg[key][1] = r[1]
tot = sum(v[0] for v in agg.values())
print(f"total samples {tot}")
for (f, ln), (n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:6d} {100.0*n/max(tot,1):5.1f}%  {f}:{ln:<5d} {src.strip()[:110]}")
