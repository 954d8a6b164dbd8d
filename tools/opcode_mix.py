"""Executed warp-instruction counts per SASS opcode of the first kernel in an ncu --set full --import-source report."""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
cnt = Counter(); smp = Counter()
k = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        if k > 1: break
        continue
    if r and "Source" in r and "Instructions Executed" in r:
        hdr = r; si = hdr.index("Source"); ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples"); continue
    if hdr and len(r) > ie:
        try: n = int(r[ie] or 0)
        except ValueError: continue
        op = r[si].strip().split()
        if not op: continue
        o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        o = o.split(".")[0] + ("." + o.split(".")[1] if o.startswith(("MUFU", "F2FP", "LDTM", "STTM", "SYNCS", "FMNMX", "UTCHMMA")) and "." in o else "")
        cnt[o] += n; smp[o] += int(r[ss] or 0)
tot = sum(cnt.values())
print(f"total warp instructions {tot}")
for o, n in cnt.most_common(30):
    print(f"{o:18s} {n:12d} {100*n/tot:5.1f}%   samples {smp[o]}")
