"""Per-kernel count of the Blackwell-only SASS mnemonics in the shipped library (no GPU needed):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UBLKCP = bulk copy,
HMMA = legacy mma.sync (must stay 0).  usage: python tools/sass_digest.py > profiles/<round>_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jyutvoice_b200", "libjyutvoice_b200.so")
PAT = [("UTCHMMA", r"\bUTCHMMA\b"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA\b"), ("UTC*MMA(other)", r"\bUTC(?!HMMA\b)[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
       ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("FFMA", r"\bFFMA")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = list(counts)
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dm):
        demangle[n] = d
    print(f"# {os.path.relpath(LIB, ROOT)}: {os.path.getsize(LIB)} bytes, {len(names)} kernels (cuobjdump -sass)")
    hdr = [p[0] for p in PAT]
    print("kernel".ljust(96) + " ".join(h.rjust(14) for h in hdr))
    tot = collections.Counter()
    for n in names:
        c = counts[n]
        tot.update(c)
        short = demangle.get(n, n).replace("void ", "").replace("jv::", "")[:94]
        print(short.ljust(96) + " ".join(str(c[h]).rjust(14) for h in hdr))
    print("TOTAL".ljust(96) + " ".join(str(tot[h]).rjust(14) for h in hdr))


if __name__ == "__main__":
    sys.exit(main())
