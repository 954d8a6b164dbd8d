"""Per-kernel digest of an `ncu --set full --import-source on` report: headline metrics + warp-stall reasons + the ten
hottest SASS instructions.  usage: python tools/stall_summary.py <report.ncu-rep> > profiles/<name>.txt  (no GPU needed)"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
for k, r in enumerate(data):
    print(f"=== launch {k}: {r[ix['Kernel Name']][:100]}")
    for key in keys:
        if key in ix:
            print(f"  {key:75s} {r[ix[key]]:>14s} {units[ix[key]]}")
    st = {h: float(r[i]) for h, i in ix.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]}
    print("  warp stall reasons (warps stalled per issue-active cycle):")
    for h, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
        print(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:2]:
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    h2 = lines[1]
    j = {h: i for i, h in enumerate(h2)}
    body = [l for l in lines[2:] if len(l) == len(h2)]
    tot = sum(int(l[j["# Samples"]] or 0) for l in body)
    print(f"  hottest instructions of launch 0 ({tot} samples, {len(body)} SASS instructions):")
    for l in sorted(body, key=lambda l: -int(l[j["# Samples"]] or 0))[:10]:
        reasons = Counter({h: int(l[i] or 0) for h, i in j.items() if h.startswith("stall_") and "Not Issued" not in h})
        top = ", ".join(f"{a[6:]} {b}" for a, b in reasons.most_common(2))
        print(f"    {int(l[j['# Samples']] or 0):5d}  {l[j['Source']][:70]:70s} [{top}]")
