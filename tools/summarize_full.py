"""Reduce an `ncu --set full` report (several launches of the tcgen05 GEMM) to the columns the roofline uses.
usage: python tools/summarize_full.py <report.ncu-rep> <out_prefix>   ->  <out_prefix>_ncu_full_gemm.csv, <out_prefix>_gemm_traffic.json
(ncu -i report --page raw --csv is run here; no GPU needed)"""
import csv
import io
import json
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


with open(prefix + "_ncu_full_gemm.csv", "w") as f:
    f.write("Kernel Name," + ",".join(f"{c} [{units[ix[c]]}]" for c in cols) + "\n")
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0]
        f.write(name + "," + ",".join(r[ix[c]] for c in cols) + "\n")
# the roofline object is about the contraction kernels: other kernels of the capture stay in the CSV only
data = [r for r in data if "gemm_taps_tc" in r[ix["Kernel Name"]] or "mlp_fused" in r[ix["Kernel Name"]]]
n = len(data)
rd = sum(to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) for r in data) / n
wr = sum(to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]) for r in data) / n
us = sum(to_us(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]]) for r in data) / n
json.dump({"kernel": "gemm_taps_tc_kernel + mlp_fused_pair_kernel", "source": f"ncu --set full --clock-control none, {n} contraction launches of one "
           "estimator forward (B=64 x 300 frames, bf16): " + prefix.split("/")[-1] + "_ncu_full_gemm.csv",
           "launches": n, "avg_duration_us": us, "avg_dram_read_MB": rd / 1e6, "avg_dram_write_MB": wr / 1e6,
           "avg_dram_bytes_per_launch": rd + wr}, open(prefix + "_gemm_traffic.json", "w"), indent=1)
print(n, "launches; avg", round(us, 2), "us; DRAM", round((rd + wr) / 1e6, 1), "MB per launch")
