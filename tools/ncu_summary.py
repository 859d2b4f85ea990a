#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) per kernel launch: duration, DRAM bytes, throughput %, occupancy, top stalls."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default


def f(r, name):
    try:
        return float(get(r, name).replace(",", ""))
    except ValueError:
        return float("nan")


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
if not stall_cols:
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled") or
                  (h.startswith("smsp__average_warps_issue_stalled"))]
print(f"{'kernel':42s} {'us':>8s} {'dramR MB':>9s} {'dramW MB':>9s} {'dram%':>6s} {'sm%':>6s} {'occ%':>6s} {'regs':>5s}  top stalls")
for r in data:
    name = get(r, "Kernel Name")[:42]
    dur = f(r, "gpu__time_duration.sum")
    du = units[col["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in col else ""
    if du in ("ns", "nsecond"):
        dur /= 1e3
    elif du in ("ms", "msecond"):
        dur *= 1e3
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    ru, wu = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    rd *= scale.get(ru, 1e-6); wr *= scale.get(wu, 1e-6)
    dram = f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    smp = f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    occ = f(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    regs = get(r, "launch__registers_per_thread")
    stalls = []
    for h in stall_cols:
        v = f(r, h)
        if v == v:
            stalls.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    stalls.sort(reverse=True)
    top = ", ".join(f"{n}={v:.1f}" for v, n in stalls[:4])
    print(f"{name:42s} {dur:8.1f} {rd:9.1f} {wr:9.1f} {dram:6.1f} {smp:6.1f} {occ:6.1f} {regs:>5s}  {top}")
