#!/usr/bin/env python
"""Per-source-line summary of one kernel from an .ncu-rep: stall samples and warp instructions executed.

    tools/ncu_lines.py <rep> <kernel regex> [top N] [launch index among matches]
"""
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "cuda,sass"]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, hdr, lines, seen_fn = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        samples = int(r[hdr.index("# Samples")]); inst = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    stalls = {k[6:]: int(v) for k, v in zip(hdr, r) if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    lines.append((samples, inst, cur_file, r[0], r[1].strip()[:90], stalls))
tot_s = sum(l[0] for l in lines) or 1
tot_i = sum(l[1] for l in lines) or 1
print(f"total samples {tot_s}  total warp instructions {tot_i}")
for s, i, f, ln, src, st in sorted(lines, key=lambda l: -l[0])[:top]:
    tops = ",".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{ln:>4s}  {src:90s} {tops}")
