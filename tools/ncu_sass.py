#!/usr/bin/env python
"""SASS windows around the most-stalled instructions of one kernel in an .ncu-rep.
    tools/ncu_sass.py <rep> <kernel regex> [n hot spots] [lines before]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
nhot = int(sys.argv[3]) if len(sys.argv) > 3 else 4
before = int(sys.argv[4]) if len(sys.argv) > 4 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "sass",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(raw)) if len(r) > 10]
h = rows[0]
si, src, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
data = [r for r in rows[1:] if r[si].isdigit()]
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:nhot]
for i in sorted(idx):
    print('-----')
    for j in range(max(0, i - before), min(len(data), i + 3)):
        print(f"{data[j][si]:>6s} {data[j][ie]:>8s}  {data[j][src][:120]}")
