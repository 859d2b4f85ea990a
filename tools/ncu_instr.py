#!/usr/bin/env python
"""Warp instructions per source line of one kernel, normalised by a unit count (e.g. warp-tiles).
   tools/ncu_instr.py <rep> <kernel regex> <units> [top]"""
import subprocess, csv, io, sys
rep, kre, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "cuda,sass",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; cur = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    try: inst = int(r[hdr.index("Instructions Executed")])
    except ValueError: continue
    lines.append((inst, cur, r[0], r[1].strip()[:100]))
tot = sum(l[0] for l in lines)
print(f"total {tot}  per unit {tot / units:.1f}")
for inst, f, ln, src in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"{inst / units:7.1f} {100 * inst / tot:5.1f}%  {f}:{ln:>4s} {src}")
