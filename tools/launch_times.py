#!/usr/bin/env python
"""Print per-launch durations from an ncu --csv launch list (gpu__time_duration.sum)."""
import csv
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
tot = 0
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    if u == 'ns': v /= 1e3
    elif u == 'ms': v *= 1e3
    tot += v
    print(f"{v:9.1f} us  {r[ki][:80]}")
print(f"{tot:9.1f} us  total")
