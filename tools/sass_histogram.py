#!/usr/bin/env python
"""Opcode histogram per kernel of libb200seg.so (cuobjdump -sass): what the hot kernels are made of -- LDGSTS (cp.async),
UTMALDG/UTMASTG (TMA), ATOMS/ATOMG/RED, MATCH/VOTE/REDUX, MUFU, BAR.   usage: tools/sass_histogram.py [lib] > profiles/..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "miccai2021_cataract_semantic_segmentation_b200", "libb200seg.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur is not None:
        cur[m.group(1).split(".")[0]] += 1
KEY = ["LDG", "STG", "LDGSTS", "UTMALDG", "UTMASTG", "UBLKCP", "LDS", "STS", "ATOMS", "ATOMG", "RED", "MATCH", "VOTE", "REDUX", "SHFL",
       "MUFU", "BAR", "FFMA", "IMAD", "HMMA", "UTCHMMA", "CCTL"]
print(f"{'kernel':70s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in KEY))
for name, cnt in kernels.items():
    d = demangle(name)
    d = re.sub(r"\((int|bool)\)", "", d)
    d = d[:d.rfind("(")] if d.endswith(")") else d
    if not any(t in d for t in ("stats_kernel_async<25", "emit_kernel", "hyb_", "backward_kernel_async<25", "sort_fallback", "sort_prepare",
                                "finalize_decide", "confmat_kernel", "jaccard_kernel", "sort_scatter", "sort_count", "ohem_", "class_map",
                                "window_iou", "metrics_kernel")):
        continue
    print(f"{d[:70]:70s} {sum(cnt.values()):6d} " + " ".join(f"{cnt.get(k, 0):7d}" for k in KEY))
tot = collections.Counter()
for cnt in kernels.values():
    tot.update(cnt)
print(f"{'ALL KERNELS OF THE LIBRARY':70s} {sum(tot.values()):6d} " + " ".join(f"{tot.get(k, 0):7d}" for k in KEY))
