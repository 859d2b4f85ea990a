#!/usr/bin/env python
"""profiles/dram_traffic.json from an `ncu --set full` capture of one step: DRAM bytes (read + write) per launch of the
kernel behind each bench.py stage.  usage: tools/dram_traffic.py <rep> [out.json]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/dram_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(h)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage_of = {"stats_kernel": "stats(+fused confmat, records)", "backward_kernel": "backward", "emit_kernel_cta": "emit",
            "jaccard_kernel": "jaccard+loss", "hyb_count": "hyb_count", "hyb_partition": "hyb_partition",
            "hyb_local": "hyb_local(rank+jaccard+loss)"}
res = {}
for r in data:
    name = r[col["Kernel Name"]]
    b = sum(float(r[col[k]].replace(",", "")) * scale[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    for key, stage in stage_of.items():
        if key in name:
            res[stage] = int(b)
res["_source"] = rep.split("/")[-1]
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
res["_csrc"] = bench.csrc_digest()
json.dump(res, open(out, "w"), indent=1)
print(res)
