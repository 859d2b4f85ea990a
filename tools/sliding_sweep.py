#!/usr/bin/env python
"""Time the windowed mean IoU map (b200seg_sliding_miou) on BASELINE-size frames, next to the oracle's torch
restatement run on the same device.  Reports the class-map kernel against the measured copy bandwidth
(algorithmic bytes: (4C + label bytes + 2) per pixel).   python tools/sliding_sweep.py [n] [c]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miccai2021_cataract_semantic_segmentation_b200 as b200
from oracle import port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h, w, k, s = 544, 960, 7, 4
peak = 6650.0                                      # fallback of B200_PROFILING.md
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
g = torch.Generator().manual_seed(0)
y = torch.randint(0, c, (n, h // 8, w // 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2).cuda()
x = (3.0 * torch.nn.functional.one_hot(y, c).permute(0, 3, 1, 2).float() + torch.randn(n, c, h, w, device="cuda")).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for ldt, lb in ((torch.int64, 8), (torch.uint8, 1)):
    yy = y.to(ldt)
    ms = timed(lambda: b200.sliding_miou(x, yy, k, s, original_size=False, validate=False))
    alg = (4 * c + lb + 2) * n * h * w
    print(json.dumps({"op": "sliding_miou", "n": n, "c": c, "labels": str(ldt), "ms": round(ms, 4),
                      "alg_GBps_whole_call": round(alg / ms / 1e6, 1), "copy_peak_GBps": peak,
                      "Mpx_per_s": round(n * h * w / ms / 1e3, 1)}))
ms_ref = timed(lambda: port.sliding_miou(x[:1], y[:1], k, s, original_size=False), reps=5)
print(json.dumps({"op": "torch restatement on the device, 1 frame", "ms": round(ms_ref, 3)}))
import time
xc, yc = x[:1].cpu(), y[:1].cpu()
port.sliding_miou(xc, yc, k, s, original_size=False)
t0 = time.perf_counter()
for _ in range(3):
    port.sliding_miou(xc, yc, k, s, original_size=False)
cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
print(json.dumps({"op": "oracle on the host CPU, 1 frame", "ms": round(cpu_ms, 1), "threads": torch.get_num_threads(),
                  "Mpx_per_s": round(h * w / cpu_ms / 1e3, 2)}))
got = b200.sliding_miou(x, y, k, s, original_size=False)
ref = torch.cat([port.sliding_miou(x[i:i + 1], y[i:i + 1], k, s, original_size=False) for i in range(n)])
print("max abs diff vs oracle on the device:", float((got - ref).abs().max()))
