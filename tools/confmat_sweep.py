#!/usr/bin/env python
"""Confusion-matrix-only sweep (BASELINE configs[4]: validation frames, C=25, 540x960): Mpixel/s and fraction of the
measured HBM peak of b200seg_confmat_accumulate at several batch sizes, int32 labels as the reference's loader gives."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miccai2021_cataract_semantic_segmentation_b200 as b200

peak = 6549.8
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
c, h, w = 25, 540, 960
g = torch.Generator(device="cuda").manual_seed(0)
for n in (1, 8, 64):
    x = torch.randn((n, c, h, w), generator=g, device="cuda")
    y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda", dtype=torch.int32)
    meter = b200.SegmentationMeter(3, c)
    reps = 200 if n == 1 else (50 if n == 8 else 10)
    for _ in range(5):
        meter.update(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        meter.update(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    px = n * h * w
    gbs = px * (4 * c + 4) / (ms * 1e-3) / 1e9
    print(f"batch {n:3d}: {ms * 1e3:8.1f} us / call  {px / ms / 1e3:9.0f} Mpx/s  {gbs:7.0f} GB/s = {100 * gbs / peak:5.1f} % of the measured HBM peak")
    meter.check()
