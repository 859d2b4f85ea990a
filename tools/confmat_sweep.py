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
# label / prediction statistics: uniform (every lane of a warp hits another bin), blocky (16 x 16 blocks of one label and a
# confident prediction of it: whole warps on ONE bin -- the worst case for the shared-memory atomics), one-class (everything
# on a single bin, the CaDIS cornea-dominated frames taken to the limit)
for dist in ("uniform", "blocky", "one-class"):
  print(f"--- labels: {dist}")
  for n in (1, 8, 64):
    x = torch.randn((n, c, h, w), generator=g, device="cuda")
    y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda", dtype=torch.int32)
    if dist == "blocky":
        coarse = torch.randint(0, c, (n, (h + 15) // 16, (w + 15) // 16), generator=g, device="cuda")
        y = coarse.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :h, :w].contiguous().int()
        for i in range(n):
            x[i] += 6.0 * torch.nn.functional.one_hot(y[i].long(), c).permute(2, 0, 1).float()
    elif dist == "one-class":
        y = torch.full((n, h, w), 3, device="cuda", dtype=torch.int32)
        x[:, 3] += 8.0
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
