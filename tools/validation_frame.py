#!/usr/bin/env python
"""Per-frame validation step of the reference (managers/OCRNet_Manager.py:146-161: batch 1, torch.no_grad(), loss +
confusion matrix on a 544x960 frame): latency of the fused forward, CUDA-event timed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai2021_cataract_semantic_segmentation_b200 as b200

c, h, w, exp = 25, 544, 960, 3
g = torch.Generator(device="cuda").manual_seed(0)
frames = [(torch.randn((1, c, h, w), generator=g, device="cuda"), torch.randint(0, c + 1, (1, h, w), generator=g, device="cuda"))
          for _ in range(16)]
meter = b200.SegmentationMeter(exp, c)
mod = b200.LovaszSoftmaxWithMetrics({"experiment": exp}, meter)
with torch.no_grad():
    for x, y in frames[:4]:
        mod(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 20
    for _ in range(reps):
        for x, y in frames:
            mod(x, y)
    e1.record()
    torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * len(frames))
print(f"forward-only Lovasz + confusion matrix, 1 x {c} x {h} x {w}: {us:.1f} us / frame  ({h * w / us:.0f} Mpx/s)")

# the same step captured in a CUDA graph (GraphedValidationStep): one launch per frame
meter2 = b200.SegmentationMeter(exp, c)
step = b200.GraphedValidationStep({"experiment": exp}, meter2, (1, c, h, w))
ref_loss = []
meter.reset()
with torch.no_grad():
    for x, y in frames:
        ref_loss.append(float(mod(x, y)))
got = [float(step(x, y)) for x, y in frames]
assert got == ref_loss and torch.equal(meter.cm, meter2.cm), "graph replay differs from the eager step"
for variant, copy in (("copying logits+labels into the static buffers", True), ("inputs already in the static buffers", False)):
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        for x, y in frames:
            if copy:
                step(x, y)
            else:
                step(step.logits, step.labels)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(frames))
    print(f"CUDA graph replay, {variant}: {us:.1f} us / frame  ({h * w / us:.0f} Mpx/s)")
