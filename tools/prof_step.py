#!/usr/bin/env python
"""One benchmark step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off ...`."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miccai2021_cataract_semantic_segmentation_b200 as b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--classes", type=int, default=25)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--per-image", action="store_true")
ap.add_argument("--blocky", action="store_true")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--confmat-only", action="store_true")
ap.add_argument("--ce", action="store_true", help="fused cross entropy + Lovasz pair (LovaszSoftmaxCE)")
a = ap.parse_args()
c, exp = a.classes, {8: 1, 17: 2, 25: 3}[a.classes]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((a.batch, c, 540, 960), generator=g, device="cuda")
y = torch.randint(0, c + (exp != 1), (a.batch, 540, 960), generator=g, device="cuda")
if a.blocky:
    coarse = torch.randint(0, c // 2, (a.batch, 27, 48), generator=g, device="cuda")
    y = coarse.repeat_interleave(20, 1).repeat_interleave(20, 2).contiguous()
    x = x + 6.0 * torch.nn.functional.one_hot(y, c).permute(0, 3, 1, 2).float()
meter = b200.SegmentationMeter(exp, c)
mod = b200.LovaszSoftmaxWithMetrics({"experiment": exp, "per_image": a.per_image}, meter)
pair = b200.LovaszSoftmaxCE({"experiment": exp, "per_image": a.per_image}, meter)
xr = x.clone().requires_grad_(True)


def step():
    if a.confmat_only:
        meter.update(x, y)
        return
    meter.reset()
    xr.grad = None
    if a.ce:
        lov, ce = pair(xr, y)
        loss = lov + ce
    else:
        loss = mod(xr, y)
    loss.backward()
    meter.summary()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
