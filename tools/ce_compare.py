#!/usr/bin/env python
"""LossWrapper({'CrossEntropyLoss': 1, 'LovaszSoftmax': 1}) on [8, 25, 540, 960] logits, forward + backward:
fused pair (one pass each way) vs. this package's Lovasz + torch's cross entropy side by side, CUDA-event timed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai2021_cataract_semantic_segmentation_b200 as b200

n, c, h, w, exp = 8, 25, 540, 960, 3
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, c, h, w), generator=g, device="cuda")
y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda")
fused = b200.LossWrapper({"losses": {"CrossEntropyLoss": 1.0, "LovaszSoftmax": 1.0}, "experiment": exp, "device": "cuda"})
lov = b200.LovaszSoftmax({"experiment": exp})
xr = x.clone().requires_grad_(True)


def step_fused():
    xr.grad = None
    fused(None, xr, y).backward()


def step_split():
    xr.grad = None
    (lov(xr, y) + torch.nn.functional.cross_entropy(xr, y, ignore_index=25)).backward()


def step_lovasz_only():
    xr.grad = None
    lov(xr, y).backward()


for name, fn in (("lovasz only", step_lovasz_only), ("fused CE + Lovasz", step_fused), ("Lovasz + torch CE", step_split)):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:20s} {e0.elapsed_time(e1) / 20:7.3f} ms / step")
