#!/usr/bin/env python
"""TwoScaleLoss (Lovasz-Lovasz, configs/OCRNet_rf_lvsz.json) on two [8, 25, 540, 960] heads, forward + backward:
heads one after the other on one stream vs. on two streams, CUDA-event timed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai2021_cataract_semantic_segmentation_b200 as b200

n, c, h, w, exp = 8, 25, 540, 960, 3
g = torch.Generator(device="cuda").manual_seed(0)
xa = torch.randn((n, c, h, w), generator=g, device="cuda").requires_grad_(True)
xb = torch.randn((n, c, h, w), generator=g, device="cuda").requires_grad_(True)
y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda")
cfg = lambda: {"interm": {"name": "LovaszSoftmax", "args": [], "weight": 0.4},
               "final": {"name": "LovaszSoftmax", "args": [], "weight": 1.0}, "experiment": exp}
two = b200.TwoScaleLoss(cfg())
la, lb = b200.LovaszSoftmax({"experiment": exp}), b200.LovaszSoftmax({"experiment": exp})


def step_two_streams():
    xa.grad = xb.grad = None
    two(xa, xb, y).backward()


def step_one_stream():
    xa.grad = xb.grad = None
    (lb(xb, y) * 1.0 + la(xa, y) * 0.4).backward()


res = {}
for name, fn in (("one stream", step_one_stream), ("two streams", step_two_streams)):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name] = (e0.elapsed_time(e1) / 20, xa.grad.clone(), xb.grad.clone())
    print(f"{name:12s} {res[name][0]:7.3f} ms / step  ({2 * n * h * w / res[name][0] / 1e3:.0f} Mpx/s over both heads)")
assert torch.equal(res["one stream"][1], res["two streams"][1]) and torch.equal(res["one stream"][2], res["two streams"][2])
print("gradients identical")
