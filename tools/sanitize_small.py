#!/usr/bin/env python
"""Small end-to-end invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miccai2021_cataract_semantic_segmentation_b200 as b200

g = torch.Generator().manual_seed(3)
for (n, c, h, w, exp, cfg) in [(2, 25, 64, 96, 3, {}), (2, 17, 64, 96, 2, {"per_image": True}),
                               (1, 8, 33, 47, 1, {}), (2, 17, 40, 64, 2, {"classes_to_consider": "all", "classes_to_ignore": 17}),
                               (1, 5, 32, 32, 1, {})]:
    x = torch.randn((n, c, h, w), generator=g).cuda().requires_grad_(True)
    y = torch.randint(0, c + (exp != 1), (n, h, w), generator=g).cuda()
    meter = b200.SegmentationMeter(exp, c)
    if c in (8, 17, 25):
        loss = b200.LovaszSoftmaxWithMetrics({"experiment": exp, **cfg}, meter)(x, y)
    else:
        loss = b200.LovaszSoftmax({"experiment": exp, **cfg})(x, y)
    loss.backward()
    cm = b200.t_get_confusion_matrix(x.detach(), y.int()) if c in (8, 17, 25) else None
    torch.cuda.synchronize()
    print(n, c, h, w, cfg, float(loss.detach()), float(x.grad.abs().max()))
# the other emission path, the fused cross-entropy pair, confident logits (streaming emission) and a big sort segment
from miccai2021_cataract_semantic_segmentation_b200 import _native
x = torch.randn((2, 25, 64, 96), generator=g).cuda().requires_grad_(True)
y = torch.randint(0, 26, (2, 64, 96), generator=g).cuda()
for path in (1, 2):
    _native.set_tuning(emit_path=path)
    lov, ce = b200.LovaszSoftmaxCE({"experiment": 3})(x, y)
    (lov + ce).backward()
    torch.cuda.synchronize()
    print("emit_path", path, float(lov.detach()), float(ce.detach()))
_native.set_tuning(emit_path=0)
xz = torch.zeros((1, 25, 160, 256)).cuda().requires_grad_(True)          # every pair is a candidate: 40 960 x 25, > 128 tiles
yz = torch.zeros((1, 160, 256), dtype=torch.long).cuda()
b200.LovaszSoftmax({"experiment": 3, "classes_to_consider": "all"})(xz, yz).backward()
torch.cuda.synchronize()
# windowed IoU map and OHEM cross entropy: pipelined and scalar kernels, ignored and rank-decided cases
for (n, c, h, w) in [(2, 25, 32, 48), (1, 12, 21, 19)]:
    xs = torch.randn((n, c, h, w), generator=g).cuda()
    ys = torch.randint(0, c, (n, h, w), generator=g).cuda()
    m = b200.sliding_miou(xs, ys, 7, 4)
    xo = xs.clone().requires_grad_(True)
    yo = ys.clone()
    yo[:, :4] = c if c == 25 else 0
    for cfg in ({"experiment": 3, "min_kept": 200, "thresh": 0.01}, {"experiment": 3}):
        lo = b200.OhemCrossEntropy(cfg if c == 25 else {k: v for k, v in cfg.items() if k != "experiment"})(xo, yo)
        lo.backward()
    torch.cuda.synchronize()
    print("sliding / ohem", n, c, h, w, float(m.mean()), float(lo.detach()))
# fused bilinear upsampling: both emission paths, per-image, fused cross entropy, stride-like and odd geometries
for (n, c, h, w, H, W, kw) in [(2, 25, 9, 12, 72, 96, {}), (2, 17, 16, 16, 64, 64, {"per_image": True}),
                               (1, 8, 5, 7, 33, 64, {"ce_ignore_index": None}), (1, 25, 1, 6, 8, 32, {})]:
    low = (torch.randn((n, c, h, w), generator=g) * 3).cuda().requires_grad_(True)
    yu = torch.randint(0, c, (n, H, W), generator=g).cuda()
    cmu = torch.zeros((c, c), dtype=torch.int64, device="cuda")
    stu = torch.zeros(1, dtype=torch.int32, device="cuda")
    for path in (1, 2):
        _native.set_tuning(emit_path=path)
        res = b200.lovasz_softmax_upsampled(low, yu, confusion=cmu, status=stu, **kw)
        tot = res[0] + res[1] if isinstance(res, tuple) else res
        tot.backward()
        torch.cuda.synchronize()
    mu = b200.SegmentationMeter({8: 1, 17: 2, 25: 3}[c], c)
    mu.update_upsampled(low, yu)                        # confusion matrix straight from the low-resolution logits
    torch.cuda.synchronize()
    print("upsampled", n, c, h, w, H, W, kw, float(tot.detach()), float(low.grad.abs().max()), int(mu.cm.sum()))
_native.set_tuning(emit_path=0)
mg = b200.SegmentationMeter(1, 12)                      # run-time class count
mg.update_upsampled(torch.randn((2, 12, 5, 7), generator=g).cuda(), torch.randint(0, 12, (2, 33, 64), generator=g).cuda().int())
torch.cuda.synchronize()
print("done")
