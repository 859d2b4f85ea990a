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
print("done")
