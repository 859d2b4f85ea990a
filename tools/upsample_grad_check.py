#!/usr/bin/env python
"""Is the low-resolution gradient of the fused path as accurate as ATen's?  Both are compared with the float64 adjoint of
the interpolation applied to the SAME full-resolution gradient (the one the full-resolution backward kernel returns)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import miccai2021_cataract_semantic_segmentation_b200 as b200
from test_gpu_upsample import _inputs, CASES

only = sys.argv[1:] 
for name, (n, c, h, w, H, W), dist, opt in CASES:
    if only and name not in only:
        continue
    low, y = _inputs(n, c, h, w, H, W, 99 + n * c + h, dist, opt.get("ignore", False))
    yd = y.cuda()
    kw = dict(per_image=opt.get("per_image", False), classes_to_ignore=opt.get("classes_to_ignore"),
              keep_absent=1 if opt.get("classes") == "all" else 0)
    lf = low.cuda().requires_grad_(True)
    b200.lovasz_softmax_upsampled(lf, yd, **kw).backward()
    lu = low.cuda().requires_grad_(True)
    full = F.interpolate(lu, size=(H, W), mode="bilinear", align_corners=True)
    full.retain_grad()
    b200.lovasz_softmax(full, yd, **kw).backward()
    # float64 adjoint of the interpolation applied to the full-resolution gradient
    l64 = low.cuda().double().requires_grad_(True)
    F.interpolate(l64, size=(H, W), mode="bilinear", align_corners=True).backward(full.grad.double())
    ref = l64.grad
    gmax = float(ref.abs().max())
    e_f = float((lf.grad.double() - ref).abs().max()) / gmax
    e_a = float((lu.grad.double() - ref).abs().max()) / gmax
    print(f"{name:28s} max|g| {gmax:.3e}   fused vs f64 {e_f:.2e}   ATen fp32 vs f64 {e_a:.2e}   fused vs ATen "
          f"{float((lf.grad - lu.grad).abs().max()) / gmax:.2e}")
