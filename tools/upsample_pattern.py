#!/usr/bin/env python
"""Which multiply-add contraction pattern does ATen's CUDA upsample_bilinear2d (align_corners=True) use?  Materialises the
36 candidates of b200seg_debug_upsample and compares each with F.interpolate bit for bit on several shapes (the reference's
stride-8 / stride-4 logits, models/OCR.py:126, models/DeepLabv3Plus.py:65, and odd ones)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miccai2021_cataract_semantic_segmentation_b200 import _native

lib = _native.load()
shapes = [(2, 25, 68, 120, 544, 960), (2, 17, 136, 240, 544, 960), (1, 8, 68, 120, 540, 960), (3, 5, 7, 9, 33, 64),
          (1, 3, 1, 5, 4, 32), (1, 2, 5, 1, 9, 32), (1, 4, 13, 17, 13, 17), (1, 4, 40, 50, 30, 32)]
ok = {p: True for p in range(36) if (p >> 1) & 3 < 3 and (p >> 3) & 3 < 3}
g = torch.Generator(device="cuda").manual_seed(0)
for (n, c, h, w, H, W) in shapes:
    lo = torch.randn((n, c, h, w), generator=g, device="cuda") * 4
    ref = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=True)
    out = torch.empty_like(ref)
    for pat in list(ok):
        _native.check(lib.b200seg_debug_upsample(lo.data_ptr(), n * c, h, w, H, W, out.data_ptr(), pat,
                                                 torch.cuda.current_stream().cuda_stream), "debug_upsample")
        same = torch.equal(out.view(torch.int32), ref.view(torch.int32))
        if not same:
            ok[pat] = False
        nd = int((out.view(torch.int32) != ref.view(torch.int32)).sum())
        print(f"shape {(n, c, h, w, H, W)} pattern {pat:2d} (lambda-fma {pat & 1}, inner {(pat >> 1) & 3}, outer {(pat >> 3) & 3}): "
              f"{'EQUAL' if same else f'{nd} of {out.numel()} differ'}")
print("patterns equal to ATen on every shape:", [p for p, v in ok.items() if v])
