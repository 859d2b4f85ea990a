#!/usr/bin/env python
"""Fused bilinear upsampling + Lovasz-Softmax (+ confusion matrix) against F.interpolate followed by the full-resolution
kernels: one training-loss step (forward + backward to the LOW-resolution logits), CUDA-event timed, on the reference's
two model geometries (stride 8: OCRNet, models/OCR.py:126; stride 4: DeepLabv3+, models/DeepLabv3Plus.py:65)."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native

n, c, H, W = 8, 25, 544, 960
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
out = {}
for name, (h, w), dist in (("stride8_d1", (68, 120), "d1"), ("stride4_d1", (136, 240), "d1"), ("stride8_trained_like", (68, 120), "d2")):
    g = torch.Generator(device="cuda").manual_seed(0)
    if dist == "d1":
        low = torch.randn((n, c, h, w), generator=g, device="cuda") * 2
        y = torch.randint(0, c + 1, (n, H, W), generator=g, device="cuda")
    else:
        coarse = torch.randint(0, c, (n, h, w), generator=g, device="cuda")
        y = F.interpolate(coarse[:, None].float(), size=(H, W), mode="nearest")[:, 0].long()
        noisy = coarse.clone()
        flips = torch.rand((n, h, w), generator=g, device="cuda") < 0.10
        noisy[flips] = torch.randint(0, c, (int(flips.sum()),), generator=g, device="cuda")
        low = 6.0 * F.one_hot(noisy, c).permute(0, 3, 1, 2).float() + torch.randn((n, c, h, w), generator=g, device="cuda")
    low.requires_grad_(True)
    cm = torch.zeros((c, c), dtype=torch.int64, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")

    def unfused():
        full = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=True)
        loss = b200.lovasz_softmax(full, y, confusion=cm, confusion_drop_label=c, status=st)
        loss.backward()
        return loss

    def fused():
        loss = b200.lovasz_softmax_upsampled(low, y, confusion=cm, confusion_drop_label=c, status=st)
        loss.backward()
        return loss

    res = {}
    for label, fn in (("interpolate_then_loss", unfused), ("fused", fused)):
        for _ in range(5):
            low.grad = None
            val = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            low.grad = None
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[label] = {"ms_per_step": e0.elapsed_time(e1) / reps, "loss": float(val.detach())}
    # stage times of both steps (CUDA events between the kernels)
    lib = _native.load()
    import ctypes
    names = ["memset", "stats", "finalize", "emit", "prepare", "count", "partition", "local", "fallback", "(autograd)", "backward"]
    for label, fn in (("interpolate_then_loss", unfused), ("fused", fused)):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        for e in evs:
            e.record()
        torch.cuda.synchronize()
        arr = (ctypes.c_void_p * 11)(*[ctypes.c_void_p(e.cuda_event) for e in evs])
        _native.check(lib.b200seg_set_stage_events(arr, 11), "set_stage_events")
        low.grad = None
        fn()
        torch.cuda.synchronize()
        _native.check(lib.b200seg_set_stage_events(None, 0), "clear stage events")
        res[label]["stage_us"] = {names[i]: round(evs[i - 1].elapsed_time(evs[i]) * 1e3, 1) for i in range(1, 11)}
    res["speedup"] = res["interpolate_then_loss"]["ms_per_step"] / res["fused"]["ms_per_step"]
    out[name] = res
    print(name, json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/upsample_compare.json", "w"), indent=1)
