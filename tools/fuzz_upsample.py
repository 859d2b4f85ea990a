#!/usr/bin/env python
"""Randomised sweep of the fused bilinear upsampling + Lovasz path on a GPU box: random low-resolution / output geometries,
class counts, modes, label dtypes, emission paths and logit styles.  Per case: the in-kernel interpolation equals ATen's bit
for bit; loss (1e-6) and confusion matrix (exact) equal F.interpolate + the full-resolution kernels; the low-resolution gradient
agrees with the float64 adjoint of the
full-resolution gradient to 1e-5 of its maximum (and with ATen's fp32 result within ATen's own distance from that referee); every fourth case is also checked against oracle/port.py on the upsampled logits.
    python tools/fuzz_upsample.py [n_cases] [seed]"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native, upsampled
from oracle import port

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
lib = _native.load()
bad = 0
worst_f = worst_a = 0.0
for case in range(n_cases):
    c, exp = [(8, 1), (17, 2), (25, 3)][rng.randint(3)]
    n = int(rng.randint(1, 4))
    W = int(rng.choice([32, 64, 96, 160, 224, 480]))
    H = int(rng.choice([8, 17, 33, 64, 100, 135, 272]))
    # low-resolution size: horizontal scale >= ~3.3 (the fused kernels' range), vertical anything from 1 row to H rows
    w = int(rng.randint(1, max(2, int(W / 3.4)) + 1))
    h = int(rng.choice([1, 2, max(1, H // 8), max(1, H // 4), max(1, H // 2), H, rng.randint(1, H + 1)]))
    style = rng.randint(3)
    g = torch.Generator().manual_seed(int(rng.randint(1 << 30)))
    hi = c + 1 if exp != 1 else c
    if style == 0:
        low = torch.randn((n, c, h, w), generator=g) * float(rng.choice([0.5, 2.0, 5.0]))
        y = torch.randint(0, hi, (n, H, W), generator=g)
    elif style == 1:                                          # trained-like
        coarse = torch.randint(0, hi, (n, h, w), generator=g)
        y = F.interpolate(coarse[:, None].float(), size=(H, W), mode="nearest")[:, 0].long()
        low = (5.0 * F.one_hot(coarse.clamp(max=c - 1), c).permute(0, 3, 1, 2).float() + torch.randn((n, c, h, w), generator=g)).contiguous()
    else:                                                     # ties: logits on a coarse grid, few classes present
        low = torch.round(torch.randn((n, c, h, w), generator=g) * 2) / 2
        y = torch.randint(0, min(3, c), (n, H, W), generator=g)
    kw = {}
    if rng.rand() < 0.4:
        kw["per_image"] = True
    mode = rng.randint(3)
    if mode == 1 and exp != 1:
        kw["classes_to_ignore"] = c
    elif mode == 2:
        kw["keep_absent"] = 1
    ldt = [torch.int64, torch.int32, torch.uint8][rng.randint(3)]
    _native.set_tuning(emit_path=int(rng.randint(3)), sort_path=int(rng.choice([0, 0, 0, 1])))
    tag = f"case {case}: C={c} n={n} {h}x{w}->{H}x{W} style={style} kw={kw} labels={ldt}"
    try:
        lowd = low.cuda()
        yd = y.cuda().to(ldt)
        if not upsampled.upsample_supported(lowd, yd):
            print(tag, "-> outside the fused kernels (skipped)")
            continue
        ref_up = F.interpolate(lowd, size=(H, W), mode="bilinear", align_corners=True)
        mine = torch.empty_like(ref_up)
        _native.check(lib.b200seg_debug_upsample(lowd.data_ptr(), n * c, h, w, H, W, mine.data_ptr(), -1,
                                                 torch.cuda.current_stream().cuda_stream), "debug_upsample")
        assert torch.equal(mine.view(torch.int32), ref_up.view(torch.int32)), "interpolation differs from ATen"
        cm_f = torch.zeros((c, c), dtype=torch.int64, device="cuda"); st_f = torch.zeros(1, dtype=torch.int32, device="cuda")
        cm_u = torch.zeros((c, c), dtype=torch.int64, device="cuda"); st_u = torch.zeros(1, dtype=torch.int32, device="cuda")
        drop = c if exp != 1 else None
        lf = lowd.clone().requires_grad_(True)
        loss_f = b200.lovasz_softmax_upsampled(lf, yd, confusion=cm_f, confusion_drop_label=drop, status=st_f, **kw)
        lu = lowd.clone().requires_grad_(True)
        full = F.interpolate(lu, size=(H, W), mode="bilinear", align_corners=True)
        full.retain_grad()
        loss_u = b200.lovasz_softmax(full, yd, confusion=cm_u, confusion_drop_label=drop, status=st_u, **kw)
        a, b = float(loss_f.detach()), float(loss_u.detach())
        assert abs(a - b) <= 1e-6 * max(abs(b), 1e-30), f"loss {a} vs {b}"
        assert torch.equal(cm_f, cm_u) and int(st_f) == int(st_u), "confusion matrix / status differ"
        cm_s = torch.zeros((c, c), dtype=torch.int64, device="cuda"); st_s = torch.zeros(1, dtype=torch.int32, device="cuda")
        b200.accumulate_confusion_matrix_upsampled(lowd, yd, cm_s, st_s, drop)      # the standalone kernel on the same pixels
        assert torch.equal(cm_s, cm_u) and int(st_s) == int(st_u), "confusion matrix from low-resolution logits differs"
        if loss_f.requires_grad and loss_u.requires_grad:
            loss_f.backward(); loss_u.backward()
            gmax = float(lu.grad.abs().max())
            if gmax > 0:
                # the referee: float64 adjoint of the interpolation applied to the (bit-identical) full-resolution gradient;
                # ATen's own fp32 atomics drift from it by up to ~2e-5 when a low-resolution cell collects 10^4 pixels
                l64 = lowd.double().requires_grad_(True)
                F.interpolate(l64, size=(H, W), mode="bilinear", align_corners=True).backward(full.grad.double())
                # scale: the largest gradient, but not less than 1 % of the largest sum of |terms| a cell collects (a 1 x 1
                # source makes the true gradient cancel to ~0, where an error relative to it means nothing)
                a64 = lowd.double().requires_grad_(True)
                F.interpolate(a64, size=(H, W), mode="bilinear", align_corners=True).backward(full.grad.abs().double())
                scale = max(gmax, 0.01 * float(a64.grad.max()))
                e_f = float((lf.grad.double() - l64.grad).abs().max()) / scale
                e_a = float((lu.grad.double() - l64.grad).abs().max()) / scale
                worst_f, worst_a = max(worst_f, e_f), max(worst_a, e_a)
                # (the referee forms its weights in float64: where h ~ H their fp32 rounding alone moves both fp32 results by
                # ~1e-5, identically -- hence the second alternative)
                assert e_f <= 1e-5 or e_f <= 1.1 * e_a, f"gradient error vs the float64 adjoint {e_f} (ATen's: {e_a})"
                err = float((lf.grad - lu.grad).abs().max()) / scale
                assert err <= 1e-5 + 1.5 * e_a, f"gradient error vs ATen {err} (ATen vs float64: {e_a})"
        if case % 4 == 0 and not kw.get("keep_absent"):
            lo = lowd.clone().requires_grad_(True)
            ref = port.lovasz_softmax(F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=True), yd.long(), exp,
                                      per_image=kw.get("per_image", False), classes_to_ignore=kw.get("classes_to_ignore"))
            r = float(ref.detach()) if torch.is_tensor(ref) and ref.numel() == 1 else 0.0
            assert abs(a - r) <= 1e-5 * max(abs(r), 1e-30) + 1e-12, f"loss {a} vs oracle {r}"
    except Exception as e:                                    # noqa: BLE001
        bad += 1
        print(tag, "FAILED:", repr(e)[:300])
_native.set_tuning(emit_path=0, sort_path=0)
print(f"{n_cases} cases, {bad} bad; worst gradient error against the float64 adjoint: fused {worst_f:.2e}, ATen fp32 {worst_a:.2e}")
sys.exit(1 if bad else 0)
