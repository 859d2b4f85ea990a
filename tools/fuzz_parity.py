#!/usr/bin/env python
"""Randomised parity sweep on a GPU box: random shapes / class counts / modes / label dtypes / logit styles against the
oracle executed on the device (loss 1e-5, gradient 1e-5 of its maximum), confusion matrix bit-exact.
    python tools/fuzz_parity.py [n_cases] [seed]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native
from oracle import port

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    c, exp = [(8, 1), (17, 2), (25, 3), (5, 1), (25, 3), (17, 2)][rng.randint(6)]
    n = int(rng.randint(1, 5))
    h = int(rng.choice([16, 33, 48, 64, 100, 135, 270]))
    w = int(rng.choice([16, 47, 64, 96, 160, 240, 480]))
    style = rng.randint(4)
    g = torch.Generator().manual_seed(int(rng.randint(1 << 30)))
    hi = c + 1 if exp != 1 else c
    if style == 0:                                            # iid
        x = torch.randn((n, c, h, w), generator=g) * float(rng.choice([0.3, 1.0, 3.0]))
        y = torch.randint(0, hi, (n, h, w), generator=g)
    elif style == 1:                                          # confident, blocky
        coarse = torch.randint(0, hi, (n, (h + 7) // 8, (w + 7) // 8), generator=g)
        y = coarse.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :h, :w].contiguous()
        x = 5.0 * torch.nn.functional.one_hot(y.clamp(max=c - 1), c).permute(0, 3, 1, 2).float() + torch.randn((n, c, h, w), generator=g)
    elif style == 2:                                          # heavy ties: logits on a coarse grid
        x = torch.round(torch.randn((n, c, h, w), generator=g) * 2) / 2
        y = torch.randint(0, hi, (n, h, w), generator=g)
    else:                                                     # few classes present, one image all-ignore when possible
        y = torch.randint(0, min(3, c), (n, h, w), generator=g)
        if exp != 1 and n > 1:
            y[0] = c
        x = torch.randn((n, c, h, w), generator=g)
    cfg = {"experiment": exp}
    kw = {}
    if rng.rand() < 0.4:
        cfg["per_image"] = kw["per_image"] = True
    mode = rng.randint(4)
    if mode == 1 and exp != 1:
        cfg["classes_to_ignore"] = kw["classes_to_ignore"] = c
    elif mode == 2:
        cfg["classes_to_consider"] = kw["classes_to_consider"] = "all"
    elif mode == 3:
        lst = sorted(set(int(v) for v in rng.randint(0, c, size=3)))
        cfg["classes_to_consider"] = kw["classes_to_consider"] = lst
    ldt = [torch.int64, torch.int32, torch.uint8][rng.randint(3)]
    _native.set_tuning(emit_path=int(rng.randint(3)), interleave=int(rng.randint(2)), stats_variant=int(rng.choice([0, 0, 1, 2, 6, 7])),
                       sort_path=int(rng.choice([0, 0, 0, 1])))
    xd = x.cuda().requires_grad_(True)
    yd = y.cuda().to(ldt)
    tag = f"case {case}: C={c} n={n} {h}x{w} style={style} cfg={cfg} labels={ldt}"
    try:
        if c in (8, 17, 25):
            meter = b200.SegmentationMeter(exp, c)
            loss = b200.LovaszSoftmaxWithMetrics(cfg, meter)(xd, yd)
        else:
            meter = None
            loss = b200.LovaszSoftmax(cfg)(xd, yd)
        loss.backward()
        filt = kw.get("classes_to_ignore")
        dead = [bool((y[i] == filt).all()) for i in range(n)] if (kw.get("per_image") and filt is not None) else []
        if any(dead):
            # the reference returns an empty tensor here; ours counts a fully filtered image as 0 in the mean
            kw1 = {k: v for k, v in kw.items() if k != "per_image"}
            ref_loss, ref_grad = 0.0, torch.zeros_like(x, device="cuda")
            for i in range(n):
                if not dead[i]:
                    li, gi = port.lovasz_softmax_with_grad(x[i:i + 1].cuda(), y[i:i + 1].cuda(), exp, **kw1)
                    ref_loss += float(li) / n
                    ref_grad[i] = gi[0] / n
        else:
            ref_loss, ref_grad = port.lovasz_softmax_with_grad(x.cuda(), y.cuda(), exp, **kw)
        loss = loss.detach()
        lerr = abs(float(loss) - float(ref_loss)) / max(abs(float(ref_loss)), 1e-12) if float(ref_loss) != 0 else abs(float(loss))
        gmax = float(ref_grad.abs().max())
        gerr = float((xd.grad - ref_grad).abs().max()) / gmax if gmax > 0 else float(xd.grad.abs().max())
        ok = lerr <= 1e-5 and gerr <= 1e-5
        if meter is not None:
            meter.check()
            ok = ok and torch.equal(meter.cm.cpu(), port.confusion_matrix(x, y.int()).to(torch.int64))
        if not ok:
            bad += 1
            print("MISMATCH", tag, "loss err", lerr, "grad err", gerr)
    except Exception as e:                                    # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e)[:300])
_native.set_tuning(emit_path=0, interleave=1, stats_variant=0, sort_path=0)

# ---- the widened rows: fused CE + Lovasz pair, OHEM cross entropy, windowed IoU map ------------------------------
for case in range(n_cases):
    c, exp = [(8, 1), (17, 2), (25, 3), (6, None)][rng.randint(4)]
    n = int(rng.randint(1, 4))
    h = int(rng.choice([16, 33, 48, 64, 100]))
    w = int(rng.choice([16, 47, 64, 96, 160]))
    g = torch.Generator().manual_seed(int(rng.randint(1 << 30)))
    hi = c + 1 if exp in (2, 3) else c
    y = torch.randint(0, hi, (n, h, w), generator=g)
    x = torch.randn((n, c, h, w), generator=g) * float(rng.choice([0.5, 1.0, 2.0]))
    if rng.rand() < 0.5:
        x = x + 4.0 * torch.nn.functional.one_hot(y.clamp(max=c - 1), c).permute(0, 3, 1, 2).float() * \
            (torch.rand((n, 1, h, w), generator=g) < 0.7).float()
    ldt = [torch.int64, torch.int32, torch.uint8][rng.randint(3)]
    xd, yd = x.cuda(), y.cuda().to(ldt)
    tag = f"widened case {case}: C={c} exp={exp} n={n} {h}x{w} labels={ldt}"
    try:
        # OHEM
        cfg = {"min_kept": int(rng.choice([1, 50, 500, 5000, 100000])), "thresh": float(rng.choice([0.05, 0.3, 0.7, 0.95]))}
        if exp is not None:
            cfg["experiment"] = exp
        mod = b200.OhemCrossEntropy(cfg)
        xo = xd.clone().requires_grad_(True)
        lo = mod(xo, yd)
        ref_l, ref_g = port.ohem_with_grad(xd, y.cuda(), thresh=mod.thresh, min_kept=mod.min_kept, ignore_label=mod.ignore_label)
        if bool(torch.isnan(ref_l)):
            ok = bool(torch.isnan(lo))
        else:
            lo.backward()
            ok = abs(float(lo.detach()) - float(ref_l)) <= 1e-5 * abs(float(ref_l)) and \
                float((xo.grad - ref_g).abs().max()) <= 1e-5 * float(ref_g.abs().max())
        if not ok:
            bad += 1
            print("MISMATCH ohem", tag, cfg, float(lo.detach()), float(ref_l))
        # windowed IoU (labels must be real classes)
        k, s = int(rng.choice([1, 3, 5, 7, 9])), int(rng.choice([1, 2, 4, 5]))
        yy = y.clamp(max=c - 1)
        full = bool(rng.randint(2))
        got = b200.sliding_miou(xd, yy.cuda().to(ldt), k, s, original_size=full)
        ref = port.sliding_miou(x, yy, k, s, original_size=full)
        if tuple(got.shape) != tuple(ref.shape) or float((got.cpu() - ref).abs().max()) > 1e-6:
            bad += 1
            print("MISMATCH sliding", tag, k, s, full)
        # fused CE + Lovasz pair
        if exp is not None:
            xp = xd.clone().requires_grad_(True)
            lov, ce = b200.LovaszSoftmaxCE({"experiment": exp})(xp, yd)
            (lov + 0.5 * ce).backward()
            xr = xd.clone().requires_grad_(True)
            rt = port.loss_wrapper_pair(xr, y.cuda(), exp, 0.5, 1.0)
            rt.backward()
            got_t = float(lov.detach()) + 0.5 * float(ce.detach())
            ok = abs(got_t - float(rt)) <= 1e-5 * abs(float(rt)) and \
                float((xp.grad - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())
            if not ok:
                bad += 1
                print("MISMATCH pair", tag, got_t, float(rt))
    except Exception as e:                                    # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e)[:300])
print(f"2 x {n_cases} cases, {bad} bad")
sys.exit(1 if bad else 0)
