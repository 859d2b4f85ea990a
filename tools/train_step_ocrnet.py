#!/usr/bin/env python
"""BASELINE configs[3]: one OCRNet-R50 (random init) training step through the reference's own LossWrapper / TwoScaleLoss
and manager-style metrics, first with the UNMODIFIED reference on the GPU, then after install() of the B200 drop-ins;
same model, same batch.  Reports parity (loss, d loss / d logits, parameter gradients) and the time of the loss + metrics
path and of the whole step, with and without the drop-in.

    python tools/train_step_ocrnet.py [--batch 8] [--height 540] [--width 960] [--steps 10] [--loss wrapper|twoscale]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step_ocrnet.py ...   (DDP)

The reference packages come from $B200SEG_REFERENCE, /root/reference or baseline/_ref (tools/stage_reference.sh).
reference: managers/OCRNet_Manager.py:67-134 (step), losses/LossWrapper.py:43-73, losses/TwoScaleLoss.py:43-52, main.py:8.
"""
import argparse
import json
import os
import sys
import time
import warnings
from unittest.mock import MagicMock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_root():
    for cand in (os.environ.get("B200SEG_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "losses")) and os.path.isdir(os.path.join(cand, "models")):
            return cand
    return None


def import_reference(root):
    for m in ("matplotlib", "matplotlib.colors", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py", "ttach"):
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, root)
    warnings.simplefilter("ignore")
    import utils, losses, models, managers   # noqa: F401
    import managers.OCRNet_Manager            # noqa: F401
    ls = sys.modules["losses.LovaszSoftmax"]  # canonical tie order for the comparison

    class _StableTorch:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def sort(input, dim=-1, descending=False):
            return torch.sort(input, dim=dim, descending=descending, stable=True)
    ls.torch = _StableTorch()
    return utils, losses, models


def blocky_labels(n, c, h, w, gen, device):
    coarse = torch.randint(0, c + 1, (n, (h + 15) // 16, (w + 15) // 16), generator=gen, device=device)
    return coarse.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :h, :w].contiguous()


def make_loss(losses, kind, exp, device):
    if kind == "wrapper":       # BASELINE configs[3]
        return losses.LossWrapper({"losses": {"CrossEntropyLoss": 1, "LovaszSoftmax": 1}, "experiment": exp, "device": device})
    return losses.TwoScaleLoss({"interm": {"name": "LovaszSoftmax", "args": [], "weight": 0.4},      # configs/OCRNet_rf_lvsz.json
                                "final": {"name": "LovaszSoftmax", "args": [], "weight": 1.0}, "experiment": exp})


def loss_call(loss_fn, kind, interm, final, lbl):
    if kind == "wrapper":
        return loss_fn(None, final, lbl.long(), interm_prediction=interm)
    return loss_fn(interm, final, lbl.long())


def metrics_call(utils_mod, final, lbl, exp):
    """what managers/OCRNet_Manager.py:108-113 does every training step"""
    cm = utils_mod.t_get_confusion_matrix(final, lbl)
    pa, pac = utils_mod.t_get_pixel_accuracy(cm)
    iou = utils_mod.t_get_mean_iou(cm, exp, categories=True, calculate_mean=False, rare=True)
    return cm, pa, pac, iou


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--loss", default="wrapper", choices=["wrapper", "twoscale"])
    ap.add_argument("--experiment", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    root = reference_root()
    if root is None:
        print(json.dumps({"unavailable": "no reference tree (run tools/stage_reference.sh where /root/reference exists)"}))
        return 0
    from miccai2021_cataract_semantic_segmentation_b200 import dist as bdist
    import torch.distributed as dist
    rank, world, local = bdist.init_from_env()
    device = torch.device("cuda", local)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.autograd.set_detect_anomaly(True)               # the reference runs with anomaly mode on (main.py:8)
    utils_mod, losses, models = import_reference(root)
    import miccai2021_cataract_semantic_segmentation_b200 as b200
    exp = args.experiment
    c = {1: 8, 2: 17, 3: 25}[exp]
    torch.manual_seed(0)
    model = models.OCRNet({"backbone": "resnet50", "pretrained": False, "out_stride": 8}, exp).to(device).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    gen = torch.Generator(device=device).manual_seed(100 + rank)
    img = torch.randn((args.batch, 3, args.height, args.width), generator=gen, device=device)
    lbl = blocky_labels(args.batch, c, args.height, args.width, gen, device).int()     # the loader hands int32 labels

    # ---- reference loss / metrics objects BEFORE install(), drop-ins after ----------------------------------------
    ref_loss = make_loss(losses, args.loss, exp, device).to(device)
    ref_metrics = {k: getattr(utils_mod, k) for k in ("t_get_confusion_matrix", "t_get_pixel_accuracy", "t_get_mean_iou")}
    rebound = b200.install(fuse_ce=True, two_stream_heads=True)
    new_loss = make_loss(losses, args.loss, exp, device).to(device)
    assert type(new_loss) is not type(ref_loss) or args.loss == "twoscale"
    mgr = sys.modules["managers.OCRNet_Manager"]

    class _RefUtils:            # the names the manager module bound before install()
        t_get_confusion_matrix = staticmethod(ref_metrics["t_get_confusion_matrix"])
        t_get_pixel_accuracy = staticmethod(ref_metrics["t_get_pixel_accuracy"])
        t_get_mean_iou = staticmethod(ref_metrics["t_get_mean_iou"])

    # ---- parity on one step: same logits, both loss implementations ---------------------------------------------------
    out = {}
    res = {}
    for tag, loss_fn in (("reference", ref_loss), ("reference_again", ref_loss), ("b200", new_loss)):
        for p_ in model.parameters():
            p_.grad = None
        interm, final = net(img)                          # a fresh forward per variant (DDP reduces gradients once per forward)
        final.retain_grad()
        interm.retain_grad()
        loss = loss_call(loss_fn, args.loss, interm, final, lbl)
        loss.backward()
        pg = torch.cat([p_.grad.flatten() for p_ in model.parameters() if p_.grad is not None])
        res[tag] = (float(loss.detach()), final.grad.clone(), None if interm.grad is None else interm.grad.clone(), pg.clone())
    lr, lb = res["reference"][0], res["b200"][0]
    rel = lambda a_, b_: float((a_ - b_).abs().max() / b_.abs().max())
    out["loss_reference"], out["loss_b200"], out["loss_rel_err"] = lr, lb, abs(lr - lb) / abs(lr)
    out["dlogits_rel_err"] = rel(res["b200"][1], res["reference"][1])
    if res["reference"][2] is not None:
        out["dlogits_interm_rel_err"] = rel(res["b200"][2], res["reference"][2])
    out["param_grad_rel_err"] = rel(res["b200"][3], res["reference"][3])
    # noise floor of the comparison: the SAME reference loss evaluated twice (atomics in the upsampling / cuDNN backward)
    out["param_grad_noise_floor"] = rel(res["reference_again"][3], res["reference"][3])
    out["dlogits_noise_floor"] = rel(res["reference_again"][1], res["reference"][1])
    cm_r, pa_r, pac_r, iou_r = metrics_call(_RefUtils, final.detach(), lbl, exp)
    cm_n, pa_n, pac_n, iou_n = metrics_call(mgr, final.detach(), lbl, exp)          # the manager's (re-bound) names
    out["confusion_matrix_equal"] = bool(torch.equal(cm_r.long(), cm_n.long()))
    out["iou_max_abs_err"] = float(max((a_ - b_).abs().max() for a_, b_ in zip(iou_r, iou_n)))
    final = final.detach()
    del res, interm
    torch.cuda.empty_cache()

    # ---- timing: loss + metrics path, and the whole step, with each implementation ------------------------------------
    opt = torch.optim.SGD(net.parameters(), lr=1e-4)

    def step(loss_fn, util_ns, detail):
        opt.zero_grad(set_to_none=True)
        t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        interm_, final_ = net(img)
        t[0].record()
        loss = loss_call(loss_fn, args.loss, interm_, final_, lbl)
        t[1].record()
        loss.backward()
        t[2].record()
        opt.step()
        metrics_call(util_ns, final_.detach(), lbl, exp)
        t[3].record()
        if detail:
            torch.cuda.synchronize()
            return t[0].elapsed_time(t[1]), t[1].elapsed_time(t[2]), t[2].elapsed_time(t[3])
        return None

    timing = {}
    for tag, loss_fn, util_ns in (("reference", ref_loss, _RefUtils), ("b200", new_loss, mgr)):
        for _ in range(2):
            step(loss_fn, util_ns, False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(loss_fn, util_ns, False)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        lf, bw, mt = step(loss_fn, util_ns, True)
        timing[tag] = {"step_ms": e0.elapsed_time(e1) / args.steps, "wall_ms": wall, "loss_forward_ms": lf,
                       "backward_all_ms": bw, "optimizer+metrics_ms": mt}
    out["timing"] = timing

    # ---- the same step with the model's final F.interpolate (models/OCR.py:126-131) folded into the loss kernels ----
    # The model stops at its stride-8 logits (the six calls below are OCR.py:110-125); CE + Lovasz + confusion matrix are
    # computed from them by lovasz_softmax_upsampled, whose backward returns the gradient at stride 8.
    if world == 1 and args.loss == "wrapper":
        from miccai2021_cataract_semantic_segmentation_b200.fused import ce_ignore_index

        def forward_low(m, x):
            feats = m.backbone(x)
            interm_ = m.interm_prediction_head(feats["low"])
            hi = m.conv_high_map(feats["high"])
            ctx = m.spatial_gather(hi, interm_)
            return m.conv_out(m.spatial_ocr_head(hi, ctx))

        ign = ce_ignore_index(exp)
        meter = b200.SegmentationMeter(exp, c, device)
        lbl64 = lbl.long()
        # parity of the fused step against the reference loss on the upsampled logits (same weights, same batch)
        for p_ in model.parameters():
            p_.grad = None
        low = forward_low(model, img)
        low.retain_grad()
        lov, ce = b200.lovasz_softmax_upsampled(low, lbl64, ce_ignore_index=ign, confusion=meter.cm,
                                                confusion_drop_label=meter.drop_label, status=meter.status)
        (lov + ce).backward()
        g_fused = low.grad.clone()
        for p_ in model.parameters():
            p_.grad = None
        low2 = forward_low(model, img)
        low2.retain_grad()
        full = torch.nn.functional.interpolate(low2, size=img.shape[-2:], mode="bilinear", align_corners=True)
        ref_total = loss_call(ref_loss, args.loss, None, full, lbl)
        ref_total.backward()
        cm_ref = _RefUtils.t_get_confusion_matrix(full.detach(), lbl)
        out["fused_upsample"] = {
            "loss_rel_err": abs(float((lov + ce).detach()) - float(ref_total.detach())) / abs(float(ref_total.detach())),
            "dlowres_rel_err": rel(g_fused, low2.grad),
            "confusion_matrix_equal": bool(torch.equal(meter.cm, cm_ref.long())),
            "low_res": list(low.shape[-2:])}
        del low, low2, full, g_fused
        torch.cuda.empty_cache()

        def step_fused(detail):
            opt.zero_grad(set_to_none=True)
            t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            meter.reset()
            low_ = forward_low(model, img)
            t[0].record()
            lov_, ce_ = b200.lovasz_softmax_upsampled(low_, lbl64, ce_ignore_index=ign, confusion=meter.cm,
                                                      confusion_drop_label=meter.drop_label, status=meter.status)
            total = lov_ + ce_
            t[1].record()
            total.backward()
            t[2].record()
            opt.step()
            meter.summary()
            t[3].record()
            if detail:
                torch.cuda.synchronize()
                return t[0].elapsed_time(t[1]), t[1].elapsed_time(t[2]), t[2].elapsed_time(t[3])
            return None

        for _ in range(2):
            step_fused(False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_fused(False)
        e1.record()
        torch.cuda.synchronize()
        lf, bw, mt = step_fused(True)
        timing["b200_fused_upsample"] = {"step_ms": e0.elapsed_time(e1) / args.steps, "loss_forward_ms": lf,
                                         "backward_all_ms": bw, "optimizer+metrics_ms": mt,
                                         "note": "model stops at stride-8 logits; the intermediate head's unused upsampling "
                                                 "(LossWrapper ignores interm_prediction) is not computed either"}
    out["config"] = {"model": "OCRNet resnet50 random init out_stride 8", "loss": args.loss, "experiment": exp, "batch_per_gpu": args.batch,
                     "size": [args.height, args.width], "world": world, "ddp": world > 1, "anomaly_mode": True,
                     "reference_root": os.path.basename(root.rstrip("/")), "rebound_modules": sorted(rebound)}
    if world > 1:
        vals = torch.tensor([out["loss_rel_err"], out["dlogits_rel_err"], out["param_grad_rel_err"]], device=device, dtype=torch.float64)
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        out["max_over_ranks"] = {"loss_rel_err": float(vals[0]), "dlogits_rel_err": float(vals[1]), "param_grad_rel_err": float(vals[2])}
    if rank == 0:
        line = json.dumps(out)
        print(line)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
