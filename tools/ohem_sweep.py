#!/usr/bin/env python
"""Time OhemCrossEntropy forward + backward on the BASELINE-size batch (8 x 25 x 544 x 960), next to the oracle's torch
restatement (softmax + CE + gather + kthvalue + masked mean, autograd backward) on the same device.
    python tools/ohem_sweep.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miccai2021_cataract_semantic_segmentation_b200 as b200
from oracle import port

n, c, h, w = 8, 25, 544, 960
peak = 6650.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
g = torch.Generator(device="cuda").manual_seed(0)
y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for style in ("iid", "confident"):
    x = torch.randn((n, c, h, w), generator=g, device="cuda")
    if style == "confident":
        x += 4.0 * torch.nn.functional.one_hot(y.clamp(max=c - 1), c).permute(0, 3, 1, 2) * \
            (torch.rand((n, 1, h, w), generator=g, device="cuda") < 0.8)
    x.requires_grad_(True)
    mod = b200.OhemCrossEntropy({"experiment": 3})

    def step():
        x.grad = None
        mod(x, y).backward()

    def fwd():
        with torch.no_grad():
            mod(x, y)

    def ref_step():
        x.grad = None
        port.ohem_cross_entropy(x, y, thresh=mod.thresh, min_kept=mod.min_kept, ignore_label=mod.ignore_label).backward()

    ms, ms_f, ms_ref = timed(step), timed(fwd), timed(ref_step, reps=5)
    import time
    xc, yc = x[:1].detach().cpu().requires_grad_(True), y[:1].cpu()

    def cpu_step():
        xc.grad = None
        port.ohem_cross_entropy(xc, yc, thresh=mod.thresh, min_kept=mod.min_kept, ignore_label=mod.ignore_label).backward()

    cpu_step()
    t0 = time.perf_counter()
    for _ in range(3):
        cpu_step()
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    kept = int((x.grad.abs().sum(1) > 0).sum())
    alg = (8 * c + 8) * n * h * w                 # logits read twice... once per pass, dlogits written, labels
    print(json.dumps({"op": "ohem fwd+bwd", "logits": style, "ms": round(ms, 4), "fwd_ms": round(ms_f, 4),
                      "kept_fraction": round(kept / (n * h * w), 3), "Mpx_per_s": round(n * h * w / ms / 1e3, 1),
                      "torch_restatement_ms": round(ms_ref, 3), "copy_peak_GBps": peak,
                      "cpu_oracle_1_image_ms": round(cpu_ms, 1), "cpu_threads": torch.get_num_threads(),
                      "cpu_Mpx_per_s": round(h * w / cpu_ms / 1e3, 2)}))
