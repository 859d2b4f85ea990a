#!/usr/bin/env python
"""Bring-up diagnostics on a GPU box: small parity checks with verbose output, then a per-kernel time breakdown of
the full-size step.  Not a test and not the benchmark -- a debugging aid (`gpurun -- python tools/first_light.py`)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import miccai2021_cataract_semantic_segmentation_b200 as b200  # noqa: E402
from oracle import port  # noqa: E402

SMALL = "--small" in sys.argv


def check_sort():
    from test_gpu_parity import _sort_segments
    rng = np.random.RandomState(0)
    cap = 20000
    counts = np.array([0, 1, 17, 4096, 4097, 20000, 12345, 2], dtype=np.int64)
    bits = np.array([1, 5, 3, 30, 25, 24, 10, 30], dtype=np.int64)
    keys = np.zeros(len(counts) * cap, dtype=np.int64)
    vals = np.zeros_like(keys)
    for s, n in enumerate(counts):
        keys[s * cap:s * cap + n] = rng.randint(0, 1 << bits[s], size=n)
        vals[s * cap:s * cap + n] = np.arange(n)
    kout, vout = _sort_segments(keys, vals, counts, bits, cap)
    ok = True
    for s, n in enumerate(counts):
        sl = slice(s * cap, s * cap + n)
        order = np.argsort(keys[sl], kind="stable")
        kk, vv = np.array_equal(kout[sl], keys[sl][order]), np.array_equal(vout[sl], vals[sl][order])
        print(f"  sort seg {s}: n={n} bits={bits[s]} keys_ok={kk} vals_ok={vv}")
        if not (kk and vv):
            ok = False
            bad = np.nonzero(kout[sl] != keys[sl][order])[0]
            print("    first bad positions:", bad[:10], "got", kout[sl][bad[:5]], "want", keys[sl][order][bad[:5]])
    return ok


def check_lovasz(n, c, h, w, exp, **cfg):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, c, h, w), generator=g)
    y = torch.randint(0, c + (exp != 1), (n, h, w), generator=g)
    kw = dict(per_image=cfg.get("per_image", False), classes_to_ignore=cfg.get("classes_to_ignore"),
              classes_to_consider=cfg.get("classes_to_consider", "present"))
    rl, rg = port.lovasz_softmax_with_grad(x, y, exp, **kw)
    xd = x.cuda().requires_grad_(True)
    loss = b200.LovaszSoftmax({"experiment": exp, **cfg})(xd, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    lerr = abs(float(loss.detach()) - float(rl)) / abs(float(rl))
    gerr = float((xd.grad.cpu() - rg).abs().max() / rg.abs().max())
    dl, dg = port.lovasz_softmax_with_grad(x.cuda(), y.cuda(), exp, **kw)
    gerr_dev = float((xd.grad - dg).abs().max() / dg.abs().max())
    nbad = int(((xd.grad.cpu() - rg).abs() > 1e-5 * rg.abs().max()).sum())
    print(f"     vs oracle-on-GPU: loss rel {abs(float(loss.detach()) - float(dl)) / abs(float(dl)):.2e} grad err {gerr_dev:.2e};"
          f" vs CPU: {nbad} elements above 1e-5")
    ref_cm = port.confusion_matrix(x, y.int()).to(torch.int64)
    cm = b200.t_get_confusion_matrix(x.cuda(), y.cuda().int())
    cm_ok = torch.equal(cm.cpu(), ref_cm)
    print(f"  lovasz {n}x{c}x{h}x{w} {cfg}: loss {float(loss):.8f} ref {float(rl):.8f} rel {lerr:.2e} | grad err {gerr:.2e}"
          f" | cm_ok={cm_ok}")
    return lerr <= 1e-5 and gerr_dev <= 1e-5 and cm_ok


def time_call(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def full_size(c, exp, per_image, blocky=False):
    n, h, w = 8, 540, 960
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, c, h, w), generator=g, device="cuda")
    y = torch.randint(0, c + (exp != 1), (n, h, w), generator=g, device="cuda")
    if blocky:
        coarse = torch.randint(0, c // 2, (n, h // 20, w // 20), generator=g, device="cuda")
        y = coarse.repeat_interleave(20, 1).repeat_interleave(20, 2).contiguous()
        x = x + 6.0 * torch.nn.functional.one_hot(y, c).permute(0, 3, 1, 2).float()
    meter = b200.SegmentationMeter(exp, c)
    fused = b200.LovaszSoftmaxWithMetrics({"experiment": exp, "per_image": per_image}, meter)
    xr = x.clone().requires_grad_(True)

    def step():
        xr.grad = None
        loss = fused(xr, y)
        loss.backward()
        return meter.summary()

    def fwd_only():
        with torch.no_grad():
            return fused(x, y)

    def cm_only():
        meter.update(x, y)

    px = n * h * w
    tag = f"C={c} per_image={per_image} blocky={blocky}"
    med, best = time_call(step)
    print(f"  [{tag}] fwd+bwd+cm+miou: median {med:.3f} ms best {best:.3f} ms -> {px / med / 1e3:.0f} Mpx/s, "
          f"{px * (8 * c + 8) / med / 1e6:.0f} GB/s algorithmic")
    med, best = time_call(fwd_only)
    print(f"  [{tag}] forward only   : median {med:.3f} ms best {best:.3f} ms")
    med, best = time_call(cm_only)
    print(f"  [{tag}] confmat only   : median {med:.3f} ms best {best:.3f} ms -> {px * (4 * c + 8) / med / 1e6:.0f} GB/s")
    meter.check()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / 3.0, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"  [{tag}] kernel breakdown per step (total {tot:.0f} us):")
    for k, t, cnt in rows[:16]:
        print(f"      {t:9.1f} us  x{cnt:<3d} {k[:90]}")


def main():
    print("device:", torch.cuda.get_device_name(0))
    t0 = time.time()
    ok = True
    print("== sort"); ok &= check_sort()
    print("== lovasz small")
    ok &= check_lovasz(1, 8, 16, 32, 1)
    ok &= check_lovasz(2, 25, 36, 64, 3)
    ok &= check_lovasz(2, 17, 36, 64, 2, per_image=True)
    ok &= check_lovasz(2, 25, 37, 63, 3)
    ok &= check_lovasz(2, 17, 36, 64, 2, classes_to_ignore=17)
    ok &= check_lovasz(2, 17, 36, 64, 2, classes_to_consider="all")
    ok &= check_lovasz(2, 25, 135, 240, 3)
    print("small checks:", "ALL OK" if ok else "FAILURES", f"({time.time() - t0:.1f}s)")
    if SMALL:
        return
    print("== full size")
    full_size(25, 3, False)
    full_size(17, 2, True)
    full_size(25, 3, False, blocky=True)
    full_size(8, 1, False)


if __name__ == "__main__":
    main()
