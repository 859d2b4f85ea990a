"""Hybrid sort path (sort_path=0) against the three-pass LSD path (sort_path=1): bit equality of loss / gradients on a
set of shapes and distributions, then in-stream stage times of both.   python tools/hybrid_check.py [--quick]"""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import miccai2021_cataract_semantic_segmentation_b200 as b200  # noqa: E402
from miccai2021_cataract_semantic_segmentation_b200 import _native  # noqa: E402
from test_gpu_parity import _blocky, _d1  # noqa: E402


def run(x, y, cfg):
    xd = x.requires_grad_(True)
    xd.grad = None
    loss = b200.LovaszSoftmax(cfg)(xd, y)
    loss.backward()
    return float(loss), xd.grad.clone()


def compare(name, x, y, cfg):
    x, y = x.cuda(), y.cuda()
    _native.set_tuning(sort_path=1)
    l1, g1 = run(x.clone(), y, cfg)
    _native.set_tuning(sort_path=0)
    l0, g0 = run(x.clone(), y, cfg)
    torch.cuda.synchronize()
    same = l0 == l1 and torch.equal(g0, g1)
    nd = int((g0 != g1).sum())
    print(f"{'OK ' if same else 'BAD'} {name}: loss {l0!r} vs {l1!r}, differing grads {nd}, max {float((g0 - g1).abs().max()):.3e}",
          flush=True)
    return same


def main():
    lib = _native.load()
    ok = True
    cases = [
        ("d1 c8 flat small", _d1(2, 8, 64, 96, 1, False), {"experiment": 1}),
        ("d1 c25 flat small", _d1(2, 25, 96, 160, 2, True), {"experiment": 3}),
        ("d1 c17 per-image", _d1(3, 17, 128, 160, 3, True), {"experiment": 2, "per_image": True}),
        ("d1 c25 all", _d1(2, 25, 96, 160, 4, True), {"experiment": 3, "classes_to_consider": "all"}),
        ("blocky c25 flat", _blocky(2, 25, 128, 160, 5, True), {"experiment": 3}),
        ("blocky c17 per-image ignore", _blocky(2, 17, 128, 160, 6, True), {"experiment": 2, "per_image": True, "classes_to_ignore": 17}),
        ("zeros (all ties) small", (torch.zeros(1, 25, 64, 96), torch.randint(0, 26, (1, 64, 96))), {"experiment": 3}),
        ("zeros (all ties) overflow", (torch.zeros(1, 25, 160, 160), torch.randint(0, 26, (1, 160, 160))), {"experiment": 3}),
        ("grid 0.5 ties", ((torch.randn(2, 25, 96, 160) * 2).round() / 2, torch.randint(0, 26, (2, 96, 160))), {"experiment": 3}),
        ("d1 c25 flat medium", _d1(2, 25, 540, 960, 7, True), {"experiment": 3}),
        ("blocky c25 flat medium", _blocky(2, 25, 540, 960, 8, True), {"experiment": 3}),
    ]
    if "--quick" not in sys.argv:
        cases += [
            ("d1 c25 flat full", _d1(8, 25, 540, 960, 9, True), {"experiment": 3}),
            ("d1 c17 per-image full", _d1(8, 17, 540, 960, 10, True), {"experiment": 2, "per_image": True}),
            ("blocky c25 flat full", _blocky(8, 25, 540, 960, 11, True), {"experiment": 3}),
            ("d1 c8 flat full", _d1(8, 8, 540, 960, 12, False), {"experiment": 1}),
        ]
    for name, (x, y), cfg in cases:
        ok &= compare(name, x, y, cfg)

    # stage times
    names = {0: ["stats", "finalize", "emit", "prepare", "hyb_count", "hyb_partition", "hyb_local", "fallback", None, "backward"],
             1: ["stats", "finalize", "emit", "prepare", "pass0", "pass1", "pass2+fgcount", "jaccard", None, "backward"]}
    n_ev = 11
    for tag, (x, y), cfg in [("d1 c25 flat full", _d1(8, 25, 540, 960, 9, True), {"experiment": 3}),
                             ("blocky c25 flat full", _blocky(8, 25, 540, 960, 11, True), {"experiment": 3}),
                             ("d1 c17 per-image full", _d1(8, 17, 540, 960, 10, True), {"experiment": 2, "per_image": True})]:
        x, y = x.cuda(), y.cuda()
        for path in (1, 0):
            _native.set_tuning(sort_path=path)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
            for ev in evs:
                ev.record()
            torch.cuda.synchronize()
            arr = (ctypes.c_void_p * n_ev)(*[ctypes.c_void_p(ev.cuda_event) for ev in evs])
            mod = b200.LovaszSoftmax(cfg)
            xr = x.clone().requires_grad_(True)
            for _ in range(3):
                xr.grad = None
                mod(xr, y).backward()
            _native.check(lib.b200seg_set_stage_events(arr, n_ev), "events")
            acc = [0.0] * 10
            reps = 10
            tot = 0.0
            for _ in range(reps):
                xr.grad = None
                mod(xr, y).backward()
                torch.cuda.synchronize()
                for i in range(10):
                    if names[path][i]:
                        acc[i] += evs[i].elapsed_time(evs[i + 1])
                tot += evs[0].elapsed_time(evs[8]) + evs[9].elapsed_time(evs[10])
            _native.check(lib.b200seg_set_stage_events(None, 0), "events off")
            print(f"[{tag}] sort_path={path} total {tot / reps * 1e3:.0f} us:",
                  {names[path][i]: round(acc[i] / reps * 1e3, 1) for i in range(10) if names[path][i]}, flush=True)
    print("ALL OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
