"""stats kernel fed by TMA (stats_variant=7: one cp.async.bulk.tensor.3d per 128-pixel x C box) against the per-warp cp.async
(LDGSTS) ring (stats_variant=0): bit equality of everything downstream, then the in-stream time of the stats stage."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native
from test_gpu_parity import _d1
lib = _native.load()
for (n, c, h, w, exp, cfg) in [(2, 25, 96, 160, 3, {}), (3, 17, 128, 192, 2, {"per_image": True}), (2, 8, 64, 96, 1, {}), (8, 25, 540, 960, 3, {})]:
    x, y = _d1(n, c, h, w, 5, exp != 1)
    x, y = x.cuda(), y.cuda()
    res = []
    for variant in (0, 7):
        _native.set_tuning(stats_variant=variant)
        meter = b200.SegmentationMeter(exp, c)
        xr = x.clone().requires_grad_(True)
        loss = b200.LovaszSoftmaxWithMetrics({"experiment": exp, **cfg}, meter)(xr, y)
        loss.backward()
        torch.cuda.synchronize()
        res.append((float(loss.detach()), xr.grad.clone(), meter.cm.clone()))
    same = res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    print(("OK " if same else "BAD"), (n, c, h, w), cfg, res[0][0], res[1][0], flush=True)
x, y = _d1(8, 25, 540, 960, 9, True)
x, y = x.cuda(), y.cuda()
for variant in (0, 7, 0, 7):
    _native.set_tuning(stats_variant=variant)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    for ev in evs: ev.record()
    torch.cuda.synchronize()
    arr = (ctypes.c_void_p * 11)(*[ctypes.c_void_p(ev.cuda_event) for ev in evs])
    meter = b200.SegmentationMeter(3, 25)
    mod = b200.LovaszSoftmaxWithMetrics({"experiment": 3}, meter)
    with torch.no_grad():
        for _ in range(3): mod(x, y)
        _native.check(lib.b200seg_set_stage_events(arr, 11), "ev")
        t = 0.0
        for _ in range(20):
            mod(x, y); torch.cuda.synchronize()
            t += evs[0].elapsed_time(evs[1])
        _native.check(lib.b200seg_set_stage_events(None, 0), "ev")
    print(f"stats stage, 8x25x540x960, {'TMA (cp.async.bulk.tensor.3d)' if variant == 7 else 'cp.async rings (LDGSTS)'}: {t / 20 * 1e3:.1f} us", flush=True)
_native.set_tuning(stats_variant=0)
