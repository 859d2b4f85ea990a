import ctypes, os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native
lib = _native.load()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((1, 25, 544, 960), generator=g, device="cuda"); y = torch.randint(0, 26, (1, 544, 960), generator=g, device="cuda")
names = ["stats", "finalize", "emit", "prepare", "hyb_count", "hyb_partition", "hyb_local", "fallback"]
evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
for ev in evs: ev.record()
torch.cuda.synchronize()
arr = (ctypes.c_void_p * 11)(*[ctypes.c_void_p(ev.cuda_event) for ev in evs])
meter = b200.SegmentationMeter(3, 25)
mod = b200.LovaszSoftmaxWithMetrics({"experiment": 3}, meter)
with torch.no_grad():
    for _ in range(5): mod(x, y)
    _native.check(lib.b200seg_set_stage_events(arr, 11), "ev")
    acc = [0.0] * 8
    for _ in range(20):
        mod(x, y); torch.cuda.synchronize()
        for i in range(8): acc[i] += evs[i].elapsed_time(evs[i + 1])
    _native.check(lib.b200seg_set_stage_events(None, 0), "ev")
print({names[i]: round(acc[i] * 50, 1) for i in range(8)}, "sum", round(sum(acc) * 50, 1))
