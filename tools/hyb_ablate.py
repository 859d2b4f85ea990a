"""Timing ablations of the hybrid local kernel (results are wrong with dbg != 0)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native
from test_gpu_parity import _d1
lib = _native.load()
x, y = _d1(8, 25, 540, 960, 9, True)
x, y = x.cuda(), y.cuda()
names = ["stats", "finalize", "emit", "prepare", "hyb_count", "hyb_partition", "hyb_local", "fallback"]
for dbg in [int(a) for a in sys.argv[1:]] or [0, 8, 16, 32]:
    _native.set_tuning(dbg=dbg)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    for ev in evs: ev.record()
    torch.cuda.synchronize()
    arr = (ctypes.c_void_p * 11)(*[ctypes.c_void_p(ev.cuda_event) for ev in evs])
    mod = b200.LovaszSoftmax({"experiment": 3})
    with torch.no_grad():
        for _ in range(3): mod(x, y)
        _native.check(lib.b200seg_set_stage_events(arr, 11), "ev")
        acc = [0.0] * 8
        for _ in range(10):
            mod(x, y); torch.cuda.synchronize()
            for i in range(8): acc[i] += evs[i].elapsed_time(evs[i + 1])
        _native.check(lib.b200seg_set_stage_events(None, 0), "ev")
    print("dbg", dbg, {names[i]: round(acc[i] * 100, 1) for i in range(8)}, flush=True)
_native.set_tuning(dbg=0)
