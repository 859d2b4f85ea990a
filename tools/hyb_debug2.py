import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native
from test_gpu_parity import _blocky
lib = _native.load()
n, c, h, w, exp = 2, 25, 540, 960, 3
hw, P = h * w, n * h * w
xa, y = _blocky(n, c, h, w, seed=21, with_ignore=True)
xb, _ = _blocky(n, c, h, w, seed=22, with_ignore=True)
_native.set_tuning(dbg=int(os.environ.get("DBG", "64")))
offs = (ctypes.c_size_t * 14)()
_native.check(lib.b200seg_debug_layout(n, c, hw, 0, offs, 14), "layout")
yd = y.cuda()
for name, x in (("b", xb), ("a", xa), ("b", xb), ("a", xa)):
    xd = x.cuda().requires_grad_(True)
    loss = b200.LovaszSoftmax({"experiment": exp})(xd, yd)
    torch.cuda.synchronize()
    ws = loss.grad_fn.saved_tensors[2]
    v = lambda o, nbytes, dt: ws[o:o + nbytes].view(dt)
    cnt = v(offs[10], 4 * c, torch.int32).cpu().tolist()
    bits = v(offs[11], 4 * c, torch.int32).cpu().tolist()
    keys = v(offs[8], 4 * c * P, torch.int32)
    print(name, "loss", float(loss.detach()), "counts", cnt[:14], flush=True)
    for s in range(c):
        ns = cnt[s]
        if ns == 0: continue
        lg = max(0, (ns - 1).bit_length()); wd = min(max(lg - 6, 0), 13, bits[s]); L = bits[s] - wd
        d = (keys[s * P: s * P + ns].long() & 0xFFFFFFFF) >> L
        bad = int((d[1:] < d[:-1]).sum())
        if bad:
            i = int((d[1:] < d[:-1]).nonzero()[0])
            print(f"   seg {s}: n={ns} out-of-order={bad} first at {i}: {d[max(0, i - 3): i + 4].tolist()}", flush=True)
