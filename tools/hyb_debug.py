import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from miccai2021_cataract_semantic_segmentation_b200 import _native
from test_gpu_parity import _blocky
lib = _native.load()
n, c, h, w = 2, 25, 540, 960
x, y = _blocky(n, c, h, w, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 22, with_ignore=True)
_, y = _blocky(n, c, h, w, seed=21, with_ignore=True)
x, y = x.cuda(), y.cuda()
hw, P = h * w, n * h * w
_native.set_tuning(dbg=32)      # local kernel stops after finding its buckets
nb = _native._sz(0)
_native.check(lib.b200seg_lovasz_workspace_bytes(n, c, hw, 0, nb), "ws")
ws = torch.full((nb.value,), 0xAB, dtype=torch.uint8, device="cuda")
loss = torch.empty((), device="cuda")
_native.check(lib.b200seg_lovasz_forward(x.data_ptr(), y.data_ptr(), _native.LABEL_I64, n, c, hw, 0, _native.NO_LABEL, 0,
                                         (1 << c) - 1, 1, ws.data_ptr(), ws.numel(), loss.data_ptr(), None, _native.NO_LABEL,
                                         None, torch.cuda.current_stream().cuda_stream), "fwd")
torch.cuda.synchronize()
offs = (ctypes.c_size_t * 14)()
_native.check(lib.b200seg_debug_layout(n, c, hw, 0, offs, 14), "layout")
v = lambda o, nbytes, dt: ws[o:o + nbytes].view(dt)
cnt = v(offs[10], 4 * c, torch.int32).cpu().tolist()
bits = v(offs[11], 4 * c, torch.int32).cpu().tolist()
keys = v(offs[8], 4 * c * P, torch.int32)
print("counts", cnt); print("bits", bits)
for s in range(c):
    ns = cnt[s]
    if ns == 0: continue
    lg = max(0, (ns - 1).bit_length()); wd = min(max(lg - 6, 0), 13, bits[s]); L = bits[s] - wd
    k = keys[s * P: s * P + ns].long() & 0xFFFFFFFF
    d = k >> L
    bad = int((d[1:] < d[:-1]).sum())
    hist = v(offs[12] + 4 * 8192 * s, 4 * 8192, torch.int32)[: 1 << wd].long()
    print(f"seg {s}: n={ns} bits={bits[s]} w={wd} L={L} maxkey={int(k.max()):#x} maxdigit={int(d.max())} nbins={1 << wd} out-of-order={bad} "
          f"cursor_end={int(hist[-1])}")
    if bad:
        i = int((d[1:] < d[:-1]).nonzero()[0])
        print("   first at", i, d[max(0, i - 3): i + 4].tolist())
