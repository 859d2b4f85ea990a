"""Debugging aid: the full-resolution gradient the fused backward kernel forms internally (dbg bit 256 dumps it into the dead
sort buffer) against the gradient of the full-resolution kernel on the ATen-upsampled logits."""
import ctypes, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import miccai2021_cataract_semantic_segmentation_b200 as b200
from miccai2021_cataract_semantic_segmentation_b200 import _native
from test_gpu_upsample import _inputs, CASES

lib = _native.load()
name, (n, c, h, w, H, W), dist, opt = CASES[4]
low, y = _inputs(n, c, h, w, H, W, 99 + n * c + h, dist, True)
yd = y.cuda(); lowd = low.cuda().contiguous()
st = torch.cuda.current_stream().cuda_stream
nb = _native._sz(0)
_native.check(lib.b200seg_lovasz_workspace_bytes(n, c, H * W, 0, nb), "ws")
ws = torch.zeros(nb.value, dtype=torch.uint8, device="cuda")
loss = torch.zeros((), device="cuda"); one = torch.ones((), device="cuda")
dlow = torch.zeros_like(lowd)
NO = _native.NO_LABEL
_native.set_tuning(dbg=256)
_native.check(lib.b200seg_lovasz_up_forward(lowd.data_ptr(), h, w, yd.data_ptr(), _native.label_code(yd), n, c, H, W, 0, NO, 0,
                                            (1 << c) - 1, 1, ws.data_ptr(), ws.numel(), loss.data_ptr(), 0, NO, None, None, NO, None, st), "fwd")
_native.check(lib.b200seg_lovasz_up_backward(lowd.data_ptr(), h, w, yd.data_ptr(), _native.label_code(yd), n, c, H, W, 0, NO, 0,
                                             (1 << c) - 1, ws.data_ptr(), ws.numel(), one.data_ptr(), 0, NO, None, dlow.data_ptr(), st), "bwd")
_native.set_tuning(dbg=0)
offs = (_native._sz * 14)()
_native.check(lib.b200seg_debug_layout(n, c, H * W, 0, offs, 14), "layout")
dz_f = ws[offs[8]:offs[8] + 4 * n * c * H * W].view(torch.float32).view(n, c, H, W).clone()
lu = lowd.clone().requires_grad_(True)
full = F.interpolate(lu, size=(H, W), mode="bilinear", align_corners=True)
full.retain_grad()
b200.lovasz_softmax(full, yd).backward()
dz_u = full.grad
diff = (dz_f - dz_u).abs()
print("loss", float(loss), "max |dz|", float(dz_u.abs().max()), "max diff", float(diff.max()), "n differing", int((dz_f != dz_u).sum()), "of", dz_u.numel())
top = torch.topk(diff.flatten(), 8)
for v, i in zip(top.values.tolist(), top.indices.tolist()):
    nn_, r = divmod(i, c * H * W); cc, r = divmod(r, H * W); yy, xx = divmod(r, W)
    print(f"  n={nn_} c={cc} Y={yy} X={xx} label={int(yd[nn_, yy, xx])} fused={float(dz_f.flatten()[i]):.6e} full-res={float(dz_u.flatten()[i]):.6e}")
# adjoint of the fused kernel's own dz in float64 against its low-resolution output
l64 = lowd.double().requires_grad_(True)
F.interpolate(l64, size=(H, W), mode="bilinear", align_corners=True).backward(dz_f.double())
gmax = float(l64.grad.abs().max())
print("fused adjoint vs f64 adjoint of its own dz:", float((dlow.double() - l64.grad).abs().max()) / gmax)
