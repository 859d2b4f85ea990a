#!/usr/bin/env python
"""Inspect the per-pixel candidate records of one forward call (GPU): guard statistics vs a torch top-k."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miccai2021_cataract_semantic_segmentation_b200 import _native
lib = _native.load()
n, c, h, w = 2, 25, 540, 960
hw = h * w
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, c, h, w), generator=g, device="cuda")
y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda")
nb = _native._sz(0)
_native.check(lib.b200seg_lovasz_workspace_bytes(n, c, hw, 0, nb), "ws")
ws = torch.zeros(nb.value, dtype=torch.uint8, device="cuda")
loss = torch.empty((), device="cuda")
cm = torch.zeros((c, c), dtype=torch.int64, device="cuda"); status = torch.zeros(1, dtype=torch.int32, device="cuda")
_native.check(lib.b200seg_lovasz_forward(x.data_ptr(), y.data_ptr(), _native.LABEL_I64, n, c, hw, 0, _native.NO_LABEL, 0, (1 << c) - 1, 1,
                                         ws.data_ptr(), ws.numel(), loss.data_ptr(), cm.data_ptr(), c, status.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "fwd")
torch.cuda.synchronize()
offs = (ctypes.c_size_t * 8)()
_native.check(lib.b200seg_debug_layout(n, c, hw, 0, offs, 8), "layout")
P = n * hw
def view(off, nbytes, dtype): return ws[off:off + nbytes].view(dtype)
rec16 = view(offs[4], 16 * P, torch.int32).view(P, 4)
rec4 = view(offs[5], 4 * P, torch.int32)
thr = view(offs[6], 4 * c, torch.float32)
tmin = view(offs[7], 4, torch.float32)
p1 = rec16[:, 1].view(torch.float32); p2 = rec16[:, 2].view(torch.float32); p3 = rec16[:, 3].view(torch.float32)
print("loss", float(loss), "tmin", float(tmin), "thr min", float(thr.min()), "thr", thr.sort().values[:4].tolist())
print("frac p1>=tmin", float((p1 >= tmin).float().mean()), "p2", float((p2 >= tmin).float().mean()), "guard", float((p3 >= tmin).float().mean()))
prob = torch.softmax(x, 1).permute(0, 2, 3, 1).reshape(P, c).clone()
yy = y.view(-1)
idx = (yy < c).nonzero().squeeze()
prob[idx, yy[idx]] = -1
top = prob.topk(3, 1)
print("max |p1-top1|", float((p1 - top.values[:, 0]).abs().max()), "|p2-top2|", float((p2 - top.values[:, 1]).abs().max()),
      "guard-top3 min", float((p3 - top.values[:, 2]).min()), "max", float((p3 - top.values[:, 2]).max()))
print("torch frac top3>=tmin", float((top.values[:, 2] >= tmin).float().mean()))
c1 = (rec4 >> 8) & 31; c2 = (rec4 >> 16) & 31
print("c1 match", float((c1 == top.indices[:, 0]).float().mean()), "c2 match", float((c2 == top.indices[:, 1]).float().mean()))
