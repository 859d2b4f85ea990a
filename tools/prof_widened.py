import os, sys, torch
sys.path.insert(0, os.getcwd())
import miccai2021_cataract_semantic_segmentation_b200 as b200
n, c, h, w = 8, 25, 544, 960
g = torch.Generator(device="cuda").manual_seed(0)
y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda")
x = torch.randn((n, c, h, w), generator=g, device="cuda").requires_grad_(True)
mod = b200.OhemCrossEntropy({"experiment": 3})
for _ in range(2):
    x.grad = None
    mod(x, y).backward()
    b200.sliding_miou(x.detach(), y.clamp(max=c - 1), 7, 4, original_size=False)
torch.cuda.synchronize()
