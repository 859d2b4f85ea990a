#!/usr/bin/env python
"""Host-to-device bandwidth per GPU with N ranks copying at once (plain pinned cudaMemcpyAsync of the step's 448 MB of logits +
labels, nothing else running): shows whether the end-to-end figure of bench.py beyond 2 GPUs is limited by the box (host
memory / PCIe root complexes shared by several GPUs) or by the code.     torchrun --nproc-per-node N tools/h2d_bandwidth.py"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miccai2021_cataract_semantic_segmentation_b200 import dist as bdist  # noqa: E402

rank, world, local = bdist.init_from_env()
dev = torch.device("cuda", local)
nbytes = 8 * 540 * 960 * (25 * 4 + 8)
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for _ in range(3):
    buf.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    buf.copy_(host, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
t = torch.tensor([gbs], device=dev, dtype=torch.float64)
allv = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, t)
else:
    allv = [t]
if rank == 0:
    v = [float(x) for x in allv]
    print(json.dumps({"ranks": world, "bytes_per_copy": nbytes, "GBps_per_rank": [round(x, 1) for x in v], "GBps_min": round(min(v), 1),
                      "GBps_sum": round(sum(v), 1), "ms_per_copy_slowest": round(nbytes / min(v) / 1e6, 2)}))
if world > 1:
    dist.destroy_process_group()
