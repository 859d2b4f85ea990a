#!/usr/bin/env python
"""Benchmark of the hot path: Lovasz-Softmax forward + backward + confusion-matrix mIoU on CaDIS-shaped logits.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric / configs[2]): flat Lovasz-Softmax fwd+bwd + mIoU, C=25 (task 3), 540x960, batch
8 per GPU, synthetic N(0,1) logits and uniform labels 0..25 (25 = ignore).  One "step" = one pass of the path over one
batch: fused loss+confusion-matrix forward, backward to the logits, (N>1: one NCCL all-reduce of the 25x25 int64
matrix), IoU/accuracy summary kernel.  Prints ONE JSON line (rank 0).

`value`       device-timed (CUDA events, max over ranks), inputs resident in HBM.
`e2e`         same metric through the public API from pinned HOST buffers: H2D of logits+labels and D2H of
              loss+summary inside the timed region (double-buffered on a copy stream).
`roofline`    dominant kernel group, timed live with CUDA events (b200seg_set_stage_events) in a separate loop of
              the same process, against that kernel's COMPULSORY bytes only (4C+L for the stats pass,
              8C for the backward pass); `roofline_step` is the whole path against 2*4*C + L bytes/pixel.
`configs`     the other measured configurations beside the headline (device-timed, N=1 only): trained-like D2 C=25
              flat, per-image C=17 (BASELINE configs[1]), C=8 flat, and a slice of the confusion-matrix sweep
              (BASELINE configs[4]).  `--dist blocky` makes D2 the main workload; `--sweep-frames F` runs the
              confusion-matrix sweep over F frames sharded over the ranks as the main workload.
`cpu_baseline` / `--impl reference`: the reference's CPU PyTorch path (oracle/port.py, the pinned restatement)
              on this box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "Mpixel/s Lovasz fwd+bwd + mIoU @540x960 C=25"
UNIT = "Mpixel/s"
# kernel groups between the stage events of the library (hybrid sort path, the default; B200SEG_SORT_PATH=1: LSD names)
STAGES_HYBRID = ["stats(+fused confmat, records)", "finalize+decide", "emit", "sort_prepare", "hyb_count", "hyb_partition",
                 "hyb_local(rank+jaccard+loss)", "lsd_fallback(idle)", None, "backward"]
STAGES_LSD = ["stats(+fused confmat, records)", "finalize+decide", "emit", "sort_prepare", "sort_pass0", "sort_pass1",
              "sort_pass2", "jaccard+loss", None, "backward"]
SORT_PATH = int(os.environ.get("B200SEG_SORT_PATH", "0"))
NO_ALLREDUCE = os.environ.get("B200SEG_BENCH_NO_ALLREDUCE", "0") == "1"    # diagnosis only: N independent replicas
STAGES = STAGES_LSD if SORT_PATH == 1 else STAGES_HYBRID
N_EV = len(STAGES) + 1
# hybrid: stats, finalize+decide, emit (record path) + emit (streaming path; one of the two exits at once), sort prepare,
# hyb_count, hyb_partition, hyb_local, fallback (exits at once unless a segment overflowed), backward, metrics = 11
# LSD: ... sort prepare, 3 x (count, scatter), fg_count, jaccard(+loss), backward, metrics = 15
KERNELS_PER_STEP = 15 if SORT_PATH == 1 else 11


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--classes", type=int, default=25)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--per-image", action="store_true")
    ap.add_argument("--dist", default="d1", choices=["d1", "blocky"],
                    help="d1: N(0,1) logits, uniform labels (BASELINE); blocky: trained-like D2 (blocky labels, confident logits)")
    ap.add_argument("--sweep-frames", type=int, default=0,
                    help="main workload = confusion-matrix sweep over this many frames (BASELINE configs[4]), sharded over the ranks")
    ap.add_argument("--sweep-batch", type=int, default=64)
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--cpu-sample-images", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [v for v in sm if mx and v > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(args, device, seed, dist=None, classes=None, batch=None):
    """D1: N(0,1) logits, uniform labels incl. the ignore label.  blocky (D2, trained-like): 16x16 blocks of
    one label, several classes absent, logits = 6 * onehot(label with 10 % random flips) + N(0,1) (pixel accuracy ~0.9)."""
    dist = dist or args.dist
    c = classes or args.classes
    n = batch or args.batch
    h, w = args.height, args.width
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((n, c, h, w), generator=g, device=device)
    hi = c + 1 if c in (17, 25) else c
    if dist == "d1":
        return x, torch.randint(0, hi, (n, h, w), generator=g, device=device)
    coarse = torch.randint(0, hi, (n, (h + 15) // 16, (w + 15) // 16), generator=g, device=device)
    coarse[coarse >= c // 2 + 2] = c if hi > c else 0
    y = coarse.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :h, :w].contiguous()
    noisy = y.clone()
    flips = torch.rand((n, h, w), generator=g, device=device) < 0.10
    noisy[flips] = torch.randint(0, c, (int(flips.sum()),), generator=g, device=device)
    for i in range(n):                                   # image by image: the one-hot of a whole batch would double the footprint
        x[i] += 6.0 * torch.nn.functional.one_hot(noisy[i].clamp(max=c - 1), c).permute(2, 0, 1).float()
    return x, y


def experiment_of(c):
    return {8: 1, 17: 2, 25: 3}[c]


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's PyTorch path (pinned restatement) on host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_step_fn(args, n_images):
    from oracle import port
    torch.set_num_threads(os.cpu_count() or 1)
    c = args.classes
    x, y = make_inputs(args, torch.device("cpu"), seed=0, batch=n_images)
    exp = experiment_of(c)
    if args.sweep_frames:
        yi = y.int()

        def sweep_step():
            cm = port.confusion_matrix(x, yi)
            port.mean_iou(cm, exp, True, rare=True)
            return float(cm.sum())

        return sweep_step, n_images * args.height * args.width

    def step():
        xr = x.clone().requires_grad_(True)
        loss = port.lovasz_softmax(xr, y, exp, per_image=args.per_image)
        loss.backward()
        cm = port.confusion_matrix(x, y.int())
        port.mean_iou(cm, exp, True, rare=True)
        port.pixel_accuracy(cm)
        return float(loss.detach())

    return step, n_images * args.height * args.width


def run_reference_arm(args, rank):
    if rank != 0:
        return
    step, px = cpu_step_fn(args, args.cpu_sample_images)
    for _ in range(min(args.warmup, 1)):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    v = px / dt / 1e6
    sample = (f"{args.cpu_sample_images} of {args.batch} images per step ({args.height}x{args.width}, C={args.classes}), "
              f"{steps} timed steps after 1 warm-up")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args):
    in_mb = args.batch * args.height * args.width * (4 * args.classes + 8) / 1e6
    if args.sweep_frames:
        return {"workload": f"confusion_matrix_sweep C={args.classes} {args.height}x{args.width} {args.sweep_frames} frames "
                            f"(BASELINE.json configs[4]), batches of {args.sweep_batch} frames, one final all-reduce",
                "labels": "int32, uniform 0..C (C = ignore)", "logits": "fp32 N(0,1)",
                "l2": "one batch of logits (3.3 GB at 64 frames) exceeds the 126 MB L2",
                "parallelism": "one process per GPU, frames sharded, one NCCL all-reduce of the CxC int64 matrix per sweep"}
    return {"workload": f"lovasz_softmax_{'per_image' if args.per_image else 'flat'}_fwd_bwd+confmat_miou C={args.classes} "
                        f"{args.height}x{args.width} batch {args.batch}/GPU (BASELINE.json configs[2])"
                        + (" on trained-like D2 inputs" if args.dist == "blocky" else ""),
            "labels": "int64, uniform 0..C (C = ignore)" if args.dist == "d1" else "int64, 16x16 blocks, classes absent",
            "logits": "fp32 N(0,1)" if args.dist == "d1" else "fp32 6*onehot(label, 10% flips) + N(0,1)",
            "l2": f"inputs {in_mb:.0f} MB/GPU exceed the 126 MB L2 (no flush needed between iterations)",
            "parallelism": "one process per GPU, images sharded, one NCCL all-reduce of the CxC int64 matrix per step"}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def _device_time(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def measure_config(b200, args, device, name, dist, classes, per_image, batch, steps, hbm):
    """One secondary configuration, device-timed like the headline: fused loss + confusion matrix forward, backward, summary."""
    exp = experiment_of(classes)
    x, y = make_inputs(args, device, seed=1, dist=dist, classes=classes, batch=batch)
    meter = b200.SegmentationMeter(exp, classes, device)
    mod = b200.LovaszSoftmaxWithMetrics({"experiment": exp, "per_image": per_image}, meter)
    xr = x.requires_grad_(True)

    def step():
        meter.reset()
        xr.grad = None
        loss = mod(xr, y)
        loss.backward()
        meter.summary()
        return loss

    ms = _device_time(step, steps, 3, torch.cuda.synchronize)
    px = batch * args.height * args.width
    bpp = 2 * 4 * classes + 8
    gbs = px * bpp / (ms * 1e-3) / 1e9
    out = {"name": name, "dist": dist, "classes": classes, "per_image": per_image, "batch": batch, "steps": steps,
           "ms_per_step": ms, "value": px / (ms * 1e-3) / 1e6, "unit": UNIT, "bytes_per_pixel": bpp,
           "roofline_step_frac": gbs / hbm, "loss": float(step().detach())}
    del x, y, xr, meter, mod
    torch.cuda.empty_cache()
    return out


def measure_upsampled(b200, args, device, name, stride, classes, batch, steps):
    """the training-loss step from the model's LOW-resolution logits (models/OCR.py:126: stride 8;
    models/DeepLabv3Plus.py:65: stride 4), fused kernels against F.interpolate(align_corners=True) + the full-resolution
    kernels; both produce loss + confusion matrix forward and the gradient w.r.t. the low-resolution logits."""
    import torch.nn.functional as F
    exp = experiment_of(classes)
    h_out, w_out = (args.height + 31) // 32 * 32, args.width          # the models' input size (544 x 960: a multiple of 32)
    h_lo, w_lo = h_out // stride, w_out // stride
    g = torch.Generator(device=device).manual_seed(5)
    # trained-like: blocky coarse label map, logits confident about a 10 % noisy copy of it
    coarse = torch.randint(0, classes, (batch, h_lo, w_lo), generator=g, device=device)
    coarse = coarse // 3 * 3 % classes if classes > 8 else coarse
    y = F.interpolate(coarse[:, None].float(), size=(h_out, w_out), mode="nearest")[:, 0].long()
    noisy = coarse.clone()
    flips = torch.rand((batch, h_lo, w_lo), generator=g, device=device) < 0.10
    noisy[flips] = torch.randint(0, classes, (int(flips.sum()),), generator=g, device=device)
    low = (6.0 * F.one_hot(noisy, classes).permute(0, 3, 1, 2).float()
           + torch.randn((batch, classes, h_lo, w_lo), generator=g, device=device)).contiguous().requires_grad_(True)
    meter = b200.SegmentationMeter(exp, classes, device)

    def fused():
        meter.reset()
        low.grad = None
        loss = b200.lovasz_softmax_upsampled(low, y, confusion=meter.cm, confusion_drop_label=meter.drop_label, status=meter.status)
        loss.backward()
        return loss

    def unfused():
        meter.reset()
        low.grad = None
        full = F.interpolate(low, size=(h_out, w_out), mode="bilinear", align_corners=True)
        loss = b200.lovasz_softmax(full, y, confusion=meter.cm, confusion_drop_label=meter.drop_label, status=meter.status)
        loss.backward()
        return loss

    ms_f = _device_time(fused, steps, 3, torch.cuda.synchronize)
    lf = float(fused().detach())
    cm_f = meter.cm.clone()
    ms_u = _device_time(unfused, steps, 3, torch.cuda.synchronize)
    lu = float(unfused().detach())
    assert lf == lu and torch.equal(cm_f, meter.cm), "fused upsampling differs from interpolate + loss"
    px = batch * h_out * w_out
    out = {"name": name, "dist": "trained-like", "classes": classes, "batch": batch, "low_res": [h_lo, w_lo],
           "size": [h_out, w_out], "steps": steps, "ms_per_step": ms_f, "value": px / (ms_f * 1e-3) / 1e6, "unit": UNIT,
           "interpolate_then_loss_ms_per_step": ms_u, "speedup_vs_interpolate_then_loss": ms_u / ms_f, "loss": lf}
    del low, y, meter, coarse, noisy
    torch.cuda.empty_cache()
    return out


def sweep_time(b200, args, device, frames_rank, world, barrier):
    """Confusion-matrix sweep (managers/BaseManager.py:640-688 of the reference: t_get_confusion_matrix per frame, summed):
    frames_rank frames on this rank in batches of --sweep-batch, int32 labels as the managers pass them, ONE all-reduce of
    the int64 matrix at the end, then the IoU summary.  Returns (ms for the whole sweep, mIoU)."""
    c, exp = args.classes, experiment_of(args.classes)
    bsz = min(args.sweep_batch, frames_rank)
    x, y = make_inputs(args, device, seed=7 + int(os.environ.get("RANK", "0")), dist="d1", batch=bsz)
    y = y.int()
    meter = b200.SegmentationMeter(exp, c, device)
    nb, rem = divmod(frames_rank, bsz)

    def sweep():
        meter.reset()
        for _ in range(nb):
            meter.update(x, y)
        if rem:
            meter.update(x[:rem], y[:rem])
        if world > 1:
            meter.all_reduce()
        return meter.summary()

    sweep()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, summary = sweep()
    e1.record()
    barrier()
    meter.check()
    assert int(meter.cm.sum()) > 0
    ms = e0.elapsed_time(e1)
    del x, y
    torch.cuda.empty_cache()
    return ms, float(summary[0])


def sweep_time_upsampled(b200, args, device, frames, stride=8):
    """The same sweep when the model hands over its stride-8 logits (models/OCR.py:126-131 not applied): the matrix straight
    from them (SegmentationMeter.update_upsampled) against F.interpolate(align_corners=True) + update.  -> (ms fused, ms unfused)."""
    import torch.nn.functional as F
    c, exp = args.classes, experiment_of(args.classes)
    h_out, w_out = (args.height + 31) // 32 * 32, args.width
    bsz = min(args.sweep_batch, frames)
    g = torch.Generator(device=device).manual_seed(11)
    low = torch.randn((bsz, c, h_out // stride, w_out // stride), generator=g, device=device)
    y = torch.randint(0, c + 1, (bsz, h_out, w_out), generator=g, device=device, dtype=torch.int32)
    nb = frames // bsz
    out = []
    mats = []
    for fused in (True, False):
        meter = b200.SegmentationMeter(exp, c, device)

        def sweep():
            meter.reset()
            for _ in range(nb):
                if fused:
                    meter.update_upsampled(low, y)
                else:
                    meter.update(F.interpolate(low, size=(h_out, w_out), mode="bilinear", align_corners=True), y)
            return meter.summary()

        sweep()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sweep()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
        mats.append(meter.cm.clone())
    assert torch.equal(mats[0], mats[1]), "confusion matrix from low-resolution logits differs from interpolate + update"
    del low, y
    torch.cuda.empty_cache()
    return out[0], out[1], nb * bsz * h_out * w_out


def run_sweep_main(b200, bdist, args, rank, world, local, device):
    """`--sweep-frames F`: the sweep is the main workload (BASELINE configs[4]); a step = one batch of frames."""
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lo, hi = bdist.shard_range(args.sweep_frames, rank, world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, miou = sweep_time(b200, args, device, hi - lo, world, barrier)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        hbm, peak_src = peaks()
        px = args.sweep_frames * args.height * args.width
        bpp = 4 * args.classes + 4
        gbs = (hi - lo) * args.height * args.width * bpp / (ms * 1e-3) / 1e9
        nbat = -(-(hi - lo) // min(args.sweep_batch, hi - lo))
        print(json.dumps({
            "metric": "Mpixel/s confusion-matrix mIoU sweep @540x960 C=25", "value": px / (ms * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": nbat, "warmup": nbat, "ms_per_step": ms / nbat, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "i64", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "gpu_launches": nbat + 1, "sweep_ms": ms, "miou": miou,
            "roofline": {"kernel": "confmat_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                         "frac": gbs / hbm, "traffic": None, "peak_source": peak_src,
                         "note": "per GPU: (4C + 4) bytes/pixel (int32 labels), whole sweep incl. the final all-reduce"}}))
    if world > 1:
        dist.destroy_process_group()


def csrc_digest():
    """Digest of the CUDA sources: stamps profiles/dram_traffic.json so a stale capture is not reported as current."""
    import hashlib
    d = os.path.join(ROOT, "miccai2021_cataract_semantic_segmentation_b200", "csrc")
    hsh = hashlib.sha256()
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                hsh.update(f.read())
    return hsh.hexdigest()[:16]


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import miccai2021_cataract_semantic_segmentation_b200 as b200
    from miccai2021_cataract_semantic_segmentation_b200 import _native, dist as bdist
    import torch.distributed as dist

    lib = _native.load()                 # hard failure without the CUDA library: there is no fallback
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (use --impl reference for the CPU arm)"
    rank, world, local = bdist.init_from_env()
    device = torch.device("cuda", local)
    if args.sweep_frames:
        run_sweep_main(b200, bdist, args, rank, world, local, device)
        return
    c, exp = args.classes, experiment_of(args.classes)
    px_rank = args.batch * args.height * args.width
    x, y = make_inputs(args, device, seed=rank)
    meter = b200.SegmentationMeter(exp, c, device)
    loss_mod = b200.LovaszSoftmaxWithMetrics({"experiment": exp, "per_image": args.per_image}, meter)

    def step(xr, yy):
        meter.reset()
        xr.grad = None
        loss = loss_mod(xr, yy)
        # the matrix is complete after the first kernel of the forward pass: the collective waits for that event only
        pending = meter.all_reduce(async_op=True) if (world > 1 and not NO_ALLREDUCE) else None
        loss.backward()
        if pending is not None:
            pending.wait()                                                   # the 5 KB all-reduce ran under the backward kernel
        iou, summary = meter.summary()
        return loss, summary

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    xr = x.clone().requires_grad_(True)
    for _ in range(max(args.warmup, 3)):
        step(xr, y)
    barrier()
    if rank == 0:
        sampler.start()

    # ---- device-resident timing -----------------------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss, summary = step(xr, y)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    ms_ranks = [float(ms) / args.steps]
    if world > 1:
        allms = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(allms, ms)
        ms_ranks = [float(v) / args.steps for v in allms]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    value = px_rank * world / (ms_step * 1e-3) / 1e6
    meter.check()
    loss_value = float(loss.detach())

    # ---- end to end from pinned host buffers ------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
        bufs = [(torch.empty_like(x).requires_grad_(True), torch.empty_like(y)) for _ in range(2)]
        outs = [torch.empty(7, dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device)
        copied = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream(device)

        def e2e_step(i):
            b = i & 1
            done[b].synchronize()                      # host has read (or may read) the previous result of buffer b
            with torch.cuda.stream(copy_stream), torch.no_grad():
                copy_stream.wait_event(done[b])
                bufs[b][0].copy_(xh, non_blocking=True)
                bufs[b][1].copy_(yh, non_blocking=True)
                copied[b].record(copy_stream)
            main_stream.wait_event(copied[b])
            l, s = step(bufs[b][0], bufs[b][1])
            with torch.no_grad():
                outs[b][:1].copy_(l.detach().reshape(1), non_blocking=True)
                outs[b][1:].copy_(s, non_blocking=True)
            done[b].record(main_stream)

        for ev in done:
            ev.record(main_stream)
        for i in range(max(args.warmup, 3)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms2 = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        ms2_step = float(ms2) / args.steps
        e2e = {"value": px_rank * world / (ms2_step * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(x.numel() * 4 + y.numel() * 8), "d2h_bytes_per_step": 28,
               "ms_per_step": ms2_step, "wall_ms_per_step": wall * 1e3 / args.steps,
               "note": "pinned host logits+labels copied every step on a copy stream, double-buffered; "
                       "loss+summary read back every step"}
        assert abs(float(outs[(args.steps - 1) & 1][0]) - loss_value) <= 1e-6 * abs(loss_value) + 1e-12
        del xh, yh, bufs

    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel-group times, live, via stage events ------------------------------------------------------------
    import ctypes
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(N_EV)]
    for ev in evs:
        ev.record()                                  # creates the underlying cudaEvent_t
    torch.cuda.synchronize()
    arr = (ctypes.c_void_p * N_EV)(*[ctypes.c_void_p(ev.cuda_event) for ev in evs])
    _native.check(lib.b200seg_set_stage_events(arr, N_EV), "set_stage_events")
    acc = [0.0] * len(STAGES)
    prof_steps = min(args.steps, 10)
    for _ in range(prof_steps):
        step(xr, y)
        torch.cuda.synchronize()
        for i in range(len(STAGES)):
            if STAGES[i] is not None:
                acc[i] += evs[i].elapsed_time(evs[i + 1])
    _native.check(lib.b200seg_set_stage_events(None, 0), "clear stage events")
    stage_ms = {STAGES[i]: acc[i] / prof_steps for i in range(len(STAGES)) if STAGES[i] is not None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, peak_src = peaks()
    lab_bytes = 8
    p = px_rank
    # COMPULSORY bytes per launch of the two streaming kernels (every input read once, every output written
    # once): stats = logits + labels in, backward = logits in + gradients out.  The pipeline's own per-pixel state (softmax
    # max / denominator, records, masks) is NOT charged; `achieved_with_own_state` adds it for the curious.
    alg = {"stats(+fused confmat, records)": p * (4 * c + lab_bytes), "backward": p * (2 * 4 * c)}
    own = {"stats(+fused confmat, records)": p * 29, "backward": p * 17}
    dom = max(stage_ms, key=stage_ms.get)
    traffic, traffic_note = None, "no capture under profiles/dram_traffic.json"
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("_csrc") == csrc_digest():
            traffic, traffic_note = tj.get(dom), f"ncu dram__bytes of {tj.get('_source')} (same kernel sources as this run)"
        else:
            traffic_note = f"capture {tj.get('_source')} is of other kernel sources: not reported"
    roof_dom = None
    if dom in alg:
        a = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9
        a2 = (alg[dom] + own[dom]) / (stage_ms[dom] * 1e-3) / 1e9
        roof_dom = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
                    "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "ms": stage_ms[dom],
                    "achieved_with_own_state": a2}
    else:       # a sort / scan stage dominates: charge it the whole path's compulsory bytes (it has none of its own)
        a = p * (2 * 4 * c + lab_bytes) / (stage_ms[dom] * 1e-3) / 1e9
        roof_dom = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
                    "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "ms": stage_ms[dom],
                    "note": "stage has no compulsory HBM bytes of its own; charged the path's 2*4*C+L bytes/pixel"}
    step_gbs = p * (2 * 4 * c + lab_bytes) / (ms_step * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "gpu_launches": KERNELS_PER_STEP * args.steps,
        "roofline": roof_dom,
        "roofline_step": {"bound": "hbm", "bytes_per_pixel": 2 * 4 * c + lab_bytes, "achieved": step_gbs, "peak": hbm,
                          "unit": "GB/s", "frac": step_gbs / hbm, "per_gpu_mpx_s": value / world,
                          "peak_source": peak_src},
        "stage_ms": stage_ms, "loss": loss_value, "miou": float(summary[0]),
        "ms_per_step_ranks": ms_ranks,
    }
    if NO_ALLREDUCE:
        out["diagnosis"] = "B200SEG_BENCH_NO_ALLREDUCE=1: independent replicas, no collective (not a valid bench line)"
    out["sort_path"] = "lsd" if SORT_PATH == 1 else "hybrid"
    if world == 1 and not args.no_configs:
        del xr, x, y
        torch.cuda.empty_cache()
        cfgs = []
        for name, dist_, cc, pi in (("d2_blocky_c25_flat", "blocky", 25, False), ("d1_c17_per_image (BASELINE configs[1])", "d1", 17, True),
                                    ("d1_c8_flat", "d1", 8, False), ("d2_blocky_c17_per_image", "blocky", 17, True)):
            if (dist_, cc, pi) == (args.dist, c, args.per_image):
                continue
            cfgs.append(measure_config(b200, args, device, name, dist_, cc, pi, args.batch, min(args.steps, 20), hbm))
        for name, stride in (("fused_upsample_stride8_c25 (SURVEY F2, OCRNet geometry)", 8),
                             ("fused_upsample_stride4_c25 (SURVEY F2, DeepLabv3+ geometry)", 4)):
            cfgs.append(measure_upsampled(b200, args, device, name, stride, 25, args.batch, min(args.steps, 10)))
        frames = 512
        sms, smiou = sweep_time(b200, args, device, frames, 1, torch.cuda.synchronize)
        spx = frames * args.height * args.width
        cfgs.append({"name": "confmat_sweep_slice (BASELINE configs[4]: 512 of 4096 frames = one GPU's share at N=8)",
                     "frames": frames, "batch": args.sweep_batch, "ms": sms, "value": spx / (sms * 1e-3) / 1e6, "unit": UNIT,
                     "bytes_per_pixel": 4 * c + 4, "roofline_frac": spx * (4 * c + 4) / (sms * 1e-3) / 1e9 / hbm, "miou": smiou})
        ums_f, ums_u, upx = sweep_time_upsampled(b200, args, device, frames)
        cfgs.append({"name": "confmat_sweep_slice_from_stride8_logits (SURVEY F2 on the metric path: 512 frames, 68x120 -> 544x960)",
                     "frames": frames, "batch": args.sweep_batch, "ms": ums_f, "value": upx / (ums_f * 1e-3) / 1e6, "unit": UNIT,
                     "interpolate_then_confmat_ms": ums_u, "speedup_vs_interpolate_then_confmat": ums_u / ums_f})
        out["configs"] = cfgs
    if world == 1 and not args.no_cpu_baseline:
        cstep, cpx = cpu_step_fn(args, args.cpu_sample_images)
        cstep()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            cstep()
        dt = (time.perf_counter() - t0) / reps
        out["cpu_baseline"] = {"value": cpx / dt / 1e6, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{args.cpu_sample_images} of {args.batch} images per step, {reps} timed steps "
                                         f"after 1 warm-up, torch CPU ops on all host threads"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
