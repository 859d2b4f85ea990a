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
              the same process; `roofline_step` is the whole path against 2*4*C + L bytes/pixel.
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
STAGES = ["stats(+fused confmat, records)", "finalize+decide", "emit", "sort_prepare", "sort_pass0", "sort_pass1",
          "sort_pass2", "jaccard+loss", None, "backward"]
N_EV = len(STAGES) + 1
# stats, finalize+decide, emit (record path) + emit (streaming path, exits at once), sort prepare, 3 x (count, scatter),
# fg_count, jaccard(+loss), backward, metrics
KERNELS_PER_STEP = 15


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--classes", type=int, default=25)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--per-image", action="store_true")
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


def make_inputs(args, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((args.batch, args.classes, args.height, args.width), generator=g, device=device)
    hi = args.classes + 1 if args.classes in (17, 25) else args.classes
    y = torch.randint(0, hi, (args.batch, args.height, args.width), generator=g, device=device)
    return x, y


def experiment_of(c):
    return {8: 1, 17: 2, 25: 3}[c]


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's PyTorch path (pinned restatement) on host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_step_fn(args, n_images):
    from oracle import port
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    c = args.classes
    x = torch.randn((n_images, c, args.height, args.width), generator=g)
    hi = c + 1 if c in (17, 25) else c
    y = torch.randint(0, hi, (n_images, args.height, args.width), generator=g)
    exp = experiment_of(c)

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
    return {"workload": f"lovasz_softmax_{'per_image' if args.per_image else 'flat'}_fwd_bwd+confmat_miou C={args.classes} "
                        f"{args.height}x{args.width} batch {args.batch}/GPU (BASELINE.json configs[2])",
            "labels": "int64, uniform 0..C (C = ignore)", "logits": "fp32 N(0,1)",
            "l2": f"inputs {in_mb:.0f} MB/GPU exceed the 126 MB L2 (no flush needed between iterations)",
            "parallelism": "one process per GPU, images sharded, one NCCL all-reduce of the CxC int64 matrix per step"}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
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
    c, exp = args.classes, experiment_of(args.classes)
    px_rank = args.batch * args.height * args.width
    x, y = make_inputs(args, device, seed=rank)
    meter = b200.SegmentationMeter(exp, c, device)
    loss_mod = b200.LovaszSoftmaxWithMetrics({"experiment": exp, "per_image": args.per_image}, meter)

    def step(xr, yy):
        meter.reset()
        xr.grad = None
        loss = loss_mod(xr, yy)
        pending = meter.all_reduce(async_op=True) if world > 1 else None    # the matrix is complete after the forward pass
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
    if world > 1:
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
    # algorithmic bytes per launch of each kernel group (DESIGN.md "Kernels"): compulsory reads + writes of that stage
    alg = {
        # logits + labels in; softmax state (8), compact label (1) and the 20-byte candidate record out
        "stats(+fused confmat, records)": p * (4 * c + lab_bytes + 29),
        # logits in, gradients out, 17 B/px of state (softmax max / denominator, own gradient, label8, candidate mask)
        "backward": p * (2 * 4 * c + 17),
    }
    dom = max(stage_ms, key=stage_ms.get)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    roof_dom = None
    if dom in alg:
        a = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9
        roof_dom = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
                    "traffic": traffic, "peak_source": peak_src, "ms": stage_ms[dom]}
    else:       # a sort / scan stage dominates: charge it the whole path's compulsory bytes (it has none of its own)
        a = p * (2 * 4 * c + lab_bytes) / (stage_ms[dom] * 1e-3) / 1e9
        roof_dom = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
                    "traffic": traffic, "peak_source": peak_src, "ms": stage_ms[dom],
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
    }
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
