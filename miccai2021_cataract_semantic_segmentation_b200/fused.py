"""Fused loss + metrics front end: the per-step work of the reference's training loops
(managers/OCRNet_Manager.py:86-117: loss -> backward -> t_get_confusion_matrix -> pixel accuracy -> mIoU)
in one pass over the logits for the forward side, with every result left on the device.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native
from .class_info import CLASS_INFO
from .lovasz import LovaszSoftmax, _PRESENT, _resolve_classes, lovasz_softmax, lovasz_softmax_ce
from .metrics import (accumulate_confusion_matrix, accumulate_confusion_matrix_upsampled, confusion_drop_label, metrics_summary,
                      raise_if_label_out_of_range)


class SegmentationMeter:
    """Running int64 confusion matrix (cm[pred, gt]) on one device, plus a sticky label-range status word.

    ``update`` is the asynchronous form of ``running = t_get_confusion_matrix(output, lbl, running)``
    (managers/OCRNet_Manager.py:161); ``all_reduce`` sums the matrix over the data-parallel ranks;
    ``summary`` launches one tiny kernel for IoU / accuracy; ``check`` is the only call that synchronises.
    """

    def __init__(self, experiment: int, num_classes: int | None = None, device=None, no_ignore_class: bool = True):
        self.experiment = experiment
        self.num_classes = num_classes if num_classes is not None else len(
            [k for k in CLASS_INFO[experiment][1] if k != 255])
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # one packed int64 buffer: the matrix, then one word whose low half is the int32 status the kernels OR into
        # (little-endian), so reset is one memset and the data-parallel sum one collective
        cc = self.num_classes * self.num_classes
        self._buf = torch.zeros(cc + 1, dtype=torch.int64, device=device)
        self.cm = self._buf[:cc].view(self.num_classes, self.num_classes)
        self.status = self._buf[cc:].view(torch.int32)[:1]
        self._reduced = False
        self.drop_label = confusion_drop_label(self.num_classes, no_ignore_class)
        self.cm_ready = None                                   # event recorded when the fused matrix of the last forward is complete
        self._side = None

    def reset(self):
        self._buf.zero_()
        self._reduced = False
        self.cm_ready = None

    def update(self, prediction: torch.Tensor, target: torch.Tensor):
        accumulate_confusion_matrix(prediction, target, self.cm, self.status, self.drop_label)

    def update_upsampled(self, low_res: torch.Tensor, target: torch.Tensor):
        """``update(F.interpolate(low_res, target.shape[-2:], 'bilinear', align_corners=True), target)`` without the upsampled
        logits: for validation loops whose model hands over its stride-8 / stride-4 logits (models/OCR.py:126-131)."""
        accumulate_confusion_matrix_upsampled(low_res, target, self.cm, self.status, self.drop_label)

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum the matrix over the data-parallel ranks.  ``async_op=True``: returns a handle whose ``wait()`` must be
        called before the matrix is read; issue it right after the forward pass and wait after ``loss.backward()`` so the
        5 KB collective hides under the backward kernel."""
        from .dist import all_reduce_packed
        self._reduced = True
        if async_op and self.cm_ready is not None and self.cm.is_cuda:
            # the matrix is complete after the FIRST kernel of the fused forward (event recorded by the library): let the
            # collective wait for that event only, on a side stream, so it runs under the emission / sort kernels instead
            # of queueing behind whatever the current stream has been given since (the backward kernel fills every SM)
            if self._side is None:
                self._side = torch.cuda.Stream(self.cm.device)
            self._side.wait_event(self.cm_ready)
            with torch.cuda.stream(self._side):
                pending = all_reduce_packed(self._buf, group=group, async_op=True)
            self._buf.record_stream(self._side)
            return pending
        pending = all_reduce_packed(self._buf, group=group, async_op=async_op)
        return pending if async_op else self.cm

    def summary(self):
        """(iou[C], [mIoU, PA, PAC, mIoU_instruments, mIoU_anatomies, mIoU_rare]) as device tensors."""
        return metrics_summary(self.cm, self.experiment)

    def check(self):
        if self._reduced:                                  # the status word is a sum over ranks now: non-zero = some rank flagged
            if int(self._buf[-1].item()) != 0:
                raise RuntimeError("Class values must be smaller than num_classes.")
            return
        raise_if_label_out_of_range(self.status)


class LovaszSoftmaxWithMetrics(nn.Module):
    """``LovaszSoftmax`` whose forward also accumulates the confusion matrix of (argmax(prediction), target) into
    a ``SegmentationMeter`` while the logits stream through the first kernel (fused argmax + histogram):
    saves the separate 4*C + L bytes/pixel pass that ``t_get_confusion_matrix`` would cost.
    Same config keys as ``LovaszSoftmax``."""

    def __init__(self, config, meter: SegmentationMeter | None = None):
        super().__init__()
        self.experiment = config['experiment']
        self.num_classes = len(CLASS_INFO[self.experiment][1])
        self.per_image = config.get('per_image', False)
        self.classes_to_ignore = config.get('classes_to_ignore', None)
        self.classes_to_consider = config.get('classes_to_consider', _PRESENT)
        self.meter = meter
        self._event = None

    def forward(self, prediction: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.meter is None:
            self.meter = SegmentationMeter(self.experiment, prediction.shape[1], prediction.device)
        keep_absent, mask = _resolve_classes(self.classes_to_consider, prediction.shape[1])
        lib = _native.load()
        if prediction.is_cuda:
            if self._event is None:
                self._event = torch.cuda.Event()
                self._event.record(torch.cuda.current_stream(prediction.device))      # creates the underlying cudaEvent_t
            _native.check(lib.b200seg_set_confmat_event(self._event.cuda_event), "b200seg_set_confmat_event")
        try:
            loss = lovasz_softmax(prediction, target, self.per_image, self.classes_to_ignore, keep_absent, mask,
                                  confusion=self.meter.cm, confusion_drop_label=self.meter.drop_label,
                                  status=self.meter.status)
        finally:
            if prediction.is_cuda:
                lib.b200seg_set_confmat_event(None)
        self.meter.cm_ready = self._event
        return loss


def ce_ignore_index(experiment: int) -> int:
    """ignore_index LossWrapper gives nn.CrossEntropyLoss (losses/LossWrapper.py:18-24)."""
    return {2: 17, 3: 25}.get(experiment, -100)


class LovaszSoftmaxCE(nn.Module):
    """``forward(prediction, target) -> (lovasz, cross_entropy)``: the two losses the reference's ``LossWrapper``
    evaluates on the same logits (losses/LossWrapper.py:43-73), from one pass over them.
    Same config keys as ``LovaszSoftmax``; the cross entropy ignores label 17 / 25 for experiments 2 / 3.
    With a ``SegmentationMeter`` the confusion matrix is accumulated in the same pass as well."""

    def __init__(self, config, meter: SegmentationMeter | None = None):
        super().__init__()
        self.experiment = config['experiment']
        self.per_image = config.get('per_image', False)
        self.classes_to_ignore = config.get('classes_to_ignore', None)
        self.classes_to_consider = config.get('classes_to_consider', _PRESENT)
        self.ignore_index = ce_ignore_index(self.experiment)
        self.meter = meter
        self._status = None                                    # sticky label-range flag of the calls without a meter

    def check(self):
        """nn.CrossEntropyLoss raises (or device-asserts) on a target outside [0, C) that is not ignore_index; the fused
        pass sets a flag instead.  It is looked at lazily -- here, and at the start of the next forward -- so the hot
        path never synchronises on the step it has just queued."""
        if self._status is not None and int(self._status.item()) & _native.STATUS_LABEL_OOB:
            self._status.zero_()
            raise IndexError("Target is out of bounds (a label outside [0, C) other than ignore_index reached the fused "
                             "cross entropy)")

    def forward(self, prediction: torch.Tensor, target: torch.Tensor):
        keep_absent, mask = _resolve_classes(self.classes_to_consider, prediction.shape[1])
        if self.meter is not None:
            kw = dict(confusion=self.meter.cm, confusion_drop_label=self.meter.drop_label, status=self.meter.status)
        else:
            if self._status is None or self._status.device != prediction.device:
                self._status = torch.zeros(1, dtype=torch.int32, device=prediction.device)
            else:
                self.check()                                   # the previous call's flag (that work has long been queued)
            kw = dict(status=self._status)
        return lovasz_softmax_ce(prediction, target, self.ignore_index, self.per_image, self.classes_to_ignore,
                                 keep_absent, mask, **kw)


class LossWrapper(nn.Module):
    """Stand-alone ``LossWrapper`` (losses/LossWrapper.py:8-74) for the loss pair on the hot path:
    ``config['losses']`` may name ``'CrossEntropyLoss'`` and / or ``'LovaszSoftmax'`` with their weights; both are
    evaluated by one fused pass.  Other loss classes belong to the reference: use ``install(fuse_ce=True)``, which
    derives from the reference's own class and only takes over this pair.  Same constructor keys (``losses``,
    ``device``, ``experiment``, optional ``dc_off_at_epoch``), same ``forward`` signature, ``loss_vals`` and
    ``info_string`` attributes."""

    def __init__(self, config: dict):
        super().__init__()
        self.config = config
        self.loss_weightings = config['losses']
        unknown = [k for k in self.loss_weightings if k not in ('CrossEntropyLoss', 'LovaszSoftmax')]
        if unknown:
            raise NotImplementedError(f"losses {unknown} are outside the accelerated path: use the reference's "
                                      "LossWrapper with miccai2021_cataract_semantic_segmentation_b200.install(fuse_ce=True)")
        self.device = config.get('device', 'cuda')
        self.total_loss = None
        self.loss_vals = {k: 0 for k in self.loss_weightings}
        self.info_string = ', '.join(self.loss_weightings)
        self.dc_off = 'dc_off_at_epoch' in config
        self.pair = LovaszSoftmaxCE(config)
        self.lovasz = LovaszSoftmax(config)
        self.ignore_index = ce_ignore_index(config['experiment'])

    def forward(self, deep_features, prediction, labels, loss_list=None, interm_prediction=None, epoch=None):
        return fused_pair_forward(self, prediction, labels, loss_list, epoch)


def fused_pair_forward(wrapper, prediction, labels, loss_list, epoch, base_total=None):
    """total = sum of w_k * loss_k over {'CrossEntropyLoss', 'LovaszSoftmax'} in ``wrapper.loss_weightings`` order
    (losses/LossWrapper.py:46-73), the pair coming from one fused pass when both are wanted."""
    names = [k for k in wrapper.loss_weightings if k in ('CrossEntropyLoss', 'LovaszSoftmax')]
    wanted = [k for k in names if loss_list is None or k in loss_list]
    lov_on = 'LovaszSoftmax' in wanted and not (wrapper.dc_off and epoch is not None and
                                                epoch < wrapper.config['dc_off_at_epoch'])
    ce_on = 'CrossEntropyLoss' in wanted
    zero = lambda: torch.zeros((), dtype=torch.float, device=prediction.device)     # (no host-to-device copy: stays async)
    lov = ce = None
    if lov_on and ce_on:
        lov, ce = wrapper.pair(prediction, labels)
    elif lov_on:
        lov = wrapper.lovasz(prediction, labels)
    elif ce_on:
        ce = torch.nn.functional.cross_entropy(prediction, labels.long(), ignore_index=wrapper.ignore_index)
    total = base_total
    for k in names:
        val = (lov if k == 'LovaszSoftmax' else ce)
        val = zero() if val is None else val
        wgt = wrapper.loss_weightings[k]
        val = val if wgt == 1 else val * wgt
        wrapper.loss_vals[k] = val
        total = val if total is None else total + val
    if total is None:
        total = zero()
    wrapper.total_loss = total
    return total


def two_heads_forward(loss_final, loss_interm, logits_interm, logits_final, target, w_final, w_interm, side_stream):
    """loss_final(logits_final, target) * w_final + loss_interm(logits_interm, target) * w_interm
    (losses/TwoScaleLoss.py:43-52) with the two heads on two CUDA streams: each head is a serial chain of
    HBM-bound passes and latency-bound sort kernels, so the chains of two independent heads fill each other's gaps.
    autograd replays each head's backward on the stream its forward ran on."""
    ph, pw = logits_interm.size(2), logits_interm.size(3)
    h, w = target.size(1), target.size(2)
    if ph != h or pw != w:                                  # F.upsample(..., mode='bilinear') of the reference
        logits_interm = torch.nn.functional.interpolate(logits_interm, size=(h, w), mode='bilinear')
    if not logits_final.is_cuda or side_stream is None:
        return loss_final(logits_final, target) * w_final + loss_interm(logits_interm, target) * w_interm
    cur = torch.cuda.current_stream(logits_final.device)
    side_stream.wait_stream(cur)
    with torch.cuda.stream(side_stream):
        li = loss_interm(logits_interm, target)
    lf = loss_final(logits_final, target)
    cur.wait_stream(side_stream)
    for t in (logits_interm, target):
        t.record_stream(side_stream)
    li.record_stream(cur)
    return lf * w_final + li * w_interm


class TwoScaleLoss(nn.Module):
    """Stand-alone ``TwoScaleLoss`` (losses/TwoScaleLoss.py:8-52) for the Lovasz-Lovasz pair of the published OCRNet
    configuration (configs/OCRNet_rf_lvsz.json:24-28) and the CE-CE pair: same constructor dict (``interm`` / ``final``
    with ``name``, ``args``, optional ``weight``; ``experiment``), same ``forward(logits_interm, logits_final, target)``.
    The two heads run on two CUDA streams (see ``two_heads_forward``)."""

    def __init__(self, config):
        super().__init__()
        names = (config['interm']['name'], config['final']['name'])
        self.w_interm = config['interm'].get('weight', 0.4)
        self.w_final = config['final'].get('weight', 1.0)
        self.ignore_label = -100
        if 'experiment' in config:
            self.ignore_label = len(CLASS_INFO[config['experiment']][1]) - 1 if config['experiment'] in [2, 3] else -100
        config['interm'].update({"experiment": config['experiment']})
        config['final'].update({"experiment": config['experiment']})
        if names == ('CrossEntropyLoss', 'CrossEntropyLoss'):
            self.loss_interm = nn.CrossEntropyLoss(*config['interm'].get('args', []), ignore_index=self.ignore_label)
            self.loss_final = nn.CrossEntropyLoss(*config['final'].get('args', []), ignore_index=self.ignore_label)
        elif names == ('LovaszSoftmax', 'LovaszSoftmax'):
            self.loss_interm = LovaszSoftmax(config['interm'])
            self.loss_final = LovaszSoftmax(config['final'])
        elif names[0] == names[1]:
            raise NotImplementedError(f"{names[0]} is outside the accelerated path: use the reference's TwoScaleLoss "
                                      "after miccai2021_cataract_semantic_segmentation_b200.install()")
        else:
            raise NotImplementedError('different losses for interm {} and final {}'.format(config['interm'], config['final']))
        self._side = None

    def forward(self, logits_interm, logits_final, target):
        if logits_final.is_cuda and self._side is None:
            self._side = torch.cuda.Stream(logits_final.device)
        return two_heads_forward(self.loss_final, self.loss_interm, logits_interm, logits_final, target,
                                 self.w_final, self.w_interm, self._side)


class IoUTracker:
    """Running per-class IoU for the reference's adaptive batch sampler without a per-step host synchronisation
    (SURVEY.md 8 F3).  The reference does, every training step (managers/OCRNet_Manager.py:114-117)::

        iou_values = (1 - a) * self.metrics['iou_values'] + a * to_numpy(iou)     # blocks on the GPU
        self.metrics['iou_values'][:] = iou_values                               # read by AdaptiveBatchSampler.get_prob

    Here the exponential average lives on the device (``update(iou)`` takes the device vector ``meter.summary()`` or
    ``t_get_mean_iou(..., calculate_mean=False)`` returns) and is mirrored into pinned host memory asynchronously, two
    buffers in turn; ``host_values()`` hands the sampler the newest mirror whose copy has completed -- at most one step
    behind, never blocking.  ``host_values(wait=True)`` gives the exact current value (epoch end)."""

    def __init__(self, num_values: int, alpha: float, device=None, init=None):
        device = torch.device("cpu") if device is None else torch.device(device)
        self.alpha = float(alpha)
        self.ema = torch.zeros(num_values, dtype=torch.float32, device=device)
        if init is not None:
            self.ema.copy_(torch.as_tensor(init, dtype=torch.float32))
        pin = device.type == "cuda"
        self._host = [torch.zeros(num_values, dtype=torch.float32, pin_memory=pin) for _ in range(2)]
        for h in self._host:
            h.copy_(self.ema)
        self._events = [None, None]
        self._next = 0                                         # buffer the next update writes
        self._latest = 1                                       # newest buffer known complete

    def update(self, iou: torch.Tensor):
        self.ema.mul_(1.0 - self.alpha).add_(iou.to(self.ema.dtype), alpha=self.alpha)
        b = self._next
        self._host[b].copy_(self.ema, non_blocking=True)
        if self.ema.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.ema.device))
            self._events[b] = ev
        else:
            self._latest = b
        self._next = 1 - b
        return self.ema

    def host_values(self, wait: bool = False):
        """numpy view of the newest completed mirror (``wait=True``: of the current value)."""
        newest = 1 - self._next                                # buffer written by the last update
        ev = self._events[newest]
        if ev is not None:
            if wait:
                ev.synchronize()
            if ev.query():
                self._latest = newest
        elif not self.ema.is_cuda:
            self._latest = newest
        return self._host[self._latest].numpy()


class AsyncToNumpy:
    """Drop-in for the reference's ``utils.to_numpy`` (utils/utils.py:463-468) inside the manager modules, installed by
    ``install(async_iou=True)``.  The one per-step caller on the training path is the adaptive sampler's
    ``to_numpy(iou)`` (managers/OCRNet_Manager.py:114-117), a blocking device-to-host read of <= 26 floats in the middle
    of the step.  For a small 1-D float CUDA tensor this version starts an asynchronous copy into pinned memory and hands
    back the newest copy that has already completed -- the value of the previous step (the very first call waits).  The
    sampler's exponential average is therefore one step behind and the training step no longer synchronises there.  Every
    other argument (images, matrices for figures, CPU tensors) takes the reference's blocking path unchanged."""

    MAX_ELEMENTS = 64

    def __init__(self, blocking):
        self.blocking = blocking
        self.slots = {}                                          # (device, n) -> [buffers, events, next, latest]

    def __call__(self, tensor):
        if not (torch.is_tensor(tensor) and tensor.is_cuda and tensor.dim() == 1 and tensor.is_floating_point()
                and tensor.numel() <= self.MAX_ELEMENTS):
            return self.blocking(tensor)
        key = (tensor.device, tensor.numel(), tensor.dtype)
        st = self.slots.get(key)
        with torch.no_grad():
            if st is None:
                bufs = [torch.empty(tensor.numel(), dtype=tensor.dtype, pin_memory=True) for _ in range(2)]
                bufs[0].copy_(tensor)                            # first call: nothing older to hand out
                torch.cuda.current_stream(tensor.device).synchronize()
                self.slots[key] = st = [bufs, [None, None], 1, 0]
                return bufs[0].numpy().copy()
            bufs, events, nxt, latest = st
            prev = 1 - nxt                                        # written by the previous call
            if events[prev] is not None and events[prev].query():
                latest = prev
            bufs[nxt].copy_(tensor.detach(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(tensor.device))
            events[nxt] = ev
            st[2], st[3] = 1 - nxt, latest
            if latest == nxt:                                     # never hand out the buffer being overwritten
                events[nxt].synchronize()
            return bufs[latest].numpy().copy()


class BestModelTracker:
    """Device-side twin of the best-model bookkeeping at the end of the reference's validation
    (managers/OCRNet_Manager.py:208-223): the mean IoUs are rounded to 4 decimals (``round(float(x), 4)``) and a new best
    is declared when the rounded mIoU exceeds the best so far.  ``update`` takes the device scalars
    (``metrics_summary`` / ``t_get_mean_iou(..., True, rare=True)``) and neither synchronises nor leaves the device;
    ``poll()`` tells the host, without blocking, whether the last update was a new best once its flag has arrived
    (``wait=True`` blocks -- at epoch end, before the checkpoint is written).
    Rounding: a float32 times 1e4 is exact in float64, so round-half-even of that product divided by 1e4 is the same
    double as Python's correctly rounded ``round(float(x), 4)`` (checked bit for bit in tests/test_gpu_ce.py)."""

    def __init__(self, device=None, best_miou: float = 0.0):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.best = torch.zeros(4, dtype=torch.float64, device=device)       # mIoU, anatomies, instruments, rare
        self.best[0] = best_miou
        pin = device.type == "cuda"
        self._host = torch.zeros(5, dtype=torch.float64, pin_memory=pin)     # flag, then the four values of the last update
        self._event = None

    @staticmethod
    def round4(x: torch.Tensor) -> torch.Tensor:
        return torch.round(x.to(torch.float64) * 1e4) / 1e4

    def update(self, m_iou, m_iou_anatomies, m_iou_instruments, m_iou_rare):
        vals = torch.stack([torch.as_tensor(v, device=self.best.device).reshape(()).to(torch.float64)
                            for v in (m_iou, m_iou_anatomies, m_iou_instruments, m_iou_rare)])
        vals = torch.cat([self.round4(vals[:3]), vals[3:]])                  # the reference rounds the first three only
        flag = vals[0] > self.best[0]
        self.best = torch.where(flag, vals, self.best)
        self._host.copy_(torch.cat([flag.to(torch.float64).reshape(1), vals]), non_blocking=True)
        if self.best.is_cuda:
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(self.best.device))
        return flag

    def poll(self, wait: bool = False):
        """(is_new_best, [mIoU, anatomies, instruments, rare]) of the last update, or None while its flag is in flight."""
        if self._event is not None:
            if wait:
                self._event.synchronize()
            elif not self._event.query():
                return None
        h = self._host.numpy()
        return bool(h[0] != 0.0), [float(v) for v in h[1:]]


class GraphedValidationStep:
    """The reference validates frame by frame (managers/OCRNet_Manager.py:146-161: batch 1, ``torch.no_grad()``, loss value
    + running confusion matrix per 544 x 960 frame).  At that size every kernel of the forward pass is a few microseconds
    long and the step is bound by its ~10 launches.  This helper captures the forward-only step (fused loss + confusion
    matrix) for one input shape in a CUDA graph and replays it per frame: one launch, no per-call allocation.

        step = GraphedValidationStep({"experiment": 3}, meter, (1, 25, 544, 960))
        for img, lbl in loader:
            loss = step(model(img), lbl)          # device scalar; accumulate with `total += loss`, read once at the end
    The logits and labels are copied into the graph's static buffers (one pass over the frame, ~20 us at 544 x 960); pass
    ``step.logits`` as the model's output buffer (``out=`` / in-place) to avoid even that."""

    def __init__(self, config, meter: SegmentationMeter, shape, label_dtype=torch.int64, device=None):
        device = meter.cm.device if device is None else torch.device(device)
        self.meter = meter
        self.module = LovaszSoftmaxWithMetrics(config, meter)
        self.logits = torch.zeros(shape, dtype=torch.float32, device=device)
        self.labels = torch.zeros((shape[0], shape[2], shape[3]), dtype=label_dtype, device=device)
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():          # warm-up outside the capture (lazy initialisation of the library)
            for _ in range(2):
                self.module(self.logits, self.labels)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        saved = meter.cm.clone()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.loss = self.module(self.logits, self.labels)
        meter.cm.copy_(saved)                                   # the warm-up frames are not part of anybody's statistics
        meter.reset()

    def __call__(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if logits.data_ptr() != self.logits.data_ptr():
            self.logits.copy_(logits, non_blocking=True)
        if labels.data_ptr() != self.labels.data_ptr():
            self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.loss
