"""Lovasz-Softmax (+ cross entropy, + confusion matrix) computed straight from the model's low-resolution logits.

The reference's models end with ``F.interpolate(logits, size=input_resolution, mode='bilinear', align_corners=True)``
(models/OCR.py:126-131: stride 8 -> 68 x 120; models/DeepLabv3Plus.py:65-68: stride 4 -> 136 x 240) and hand the
full-resolution tensor to the loss.  ``lovasz_softmax_upsampled(low, target)`` equals
``lovasz_softmax(F.interpolate(low, target.shape[-2:], mode='bilinear', align_corners=True), target)`` -- same loss, same
confusion matrix bit for bit, same gradient w.r.t. ``low`` to rounding -- without ever forming the upsampled tensor or its
gradient: the kernels interpolate in shared memory (b200seg_lovasz_up_forward / _backward, csrc/lovasz_up.cuh).

Shapes outside the fused kernels (C not in {8, 17, 25}, output width not a multiple of 32, horizontal scale below ~3.2)
take ``F.interpolate`` + the full-resolution path and are counted in ``FALLBACK_COUNTS``.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native
from .class_info import CLASS_INFO
from .lovasz import _PRESENT, _resolve_classes, lovasz_softmax, lovasz_softmax_ce

FALLBACK_COUNTS = {"interpolate_torch": 0}


class _LovaszUpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, low, target, per_image, filter_label, keep_absent, class_mask, ce_enabled, ce_ignore, cm, cm_drop,
                status):
        with torch.cuda.device(low.device):
            lib = _native.load()
            n, c, h, w = low.shape
            H, W = target.shape[-2:]
            need_grad = bool(ctx.needs_input_grad[0])
            nbytes = _native._sz(0)
            _native.check(lib.b200seg_lovasz_workspace_bytes(n, c, H * W, int(per_image), nbytes), "workspace query")
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=low.device)
            loss = torch.empty((), dtype=torch.float32, device=low.device)
            ce = torch.empty((), dtype=torch.float32, device=low.device) if ce_enabled else None
            _native.check(lib.b200seg_lovasz_up_forward(
                low.data_ptr(), h, w, target.data_ptr(), _native.label_code(target), n, c, H, W, int(per_image),
                filter_label, keep_absent, class_mask, int(need_grad), ws.data_ptr(), ws.numel(), loss.data_ptr(),
                int(ce_enabled), ce_ignore, ce.data_ptr() if ce_enabled else None,
                cm.data_ptr() if cm is not None else None, cm_drop,
                status.data_ptr() if status is not None else None, _native.stream_ptr(low.device)),
                "b200seg_lovasz_up_forward")
            if need_grad:
                ctx.save_for_backward(low, target, ws)
                ctx.opts = (int(per_image), filter_label, keep_absent, class_mask, int(ce_enabled), ce_ignore)
            if ce_enabled:
                return loss, ce
            return loss

    @staticmethod
    def backward(ctx, grad_lovasz, grad_ce=None):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            low, target, ws = ctx.saved_tensors
            per_image, filter_label, keep_absent, class_mask, ce_enabled, ce_ignore = ctx.opts
            lib = _native.load()
            n, c, h, w = low.shape
            H, W = target.shape[-2:]
            zero = torch.zeros((), dtype=torch.float32, device=low.device)
            gl = (zero if grad_lovasz is None else grad_lovasz.detach().to(torch.float32)).contiguous()
            gc = (zero if grad_ce is None else grad_ce.detach().to(torch.float32)).contiguous()
            dlow = torch.empty_like(low)
            _native.check(lib.b200seg_lovasz_up_backward(
                low.data_ptr(), h, w, target.data_ptr(), _native.label_code(target), n, c, H, W, per_image, filter_label,
                keep_absent, class_mask, ws.data_ptr(), ws.numel(), gl.data_ptr(), ce_enabled, ce_ignore,
                gc.data_ptr() if ce_enabled else None, dlow.data_ptr(), _native.stream_ptr(low.device)),
                "b200seg_lovasz_up_backward")
            return (dlow,) + (None,) * 10


def upsample_supported(low: torch.Tensor, target: torch.Tensor) -> bool:
    """True when the fused kernels cover this (low-resolution logits, full-resolution target) pair."""
    n, c, h, w = low.shape
    H, W = target.shape[-2:]
    return n * H * W > 0 and bool(_native.load().b200seg_lovasz_up_supported(n, c, h, w, H, W))


def lovasz_softmax_upsampled(low: torch.Tensor, target: torch.Tensor, per_image: bool = False, classes_to_ignore=None,
                             keep_absent: int = 0, class_mask: int | None = None, ce_ignore_index="off",
                             confusion: torch.Tensor | None = None, confusion_drop_label: int | None = None,
                             status: torch.Tensor | None = None):
    """Lovasz-Softmax of ``interpolate(low, target.shape[-2:], 'bilinear', align_corners=True)`` against ``target``.

    ``ce_ignore_index`` other than ``"off"`` (an int or None) adds nn.CrossEntropyLoss(ignore_index) of the same upsampled
    logits and makes the call return ``(lovasz, cross_entropy)``.  ``confusion`` / ``status`` as in ``lovasz_softmax``."""
    if low.dim() != 4 or target.dim() != 3 or target.shape[0] != low.shape[0]:
        raise ValueError("low must be [N, C, h, w] and target [N, H, W]")
    _native.require_cuda(low, target)
    n, c, h, w = low.shape
    with_ce = not (isinstance(ce_ignore_index, str) and ce_ignore_index == "off")
    lowf = (low if low.dtype == torch.float32 else low.float()).contiguous()
    target = _native.as_label_tensor(target)
    fused = upsample_supported(lowf, target)
    ce_ign = _native.NO_LABEL if (not with_ce or ce_ignore_index is None) else int(ce_ignore_index)
    if fused and with_ce and classes_to_ignore is not None and 0 <= int(classes_to_ignore) < c and int(classes_to_ignore) != ce_ign:
        fused = False
    if not fused:
        FALLBACK_COUNTS["interpolate_torch"] += 1
        if FALLBACK_COUNTS["interpolate_torch"] == 1:
            warnings.warn("lovasz_softmax_upsampled: shape outside the fused kernels (C in {8, 17, 25}, output width % 32 "
                          "== 0, scale >= ~3.2); such calls go through F.interpolate (counted in upsampled.FALLBACK_COUNTS)",
                          stacklevel=2)
        full = F.interpolate(low, size=tuple(target.shape[-2:]), mode="bilinear", align_corners=True)
        if with_ce:
            return lovasz_softmax_ce(full, target, ce_ignore_index, per_image, classes_to_ignore, keep_absent, class_mask,
                                     confusion, confusion_drop_label, status)
        return lovasz_softmax(full, target, per_image, classes_to_ignore, keep_absent, class_mask, confusion,
                              confusion_drop_label, status)
    if class_mask is None:
        class_mask = (1 << c) - 1
    filt = _native.NO_LABEL if classes_to_ignore is None else int(classes_to_ignore)
    drop = _native.NO_LABEL if confusion_drop_label is None else int(confusion_drop_label)
    if confusion is not None:
        if confusion.dtype != torch.int64 or tuple(confusion.shape) != (c, c) or not confusion.is_contiguous():
            raise ValueError("confusion must be a contiguous int64 [C, C] tensor")
        if status is None:
            raise ValueError("status (int32 [1]) is required with confusion")
    if with_ce and status is None:
        status = torch.zeros(1, dtype=torch.int32, device=low.device)
    return _LovaszUpFunction.apply(lowf, target, bool(per_image), filt, int(keep_absent), int(class_mask), with_ce, ce_ign,
                                   confusion, drop, status)


class LovaszSoftmaxUpsampled(nn.Module):
    """``LovaszSoftmax`` (losses/LovaszSoftmax.py:8-32: same config keys) fed with the model's logits BEFORE its final
    ``F.interpolate(..., mode='bilinear', align_corners=True)`` (models/OCR.py:126, models/DeepLabv3Plus.py:65):
    ``forward(low_res_logits, target)``; the target's size is the size the reference would have upsampled to."""

    def __init__(self, config):
        super().__init__()
        self.experiment = config['experiment']
        self.num_classes = len(CLASS_INFO[self.experiment][1])
        self.per_image = False if 'per_image' not in config else config['per_image']
        self.classes_to_ignore = None if 'classes_to_ignore' not in config else config['classes_to_ignore']
        self.classes_to_consider = _PRESENT if 'classes_to_consider' not in config else config['classes_to_consider']

    def forward(self, low_res_logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        keep_absent, mask = _resolve_classes(self.classes_to_consider, low_res_logits.shape[1])
        return lovasz_softmax_upsampled(low_res_logits, target, self.per_image, self.classes_to_ignore, keep_absent, mask)
