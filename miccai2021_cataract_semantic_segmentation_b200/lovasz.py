"""Drop-in ``LovaszSoftmax`` (reference: losses/LovaszSoftmax.py:8-32) backed by the sm_100a kernels.

Same constructor (a config dict read with the reference's four keys), same ``forward(prediction, target)``,
differentiable w.r.t. ``prediction``; an ``nn.Module`` without parameters or buffers, so checkpoints are unchanged.
"""
from __future__ import annotations

import sys
import warnings

import torch
import torch.nn as nn

from . import _native
from .class_info import CLASS_INFO

_PRESENT = sys.intern("present")
# calls of lovasz_softmax_ce whose cross entropy was evaluated by torch.nn.functional.cross_entropy because the shape is
# outside the pipelined kernels (C not in {8, 17, 25}, plane % 16 != 0, unaligned tensors): a library fallback, counted
FALLBACK_COUNTS = {"cross_entropy_torch": 0}


def _resolve_classes(classes_to_consider, n_classes: int):
    """-> (keep_absent, class_mask).  LovaszSoftmax.py:46,53: 'present' skips classes without foreground; 'all'
    and explicit lists do not.  The reference tests ``is 'present'`` (identity): the default literal passes, an
    equal string built at run time (e.g. read from JSON) does not and silently behaves like 'all'.  Reproduced."""
    full = (1 << n_classes) - 1
    if isinstance(classes_to_consider, str):
        if classes_to_consider is _PRESENT:
            return 0, full
        if classes_to_consider == "present":
            warnings.warn("classes_to_consider == 'present' but is not the interned literal (e.g. it came from a JSON "
                          "file): the reference's `is 'present'` test fails for it and absent classes are NOT skipped; "
                          "reproducing that. Pass sys.intern('present') or omit the key to skip absent classes.",
                          stacklevel=3)
            return 1, full
        if classes_to_consider == "all":
            return 1, full
        raise ValueError("classes_to_consider must be 'present', 'all' or a list of class indices")
    mask = 0
    for c in classes_to_consider:
        c = int(c)
        if c == n_classes:          # LovaszSoftmax.py:48-49 drops index C for experiments 2/3; out of range otherwise
            continue
        if not 0 <= c < n_classes:
            raise IndexError(f"class index {c} out of range for {n_classes} classes")
        if mask & (1 << c):
            # the reference would add the class's term twice and divide by len(list); a weighted mean is outside this path
            raise ValueError(f"class index {c} is listed twice in classes_to_consider: duplicates are not supported")
        mask |= 1 << c
    return 1, mask


class _LovaszFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, per_image, filter_label, keep_absent, class_mask, cm, cm_drop, status):
        with torch.cuda.device(logits.device):
            lib = _native.load()
            n, c, h, w = logits.shape
            hw = h * w
            need_grad = bool(ctx.needs_input_grad[0])
            nbytes = _native._sz(0)
            _native.check(lib.b200seg_lovasz_workspace_bytes(n, c, hw, int(per_image), nbytes), "workspace query")
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=logits.device)
            loss = torch.empty((), dtype=torch.float32, device=logits.device)
            _native.check(lib.b200seg_lovasz_forward(
                logits.data_ptr(), target.data_ptr(), _native.label_code(target), n, c, hw, int(per_image),
                filter_label, keep_absent, class_mask, int(need_grad), ws.data_ptr(), ws.numel(), loss.data_ptr(),
                cm.data_ptr() if cm is not None else None, cm_drop,
                status.data_ptr() if status is not None else None, _native.stream_ptr(logits.device)),
                "b200seg_lovasz_forward")
            if need_grad:
                ctx.save_for_backward(logits, target, ws)
                ctx.opts = (int(per_image), filter_label, keep_absent, class_mask)
            return loss

    @staticmethod
    def backward(ctx, grad_out):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            logits, target, ws = ctx.saved_tensors
            per_image, filter_label, keep_absent, class_mask = ctx.opts
            lib = _native.load()
            n, c, h, w = logits.shape
            go = grad_out.detach().to(torch.float32).contiguous()
            dlogits = torch.empty_like(logits)
            _native.check(lib.b200seg_lovasz_backward(
                logits.data_ptr(), target.data_ptr(), _native.label_code(target), n, c, h * w, per_image, filter_label,
                keep_absent, class_mask, ws.data_ptr(), ws.numel(), go.data_ptr(), dlogits.data_ptr(),
                _native.stream_ptr(logits.device)), "b200seg_lovasz_backward")
            return dlogits, None, None, None, None, None, None, None, None


class _LovaszCEFunction(torch.autograd.Function):
    """Lovasz-Softmax and cross entropy of the same logits in one forward and one backward pass
    (b200seg_lovasz_ce_forward / _backward); returns the two scalars."""

    @staticmethod
    def forward(ctx, logits, target, per_image, filter_label, keep_absent, class_mask, ce_ignore, cm, cm_drop, status):
        with torch.cuda.device(logits.device):
            lib = _native.load()
            n, c, h, w = logits.shape
            hw = h * w
            need_grad = bool(ctx.needs_input_grad[0])
            nbytes = _native._sz(0)
            _native.check(lib.b200seg_lovasz_workspace_bytes(n, c, hw, int(per_image), nbytes), "workspace query")
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=logits.device)
            # two separate 0-dim tensors (not views of one buffer): callers mutate losses in place (LossWrapper.py:71)
            loss = torch.empty((), dtype=torch.float32, device=logits.device)
            ce = torch.empty((), dtype=torch.float32, device=logits.device)
            _native.check(lib.b200seg_lovasz_ce_forward(
                logits.data_ptr(), target.data_ptr(), _native.label_code(target), n, c, hw, int(per_image),
                filter_label, keep_absent, class_mask, int(need_grad), ws.data_ptr(), ws.numel(), loss.data_ptr(),
                ce_ignore, ce.data_ptr(), cm.data_ptr() if cm is not None else None, cm_drop,
                status.data_ptr(), _native.stream_ptr(logits.device)), "b200seg_lovasz_ce_forward")
            if need_grad:
                ctx.save_for_backward(logits, target, ws)
                ctx.opts = (int(per_image), filter_label, keep_absent, class_mask, ce_ignore)
            return loss, ce

    @staticmethod
    def backward(ctx, grad_lovasz, grad_ce):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            logits, target, ws = ctx.saved_tensors
            per_image, filter_label, keep_absent, class_mask, ce_ignore = ctx.opts
            lib = _native.load()
            n, c, h, w = logits.shape
            zero = torch.zeros((), dtype=torch.float32, device=logits.device)
            gl = (zero if grad_lovasz is None else grad_lovasz.detach().to(torch.float32)).contiguous()
            gc = (zero if grad_ce is None else grad_ce.detach().to(torch.float32)).contiguous()
            dlogits = torch.empty_like(logits)
            _native.check(lib.b200seg_lovasz_ce_backward(
                logits.data_ptr(), target.data_ptr(), _native.label_code(target), n, c, h * w, per_image, filter_label,
                keep_absent, class_mask, ws.data_ptr(), ws.numel(), gl.data_ptr(), ce_ignore, gc.data_ptr(),
                dlogits.data_ptr(), _native.stream_ptr(logits.device)), "b200seg_lovasz_ce_backward")
            return dlogits, None, None, None, None, None, None, None, None, None


def lovasz_softmax_ce(prediction: torch.Tensor, target: torch.Tensor, ce_ignore_index: int | None,
                      per_image: bool = False, classes_to_ignore=None, keep_absent: int = 0,
                      class_mask: int | None = None, confusion: torch.Tensor | None = None,
                      confusion_drop_label: int | None = None, status: torch.Tensor | None = None):
    """(Lovasz-Softmax, cross entropy) of the same logits, the cross entropy with nn.CrossEntropyLoss(ignore_index)
    semantics (mean over the non-ignored pixels).  One pass over the logits forward, one backward, whenever the
    library's pipelined kernels cover the call; otherwise the cross entropy is evaluated by torch on the side (counted in
    ``FALLBACK_COUNTS``).  A target outside [0, C) other than ``ce_ignore_index`` does not raise here as torch would: the
    fused pass ORs ``STATUS_LABEL_OOB`` into ``status`` -- pass a status tensor and look at it, or use the modules
    (``LovaszSoftmaxCE`` / ``LossWrapper``), which keep one and raise at their next call / ``check()``."""
    if prediction.dim() != 4:
        raise ValueError("prediction must be [N, C, H, W]")
    _native.require_cuda(prediction, target)
    n, c, h, w = prediction.shape
    if tuple(target.shape) != (n, h, w):
        raise ValueError(f"target shape {tuple(target.shape)} does not match prediction {tuple(prediction.shape)}")
    logits = prediction if prediction.dtype == torch.float32 else prediction.float()
    logits = logits.contiguous()
    target = _native.as_label_tensor(target)
    if class_mask is None:
        class_mask = (1 << c) - 1
    filt = _native.NO_LABEL if classes_to_ignore is None else int(classes_to_ignore)
    drop = _native.NO_LABEL if confusion_drop_label is None else int(confusion_drop_label)
    ce_ign = _native.NO_LABEL if ce_ignore_index is None else int(ce_ignore_index)
    lib = _native.load()
    fused = n * h * w > 0 and bool(lib.b200seg_lovasz_ce_supported(
        logits.data_ptr(), target.data_ptr(), _native.label_code(target), n, c, h * w, None))
    if fused and classes_to_ignore is not None and 0 <= int(classes_to_ignore) < c and int(classes_to_ignore) != ce_ign:
        fused = False
    if not fused:
        FALLBACK_COUNTS["cross_entropy_torch"] += 1
        if FALLBACK_COUNTS["cross_entropy_torch"] == 1:
            warnings.warn("lovasz_softmax_ce: shape outside the fused kernels (C in {8, 17, 25}, plane % 16 == 0, aligned "
                          "tensors); the cross entropy of such calls is evaluated by torch (counted in "
                          "lovasz.FALLBACK_COUNTS)", stacklevel=2)
        lov = lovasz_softmax(prediction, target, per_image, classes_to_ignore, keep_absent, class_mask, confusion,
                             confusion_drop_label, status)
        ce = torch.nn.functional.cross_entropy(prediction, target.long(),
                                               ignore_index=-100 if ce_ignore_index is None else int(ce_ignore_index))
        return lov, ce
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=logits.device)
    return _LovaszCEFunction.apply(logits, target, bool(per_image), filt, int(keep_absent), int(class_mask), ce_ign,
                                   confusion, drop, status)


def lovasz_softmax(prediction: torch.Tensor, target: torch.Tensor, per_image: bool = False,
                   classes_to_ignore=None, keep_absent: int = 0, class_mask: int | None = None,
                   confusion: torch.Tensor | None = None, confusion_drop_label: int | None = None,
                   status: torch.Tensor | None = None) -> torch.Tensor:
    """Functional form.  ``confusion`` (int64 [C, C], accumulated in place) fuses the confusion matrix of
    (argmax(prediction), target) into the same pass over the logits; ``status`` (int32 [1]) then receives the
    out-of-range-label flag."""
    if prediction.dim() != 4:
        raise ValueError("prediction must be [N, C, H, W]")
    _native.require_cuda(prediction, target)
    n, c, h, w = prediction.shape
    if tuple(target.shape) != (n, h, w):
        raise ValueError(f"target shape {tuple(target.shape)} does not match prediction {tuple(prediction.shape)}")
    logits = prediction if prediction.dtype == torch.float32 else prediction.float()
    logits = logits.contiguous()
    target = _native.as_label_tensor(target)
    if class_mask is None:
        class_mask = (1 << c) - 1
    filt = _native.NO_LABEL if classes_to_ignore is None else int(classes_to_ignore)
    drop = _native.NO_LABEL if confusion_drop_label is None else int(confusion_drop_label)
    if confusion is not None:
        if confusion.dtype != torch.int64 or tuple(confusion.shape) != (c, c) or not confusion.is_contiguous():
            raise ValueError("confusion must be a contiguous int64 [C, C] tensor")
        if status is None:
            raise ValueError("status (int32 [1]) is required with confusion")
    return _LovaszFunction.apply(logits, target, bool(per_image), filt, int(keep_absent), int(class_mask),
                                 confusion, drop, status)


class LovaszSoftmax(nn.Module):
    """Multi-class Lovasz-Softmax loss; reference losses/LovaszSoftmax.py:8-32.

    config keys (LovaszSoftmax.py:12-16): ``experiment`` (required, 1/2/3), ``per_image`` (False),
    ``classes_to_ignore`` (None; a single label value whose pixels are removed), ``classes_to_consider``
    ('present' | 'all' | list of class indices).

    Differences from the reference, all supersets: returns a 0-dim zero tensor where the reference returns the
    python int 0 or an empty tensor (no class kept / every pixel filtered); handles the single-valid-pixel case
    the reference crashes on (LovaszSoftmax.py:78).  Ties in the per-class sort are broken by ascending pixel
    index (torch.sort(stable=True)), the canonical order.
    """

    def __init__(self, config):
        super().__init__()
        self.eps = torch.as_tensor(1e-10)
        self.experiment = config['experiment']
        self.num_classes = len(CLASS_INFO[self.experiment][1])
        self.per_image = False if 'per_image' not in config else config['per_image']
        self.classes_to_ignore = None if 'classes_to_ignore' not in config else config['classes_to_ignore']
        self.classes_to_consider = _PRESENT if 'classes_to_consider' not in config else config['classes_to_consider']

    def forward(self, prediction: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        keep_absent, mask = _resolve_classes(self.classes_to_consider, prediction.shape[1])
        return lovasz_softmax(prediction, target, self.per_image, self.classes_to_ignore, keep_absent, mask)
