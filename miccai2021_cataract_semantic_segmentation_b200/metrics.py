"""Drop-in confusion-matrix metrics (reference: utils/torch_utils.py:221-346, numpy twins utils/metrics.py).

``t_get_confusion_matrix`` runs the fused argmax + histogram kernel; the C x C post-processing
(``t_get_mean_iou`` & co.) stays in torch with the reference's formulas, dtypes and op order so the floats
are bit-identical to the reference's for the same matrix.  Signatures and return conventions are the
reference's.  The one documented difference: the matrix is int64 (reference: int32; values equal, and the
4096-frame validation sweep of BASELINE.json config 5 would overflow int32).  Pass ``dtype=torch.int32`` or call
``set_confusion_dtype(torch.int32)`` for the reference's dtype.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native
from .class_info import CLASS_INFO, mask_of

_CM_DTYPE = torch.int64


def set_confusion_dtype(dtype: torch.dtype):
    global _CM_DTYPE
    assert dtype in (torch.int32, torch.int64)
    _CM_DTYPE = dtype


def confusion_drop_label(num_classes: int, no_ignore_class: bool = True):
    """utils/torch_utils.py:232-236: for 17 / 25 classes the target one-hot gets an extra (ignore) column that
    is sliced off, i.e. label == C is dropped; otherwise no label is dropped."""
    return num_classes if (no_ignore_class and num_classes in (17, 25)) else None


def accumulate_confusion_matrix(prediction: torch.Tensor, target: torch.Tensor, cm: torch.Tensor,
                                status: torch.Tensor, drop_label=None) -> None:
    """cm (int64 [C, C], cm[pred, gt]) += histogram; status (int32 [1]) |= out-of-range flag.  Asynchronous."""
    _native.require_cuda(prediction, target, cm, status)
    if prediction.dim() != 4:
        raise ValueError("prediction must be [N, C, H, W]")
    n, c, h, w = prediction.shape
    if target.numel() != n * h * w:
        raise ValueError("target must hold N*H*W labels")
    pred = prediction.detach()
    pred = (pred if pred.dtype == torch.float32 else pred.float()).contiguous()
    tgt = _native.as_label_tensor(target.detach())
    assert cm.dtype == torch.int64 and cm.is_contiguous() and tuple(cm.shape) == (c, c)
    assert status.dtype == torch.int32
    drop = _native.NO_LABEL if drop_label is None else int(drop_label)
    with torch.cuda.device(pred.device):                      # launches go to the current device: make it the tensors'
        _native.check(_native.load().b200seg_confmat_accumulate(
            pred.data_ptr(), tgt.data_ptr(), _native.label_code(tgt), n, c, h * w, drop, cm.data_ptr(),
            status.data_ptr(), _native.stream_ptr(pred.device)), "b200seg_confmat_accumulate")


# calls of accumulate_confusion_matrix_upsampled that went through F.interpolate (output width not a multiple of 32)
UPSAMPLED_FALLBACK_COUNTS = {"interpolate_torch": 0}


def accumulate_confusion_matrix_upsampled(low_res: torch.Tensor, target: torch.Tensor, cm: torch.Tensor,
                                          status: torch.Tensor, drop_label=None) -> None:
    """``accumulate_confusion_matrix(F.interpolate(low_res, target.shape[-2:], mode='bilinear', align_corners=True), ...)``
    without forming the upsampled logits (models/OCR.py:126-131 followed by utils/torch_utils.py:221-241): the kernel
    interpolates in shared memory with ATen's arithmetic, so the matrix is the same bit for bit.  Asynchronous."""
    _native.require_cuda(low_res, target, cm, status)
    if low_res.dim() != 4 or target.dim() != 3 or target.shape[0] != low_res.shape[0]:
        raise ValueError("low_res must be [N, C, h, w] and target [N, H, W]")
    n, c, h, w = low_res.shape
    big_h, big_w = target.shape[-2:]
    low = low_res.detach()
    low = (low if low.dtype == torch.float32 else low.float()).contiguous()
    tgt = _native.as_label_tensor(target.detach())
    assert cm.dtype == torch.int64 and cm.is_contiguous() and tuple(cm.shape) == (c, c)
    assert status.dtype == torch.int32
    lib = _native.load()
    if not lib.b200seg_confmat_up_supported(n, c, h, w, big_h, big_w):
        UPSAMPLED_FALLBACK_COUNTS["interpolate_torch"] += 1
        full = torch.nn.functional.interpolate(low, size=(big_h, big_w), mode="bilinear", align_corners=True)
        return accumulate_confusion_matrix(full, tgt, cm, status, drop_label)
    drop = _native.NO_LABEL if drop_label is None else int(drop_label)
    with torch.cuda.device(low.device):
        _native.check(lib.b200seg_confmat_up_accumulate(
            low.data_ptr(), h, w, tgt.data_ptr(), _native.label_code(tgt), n, c, big_h, big_w, drop, cm.data_ptr(),
            status.data_ptr(), _native.stream_ptr(low.device)), "b200seg_confmat_up_accumulate")


def raise_if_label_out_of_range(status: torch.Tensor):
    """Synchronises.  Mirrors the RuntimeError torch's one_hot raises inside the reference."""
    s = int(status.item())
    if s & _native.STATUS_SPIN_TIMEOUT:
        raise RuntimeError("b200seg internal error: chained-scan watchdog fired")
    if s & _native.STATUS_LABEL_OOB:
        raise RuntimeError("Class values must be smaller than num_classes.")


def t_get_confusion_matrix(prediction: torch.Tensor, target: torch.Tensor, existing_matrix: torch.Tensor = None,
                           no_ignore_class: bool = True, dtype: torch.dtype = None, validate: bool = True):
    """Expects prediction logits (or probabilities) NCHW and target classes NHW; returns cm[pred, gt].
    reference: utils/torch_utils.py:221-241.  ``validate=False`` skips the (synchronising) label-range check."""
    c = prediction.shape[1]
    dev = prediction.device
    if existing_matrix is not None:
        cm = existing_matrix.to(device=dev, dtype=torch.int64, copy=True).contiguous()
    else:
        cm = torch.zeros((c, c), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    accumulate_confusion_matrix(prediction, target, cm, status, confusion_drop_label(c, no_ignore_class))
    if validate:
        raise_if_label_out_of_range(status)
    dtype = _CM_DTYPE if dtype is None else dtype
    return cm if dtype == torch.int64 else cm.to(dtype)


def sliding_miou(prediction: torch.Tensor, target: torch.Tensor, kernel_size: int, stride: int,
                 original_size: bool = True, validate: bool = True) -> torch.Tensor:
    """Mean IoU inside every kernel_size x kernel_size window (reference: utils/torch_utils.py:189-218).
    Returns [N, windows_v, windows_h], or, with ``original_size``, that map repeated ``stride`` times along both axes
    and zero-padded to [N, H, W] exactly like :208-214.  One pass over the logits plus an L2-resident window pass
    instead of the reference's two [N, C*k*k, windows] unfolds."""
    assert (kernel_size % 2 == 1), "Kernel size needs to be odd"
    _native.require_cuda(prediction, target)
    if prediction.dim() != 4:
        raise ValueError("prediction must be [N, C, H, W]")
    n, c, h, w = prediction.shape
    if target.numel() != n * h * w:
        raise ValueError("target must hold N*H*W labels")
    if kernel_size > h or kernel_size > w:
        raise RuntimeError("sliding_miou: kernel size (%d) exceeds the input (%d x %d)" % (kernel_size, h, w))
    pred = prediction.detach()
    pred = (pred if pred.dtype == torch.float32 else pred.float()).contiguous()
    tgt = _native.as_label_tensor(target.detach())
    dev = pred.device
    lib = _native.load()
    need = _native._sz()
    _native.check(lib.b200seg_sliding_miou_scratch_bytes(n, h, w, need), "b200seg_sliding_miou_scratch_bytes")
    scratch = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
    ver_num_wins, hor_num_wins = (h - kernel_size) // stride + 1, (w - kernel_size) // stride + 1
    out = torch.empty((n, ver_num_wins, hor_num_wins), dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _native.check(lib.b200seg_sliding_miou(
            pred.data_ptr(), tgt.data_ptr(), _native.label_code(tgt), n, c, h, w, kernel_size, stride,
            scratch.data_ptr(), scratch.numel(), out.data_ptr(), status.data_ptr(), _native.stream_ptr(dev)),
            "b200seg_sliding_miou")
    if validate:
        raise_if_label_out_of_range(status)
    if not original_size:
        return out
    out = torch.repeat_interleave(torch.repeat_interleave(out, stride, dim=-2), stride, dim=-1)
    offset = kernel_size // 2
    return torch.nn.functional.pad(out, (offset, w - out.shape[-1] - offset, offset, h - out.shape[-2] - offset))


def t_normalise_confusion_matrix(matrix: torch.Tensor, mode: str):
    """reference: utils/torch_utils.py:244-256."""
    with torch.no_grad():
        if mode not in ('row', 'col'):
            raise ValueError("Normalise confusion matrix: mode needs to be either 'row' or 'col'.")
        dim = 1 if mode == 'row' else 0
        sums = torch.sum(matrix, dim=dim, dtype=torch.float)
        sums[sums == 0] = 1
        return matrix.to(torch.float) / sums.unsqueeze(dim)


def t_get_pixel_accuracy(confusion_matrix: torch.Tensor):
    """(PA, PAC); reference: utils/torch_utils.py:259-271."""
    with torch.no_grad():
        correct = torch.diag(confusion_matrix).to(torch.float)
        acc = torch.sum(correct) / torch.sum(confusion_matrix)
        pred_sums = torch.sum(confusion_matrix, dim=1, dtype=torch.float)
        pred_sums[pred_sums == 0] = 1
        return acc, torch.mean(correct / pred_sums)


def t_get_miou(confusion_matrix: torch.Tensor, experiment: int, indices=None, calculate_mean: bool = None):
    """reference: utils/torch_utils.py:306-332."""
    calculate_mean = True if calculate_mean is None else calculate_mean
    cats = CLASS_INFO[experiment][2]
    if indices is None:
        indices = [c for c in CLASS_INFO[experiment][1].keys() if not c == 255]
    else:
        assert (indices == cats['anatomies'] or indices == cats['instruments'] or
                indices == cats.get('rare') or indices == cats['others']), \
            'indices must be any of the entries of {}'.format(cats)
        indices = [c for c in indices if not c == 255]
    with torch.no_grad():
        diagonal = confusion_matrix.diag()[indices].to(torch.float)
        gt_totals = torch.sum(confusion_matrix, dim=0, dtype=torch.float)[indices]
        pred_totals = torch.sum(confusion_matrix, dim=1, dtype=torch.float)[indices]
        iou = diagonal / (gt_totals + pred_totals - diagonal)
        iou[iou != iou] = 0
        return iou.mean() if calculate_mean else iou


def t_get_single_class_iou(confusion_matrix: torch.Tensor, experiment: int, single_class: int):
    """reference: utils/torch_utils.py:335-346."""
    with torch.no_grad():
        if single_class == 255:
            single_class = confusion_matrix.shape[0] - 1
        indices = [c for c in CLASS_INFO[experiment][1].keys() if not (c == 255 or c == single_class)]
        tp = confusion_matrix[single_class, single_class]
        fn = torch.sum(confusion_matrix[:, single_class]) - tp
        fp = torch.sum(confusion_matrix[single_class, indices])
        denom = tp + fp + fn
        if int(denom) == 0:
            return torch.zeros(1)
        return tp.to(torch.float) / denom.to(torch.float)


def t_get_mean_iou(confusion_matrix: torch.Tensor, experiment: int, categories=False, single_class=None,
                   calculate_mean=None, rare=False):
    """reference: utils/torch_utils.py:274-303.  (The reference's ``single_class in CLASS_INFO[experiment]``
    assert at :282 tests membership in a list of dicts and always fails; here the class id is checked instead.)"""
    calculate_mean = True if calculate_mean is None else calculate_mean
    assert experiment in [1, 2, 3], 'experiment must be in [1,2,3] instead got [{}]'.format(experiment)
    if single_class is not None:
        assert not categories, 'when single_class is not None, category must be False instead got [{}]'.format(categories)
        assert single_class in CLASS_INFO[experiment][1], \
            'single_class must be {} instead got [{}]'.format(CLASS_INFO[experiment][1].keys(), single_class)
        return t_get_single_class_iou(confusion_matrix, experiment, single_class)
    if categories:
        cats = CLASS_INFO[experiment][2]
        out = (t_get_miou(confusion_matrix, experiment, calculate_mean=calculate_mean),
               t_get_miou(confusion_matrix, experiment, indices=cats['instruments'], calculate_mean=calculate_mean),
               t_get_miou(confusion_matrix, experiment, indices=cats['anatomies'], calculate_mean=calculate_mean))
        if rare:
            out = out + (t_get_miou(confusion_matrix, experiment, indices=cats['rare'], calculate_mean=calculate_mean),)
        return out
    return t_get_miou(confusion_matrix, experiment, calculate_mean=calculate_mean)


# ---- device-side summary (one tiny kernel instead of ~15 torch launches per step) --------------------------------
def metrics_summary(cm: torch.Tensor, experiment: int):
    """-> (iou[C] fp32, summary fp32 [6] = mIoU, PA, PAC, mIoU instruments, anatomies, rare), all on the device,
    no synchronisation.  Same formulas as t_get_miou / t_get_pixel_accuracy; equal to them whenever the
    reference's fp32 sums are exact (all counts < 2**24), within 1 ulp of the class mean otherwise."""
    _native.require_cuda(cm)
    assert cm.dtype == torch.int64 and cm.is_contiguous()
    c = cm.shape[0]
    cats = CLASS_INFO[experiment][2]
    keys = [k for k in CLASS_INFO[experiment][1].keys() if k != 255]
    sets = (_native._u32 * 3)(mask_of(cats['instruments']), mask_of(cats['anatomies']), mask_of(cats['rare']))
    iou = torch.empty(c, dtype=torch.float32, device=cm.device)
    summary = torch.empty(6, dtype=torch.float32, device=cm.device)
    with torch.cuda.device(cm.device):
        _native.check(_native.load().b200seg_metrics_from_confmat(
            cm.data_ptr(), c, mask_of(keys), sets, 3, iou.data_ptr(), summary.data_ptr(),
            _native.stream_ptr(cm.device)), "b200seg_metrics_from_confmat")
    return iou, summary


# ---- numpy-facing twins (reference: utils/metrics.py, dead code there; kept for API completeness) ----------------
def _to_numpy(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def get_confusion_matrix(prediction, target, existing_matrix=None):
    """reference: utils/metrics.py:5-25 -- no ignore handling: every label must be < C (IndexError otherwise),
    int32 result, consistency asserts kept.  The counting itself runs on the GPU kernel."""
    pred = torch.as_tensor(_to_numpy(prediction)).cuda() if not (torch.is_tensor(prediction) and prediction.is_cuda) \
        else prediction
    tgt = torch.as_tensor(_to_numpy(target)).cuda() if not (torch.is_tensor(target) and target.is_cuda) else target
    c = pred.shape[1]
    cm = torch.zeros((c, c), dtype=torch.int64, device=pred.device)
    status = torch.zeros(1, dtype=torch.int32, device=pred.device)
    accumulate_confusion_matrix(pred, tgt, cm, status, None)
    if int(status.item()) & _native.STATUS_LABEL_OOB:
        raise IndexError("label out of range: the numpy confusion matrix has no ignore handling")
    out = cm.cpu().numpy().astype('i')
    assert np.sum(out) == tgt.numel()
    if existing_matrix is not None:
        assert existing_matrix.shape == out.shape
        out += existing_matrix
    return out


def normalise_confusion_matrix(matrix, mode):
    """reference: utils/metrics.py:28-40."""
    if mode not in ('row', 'col'):
        raise ValueError("Normalise confusion matrix: mode needs to be either 'row' or 'col'.")
    axis = 1 if mode == 'row' else 0
    sums = matrix.sum(axis=axis)
    sums[sums == 0] = 1
    return matrix / np.expand_dims(sums, axis)


def get_pixel_accuracy(confusion_matrix):
    """reference: utils/metrics.py:43-54."""
    correct = np.diag(confusion_matrix)
    pred_sums = np.sum(confusion_matrix, axis=1)
    pred_sums[pred_sums == 0] = 1
    return np.sum(correct) / np.sum(confusion_matrix), np.mean(correct / pred_sums)


def get_single_class_iou(confusion_matrix, experiment, single_class):
    """reference: utils/metrics.py:87-114."""
    if single_class == 255:
        single_class = confusion_matrix.shape[0] - 1
    not_ignored = [c for c in CLASS_INFO[experiment][1].keys() if not (c == 255 or c == single_class)]
    tp = confusion_matrix[single_class, single_class]
    fn = confusion_matrix[:, single_class].sum() - tp
    fp = confusion_matrix[single_class, not_ignored].sum()
    denom = tp + fp + fn
    return 0 if denom == 0 else float(tp) / denom


def get_mean_iou(confusion_matrix, experiment, categories=False, single_class=None):
    """reference: utils/metrics.py:57-84."""
    assert experiment in [1, 2, 3], 'experiment must be in [1,2,3] instead got [{}]'.format(experiment)
    if single_class is not None:
        assert not categories
        return get_single_class_iou(confusion_matrix, experiment, single_class)
    every = np.mean([get_single_class_iou(confusion_matrix, experiment, c) for c in CLASS_INFO[experiment][1].keys()])
    if not categories:
        return every
    cats = CLASS_INFO[experiment][2]
    return (every,
            np.mean([get_single_class_iou(confusion_matrix, experiment, c) for c in cats['instruments']]),
            np.mean([get_single_class_iou(confusion_matrix, experiment, c) for c in cats['anatomies']]))


def IoU(input: torch.Tensor, target: torch.Tensor, epsilon=torch.finfo(torch.float32).eps):
    """Soft IoU over the last two dims; reference: losses/iou.py:31-35 (the function; the classes there are broken).
    A plain reduction over an H x W map -- not on the timed path, kept in torch."""
    intersection = input.mul(target).sum(dim=[-2, -1])
    union = (input.mul(1 - target) + target).sum(dim=[-2, -1])
    return intersection / (union + epsilon)
