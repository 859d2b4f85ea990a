"""Torch-CPU restatement ("port") of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Arithmetic is kept in the reference's dtype and op order so
that, on the same machine and torch build, results agree with the reference
bit for bit (checked in tests/test_oracle_golden.py against tests/golden/).

The one deliberate difference: sorting uses ``stable=True`` so ties are broken
by ascending flattened pixel index -- the canonical order BASELINE.json's
north_star prescribes ("compared against the reference run with
torch.sort(stable=True)").
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------
# Class tables (indices only).  utils/defaults.py:16-33 (categories) and
# :112-237 (class keys).  Experiment 1: 8 classes, no ignore; 2: 17 + ignore
# label 17; 3: 25 + ignore label 25.
# --------------------------------------------------------------------------
NUM_CLASSES = {1: 8, 2: 17, 3: 25}
CATEGORIES = {
    1: {"anatomies": [0, 4, 5, 6], "instruments": [7], "others": [1, 2, 3], "rare": [2]},
    2: {"anatomies": [0, 4, 5, 6], "instruments": list(range(7, 17)), "others": [1, 2, 3],
        "rare": [16, 10, 9, 12, 14]},
    3: {"anatomies": [0, 4, 5, 6], "instruments": list(range(7, 25)), "others": [1, 2, 3],
        "rare": [24, 20, 21, 22, 18, 23, 19, 16, 12, 11, 14]},
}


def class_keys(experiment: int):
    """Keys of CLASS_INFO[experiment][1] (utils/defaults.py:123-230): 0..C-1 plus 255 for exp 2/3."""
    keys = list(range(NUM_CLASSES[experiment]))
    if experiment in (2, 3):
        keys.append(255)
    return keys


# --------------------------------------------------------------------------
# Lovász-Softmax
# --------------------------------------------------------------------------
def jaccard_gradient(fg_sorted: torch.Tensor) -> torch.Tensor:
    """losses/LovaszSoftmax.py:83-95 (lovasz_grad): fp32 cumsums, J = 1 - I/U, first difference."""
    total_fg = fg_sorted.sum()
    inter = total_fg - fg_sorted.float().cumsum(0)
    union = total_fg + (1 - fg_sorted).float().cumsum(0)
    jac = 1.0 - inter / union
    n = fg_sorted.numel()
    if n > 1:
        jac[1:n] = jac[1:n] - jac[0:-1]
    return jac


def _seq_mean(values):
    """losses/LovaszSoftmax.py:102-120 (mean): sequential add in list order, divide by count; 0 if empty."""
    if len(values) == 0:
        return 0
    acc = values[0]
    for v in values[1:]:
        acc = acc + v
    if len(values) == 1:
        return acc
    return acc / len(values)


def _flatten(prob_nchw: torch.Tensor, lbl_nhw: torch.Tensor, filter_label):
    """losses/LovaszSoftmax.py:63-80 (flatten_probabilities): NCHW -> [P, C]; optional label filter."""
    c = prob_nchw.shape[1]
    flat = prob_nchw.permute(0, 2, 3, 1).contiguous().view(-1, c)
    lbl = lbl_nhw.reshape(-1)
    if filter_label is None:
        return flat, lbl
    keep = lbl != filter_label
    idx = keep.nonzero().squeeze()          # NB: reference crashes downstream when exactly one pixel is kept
    return flat[idx], lbl[keep]


def _flat_loss(prob_pc: torch.Tensor, lbl_p: torch.Tensor, experiment: int, consider, present_only: bool):
    """losses/LovaszSoftmax.py:34-61 (lovasz_softmax_flat)."""
    if prob_pc.numel() == 0:
        return prob_pc * 0.0
    c_total = prob_pc.shape[1]
    classes = list(range(c_total)) if consider in ("all", "present") else list(consider)
    if experiment in (2, 3) and c_total in classes:     # :48-49 (no-op for 'all'/'present')
        classes.remove(c_total)
    terms = []
    for c in classes:
        fg = (lbl_p == c).float()
        if present_only and fg.sum() == 0:              # :53
            continue
        err = (fg - prob_pc[:, c]).abs()                # :56
        err_sorted, order = torch.sort(err, dim=0, descending=True, stable=True)   # :57, canonical ties
        terms.append(torch.dot(err_sorted, jaccard_gradient(fg[order.detach()])))  # :58-60
    return _seq_mean(terms)                             # :61


def lovasz_softmax(logits: torch.Tensor, target: torch.Tensor, experiment: int, per_image: bool = False,
                   classes_to_ignore=None, classes_to_consider="present", present_only=None):
    """losses/LovaszSoftmax.py:19-32 (forward).  Returns what the reference returns
    (0-dim tensor, or python int 0 / an empty tensor in its degenerate cases).

    ``present_only`` reproduces the ``is 'present'`` identity test at :53: the
    default literal is interned (True); a 'present' string read from JSON is
    not (False).  None = decide like CPython would for an interned literal.
    """
    if present_only is None:
        present_only = isinstance(classes_to_consider, str) and classes_to_consider == "present"
    prob = torch.softmax(logits, dim=1)                 # :26
    if per_image:                                       # :27-29
        per = [_flat_loss(*_flatten(p.unsqueeze(0), t.unsqueeze(0), classes_to_ignore),
                          experiment, classes_to_consider, present_only)
               for p, t in zip(prob, target)]
        return _seq_mean(per)
    return _flat_loss(*_flatten(prob, target, classes_to_ignore), experiment, classes_to_consider, present_only)


def lovasz_softmax_with_grad(logits: torch.Tensor, target: torch.Tensor, experiment: int, **kw):
    """Loss and d(loss)/d(logits) through autograd of the restatement (A5b in SURVEY.md §8a)."""
    x = logits.detach().clone().requires_grad_(True)
    loss = lovasz_softmax(x, target, experiment, **kw)
    if not torch.is_tensor(loss) or loss.numel() != 1 or not loss.requires_grad:
        return (float(loss) if not torch.is_tensor(loss) else float(loss.sum())), torch.zeros_like(logits)
    (g,) = torch.autograd.grad(loss, x)
    return loss.detach(), g


# --------------------------------------------------------------------------
# Confusion matrix + metrics (torch twins)
# --------------------------------------------------------------------------
def confusion_matrix(prediction: torch.Tensor, target: torch.Tensor, existing=None, no_ignore_class: bool = True):
    """utils/torch_utils.py:221-241 (t_get_confusion_matrix): cm[pred, gt], int32.

    Counted directly (bincount) instead of through the reference's one-hot fp32
    GEMM; identical while every entry < 2**24 (the GEMM is exact there).  Label
    range errors mirror torch's one_hot: labels must be in [0, K) with K = C+1
    when the ignore column is dropped, else C.
    """
    c = prediction.shape[1]
    pred = prediction.transpose(1, 0).contiguous().view(c, -1).argmax(0)
    t = target.reshape(-1).to(torch.int64)
    k = c + 1 if (no_ignore_class and c in (17, 25)) else c
    if t.numel() and (int(t.min()) < 0 or int(t.max()) >= k):
        raise RuntimeError("Class values must be smaller than num_classes.")
    counts = torch.bincount(pred * k + t, minlength=c * k).view(c, k)[:, :c]
    cm = counts.to(torch.int32)
    if existing is not None:
        cm = cm + existing
    return cm


def ohem_cross_entropy(score: torch.Tensor, target: torch.Tensor, thresh: float = 0.7, min_kept: int = 100000,
                       ignore_label: int = -100):
    """losses/OhemCrossEntropy.py:22-40 (forward; score and target of equal size).  The reference sorts the label
    probabilities of the non-ignored pixels and reads element min(min_kept, n - 1); kthvalue reads the same element.
    Differentiable through the per-pixel losses exactly like the reference (the threshold is a constant)."""
    logp = torch.log_softmax(score, dim=1)
    t = target.to(torch.int64)
    keep = t != ignore_label
    safe = torch.where(keep, t, torch.zeros_like(t))
    lab_logp = logp.gather(1, safe.unsqueeze(1)).squeeze(1)
    losses = -lab_logp[keep]                                               # :28, :36
    p_lab = torch.softmax(score, dim=1).gather(1, safe.unsqueeze(1)).squeeze(1)[keep].detach()   # :27-32
    k = min(min_kept, p_lab.numel() - 1)
    min_value = torch.kthvalue(p_lab, k + 1).values                        # 0-based index k of the ascending sort
    threshold = max(float(min_value), thresh)                              # :34
    return losses[p_lab < threshold].mean()                                # :37-39


def ohem_with_grad(score: torch.Tensor, target: torch.Tensor, **kw):
    x = score.detach().clone().requires_grad_(True)
    loss = ohem_cross_entropy(x, target, **kw)
    (g,) = torch.autograd.grad(loss, x)
    return loss.detach(), g


def sliding_miou(prediction: torch.Tensor, target: torch.Tensor, kernel_size: int, stride: int,
                 original_size: bool = True):
    """utils/torch_utils.py:189-218 (sliding_miou), counted directly: per window and class, I = #(pred == c and
    label == c), U = #(pred == c or label == c) as integers, IoU = I / U in fp32 with 0/0 -> 1, mean over the classes.
    The reference gets the same integers through one-hot + unfold + int and/or sums.  Labels must be in [0, C)
    (the reference's scatter_ rejects anything else)."""
    assert kernel_size % 2 == 1, "Kernel size needs to be odd"
    n, c, h, w = prediction.shape
    t = target.reshape(n, h, w).to(torch.int64)
    if t.numel() and (int(t.min()) < 0 or int(t.max()) >= c):
        raise RuntimeError("Class values must be smaller than num_classes.")
    p = prediction.argmax(1)
    win_p = p.unfold(1, kernel_size, stride).unfold(2, kernel_size, stride)      # [N, vw, hw, k, k] views
    win_t = t.unfold(1, kernel_size, stride).unfold(2, kernel_size, stride)
    ious = []
    for cls in range(c):
        a, b = win_p == cls, win_t == cls
        inter = (a & b).sum((-1, -2)).to(torch.float)
        union = (a | b).sum((-1, -2)).to(torch.float)
        iou = inter / union
        iou[union == 0] = 1
        ious.append(iou)
    m = torch.mean(torch.stack(ious, 1), dim=1)                                  # :205
    if not original_size:
        return m
    m = torch.repeat_interleave(torch.repeat_interleave(m, stride, dim=-2), stride, dim=-1)
    off = kernel_size // 2
    return torch.nn.functional.pad(m, (off, w - m.shape[-1] - off, off, h - m.shape[-2] - off))


def normalise_confusion_matrix(cm: torch.Tensor, mode: str):
    """utils/torch_utils.py:244-256."""
    if mode not in ("row", "col"):
        raise ValueError("Normalise confusion matrix: mode needs to be either 'row' or 'col'.")
    dim = 1 if mode == "row" else 0
    sums = torch.sum(cm, dim=dim, dtype=torch.float)
    sums[sums == 0] = 1
    return cm.to(torch.float) / sums.unsqueeze(dim)


def pixel_accuracy(cm: torch.Tensor):
    """utils/torch_utils.py:259-271: (PA, PAC); PAC divides by prediction-row sums with 0 -> 1."""
    diag = torch.diag(cm).to(torch.float)
    acc = torch.sum(diag) / torch.sum(cm)
    rows = torch.sum(cm, dim=1, dtype=torch.float)
    rows[rows == 0] = 1
    return acc, torch.mean(diag / rows)


def _iou_vector(cm: torch.Tensor, indices):
    """utils/torch_utils.py:321-327: diag / (colsum + rowsum - diag) in fp32, NaN -> 0."""
    diag = cm.diag()[indices].to(torch.float)
    gt_tot = torch.sum(cm, dim=0, dtype=torch.float)[indices]
    pr_tot = torch.sum(cm, dim=1, dtype=torch.float)[indices]
    iou = diag / (gt_tot + pr_tot - diag)
    iou[iou != iou] = 0
    return iou


def miou(cm: torch.Tensor, experiment: int, indices=None, calculate_mean=True):
    """utils/torch_utils.py:306-332 (t_get_miou)."""
    if indices is None:
        indices = [c for c in class_keys(experiment) if c != 255]
    else:
        assert indices in CATEGORIES[experiment].values()
        indices = [c for c in indices if c != 255]
    iou = _iou_vector(cm, indices)
    return iou.mean() if calculate_mean else iou


def single_class_iou(cm: torch.Tensor, experiment: int, single_class: int):
    """utils/torch_utils.py:335-346 (t_get_single_class_iou)."""
    if single_class == 255:
        single_class = cm.shape[0] - 1
    others = [c for c in class_keys(experiment) if not (c == 255 or c == single_class)]
    tp = cm[single_class, single_class]
    fn = torch.sum(cm[:, single_class]) - tp
    fp = torch.sum(cm[single_class, others])
    denom = tp + fp + fn
    if int(denom) == 0:
        return torch.zeros(1)
    return tp.to(torch.float) / denom.to(torch.float)


def mean_iou(cm: torch.Tensor, experiment: int, categories=False, single_class=None, calculate_mean=None,
             rare=False):
    """utils/torch_utils.py:274-303 (t_get_mean_iou) without its (buggy) single_class membership assert."""
    calculate_mean = True if calculate_mean is None else calculate_mean
    assert experiment in (1, 2, 3)
    if single_class is not None:
        assert not categories
        return single_class_iou(cm, experiment, single_class)
    if categories:
        cats = CATEGORIES[experiment]
        out = (miou(cm, experiment, calculate_mean=calculate_mean),
               miou(cm, experiment, cats["instruments"], calculate_mean),
               miou(cm, experiment, cats["anatomies"], calculate_mean))
        if rare:
            out = out + (miou(cm, experiment, cats["rare"], calculate_mean),)
        return out
    return miou(cm, experiment, calculate_mean=calculate_mean)


# --------------------------------------------------------------------------
# Dead-code twins named by north_star (numpy metrics, losses/iou.py)
# --------------------------------------------------------------------------
def np_confusion_matrix(prediction: np.ndarray, target: np.ndarray, existing=None):
    """utils/metrics.py:5-25: numpy twin, no ignore handling (labels must be < C), int32, cm[pred, gt]."""
    c = prediction.shape[1]
    pred = np.argmax(np.moveaxis(prediction, 1, 0).reshape(c, -1), 0)
    t = target.reshape(-1)
    if t.size and (t.min() < 0 or t.max() >= c):
        raise IndexError("label out of range for the numpy confusion matrix")
    cm = np.bincount(pred.astype(np.int64) * c + t.astype(np.int64), minlength=c * c).reshape(c, c).astype("i")
    assert cm.sum() == t.size
    if existing is not None:
        cm = cm + existing
    return cm


def np_single_class_iou(cm: np.ndarray, experiment: int, single_class: int):
    """utils/metrics.py:87-114: tp / (tp + fp + fn) in float64, 0 when the denominator is 0."""
    if single_class == 255:
        single_class = cm.shape[0] - 1
    others = [c for c in class_keys(experiment) if not (c == 255 or c == single_class)]
    tp = cm[single_class, single_class]
    fn = cm[:, single_class].sum() - tp
    fp = cm[single_class, others].sum()
    denom = tp + fp + fn
    return 0 if denom == 0 else float(tp) / denom


def np_mean_iou(cm: np.ndarray, experiment: int, categories=False):
    """utils/metrics.py:57-84 (get_mean_iou): mean over ALL keys incl. 255 -> last class for exp 2/3."""
    every = np.mean([np_single_class_iou(cm, experiment, c) for c in class_keys(experiment)])
    if not categories:
        return every
    cats = CATEGORIES[experiment]
    return (every,
            np.mean([np_single_class_iou(cm, experiment, c) for c in cats["instruments"]]),
            np.mean([np_single_class_iou(cm, experiment, c) for c in cats["anatomies"]]))


def np_pixel_accuracy(cm: np.ndarray):
    """utils/metrics.py:43-54."""
    diag = np.diag(cm)
    rows = np.sum(cm, axis=1)
    rows[rows == 0] = 1
    return np.sum(diag) / np.sum(cm), np.mean(diag / rows)


def soft_iou(x: torch.Tensor, t: torch.Tensor, epsilon=torch.finfo(torch.float32).eps):
    """losses/iou.py:31-35 (function IoU): sum(x*t) / (sum(x*(1-t) + t) + eps) over the last two dims."""
    inter = x.mul(t).sum(dim=[-2, -1])
    union = (x.mul(1 - t) + t).sum(dim=[-2, -1])
    return inter / (union + epsilon)


def cross_entropy(prediction: torch.Tensor, target: torch.Tensor, experiment: int) -> torch.Tensor:
    """nn.CrossEntropyLoss(ignore_index = 17 | 25 | -100) as LossWrapper builds it (losses/LossWrapper.py:17-24,61-62)."""
    ignore = {2: 17, 3: 25}.get(experiment, -100)
    return torch.nn.functional.cross_entropy(prediction, target.long(), ignore_index=ignore)


def loss_wrapper_pair(prediction: torch.Tensor, target: torch.Tensor, experiment: int, w_ce: float, w_lovasz: float,
                      **lovasz_kw) -> torch.Tensor:
    """total = w_ce * CrossEntropyLoss + w_lovasz * LovaszSoftmax (losses/LossWrapper.py:43-73, in that order)."""
    total = torch.zeros((), dtype=torch.float32, device=prediction.device)
    total = total + cross_entropy(prediction, target, experiment) * w_ce
    total = total + lovasz_softmax(prediction, target, experiment, **lovasz_kw) * w_lovasz
    return total
