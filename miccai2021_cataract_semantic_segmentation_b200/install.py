"""Install the B200 drop-ins into an importable copy of the reference code base.

The reference binds names at import time: ``losses/__init__.py`` re-exports the class,
``losses/LossWrapper.py`` / ``TwoScaleLoss.py`` / ``SemiSupervisedLoss.py`` look losses up by name in their module
globals, and every manager does ``from utils import t_get_confusion_matrix, ...`` and ``from losses import *``.
``install()`` rebinds those names in every already-imported module of the ``losses``, ``utils`` and ``managers``
packages; call it after importing the reference packages and before constructing a manager / loss.
"""
from __future__ import annotations

import sys

from . import lovasz as _lovasz
from . import metrics as _metrics
from . import ohem as _ohem

_LOSS_NAMES = {"LovaszSoftmax": _lovasz.LovaszSoftmax, "OhemCrossEntropy": _ohem.OhemCrossEntropy}
_METRIC_NAMES = {name: getattr(_metrics, name) for name in (
    "t_get_confusion_matrix", "t_normalise_confusion_matrix", "t_get_pixel_accuracy", "t_get_mean_iou",
    "t_get_miou", "t_get_single_class_iou", "sliding_miou", "get_confusion_matrix", "normalise_confusion_matrix",
    "get_pixel_accuracy", "get_mean_iou", "get_single_class_iou")}


_FUSED_CACHE = {}


def _fused_loss_wrapper(ref_cls):
    if ref_cls in _FUSED_CACHE:
        return _FUSED_CACHE[ref_cls]
    _FUSED_CACHE[ref_cls] = cls = _make_fused_loss_wrapper(ref_cls)
    return cls


def _make_fused_loss_wrapper(ref_cls):
    """The reference's LossWrapper with the CrossEntropyLoss + LovaszSoftmax pair routed through one fused pass;
    every other loss class (DenseContrastiveLoss, TwoScaleLoss, OhemCrossEntropy ...) still runs the reference code."""
    from .fused import LovaszSoftmaxCE, ce_ignore_index, fused_pair_forward

    class LossWrapper(ref_cls):
        def __init__(self, config):
            super().__init__(config)
            if 'LovaszSoftmax' in self.loss_weightings and 'CrossEntropyLoss' in self.loss_weightings:
                self.pair = LovaszSoftmaxCE(config)
                self.lovasz = self.loss_classes['LovaszSoftmax']
                self.ignore_index = ce_ignore_index(config['experiment'])

        def forward(self, deep_features, prediction, labels, loss_list=None, interm_prediction=None, epoch=None):
            if not hasattr(self, 'pair'):
                return super().forward(deep_features, prediction, labels, loss_list, interm_prediction, epoch)
            pair = ('CrossEntropyLoss', 'LovaszSoftmax')
            others = [k for k in self.loss_weightings if k not in pair and (loss_list is None or k in loss_list)]
            base = super().forward(deep_features, prediction, labels, others, interm_prediction, epoch)
            return fused_pair_forward(self, prediction, labels, loss_list, epoch, base_total=base)

    LossWrapper.__qualname__ = LossWrapper.__name__ = "LossWrapper"
    LossWrapper._b200_fused = True
    return LossWrapper


def _two_stream_two_scale(ref_cls):
    """The reference's TwoScaleLoss with its two heads evaluated on two CUDA streams (fused.two_heads_forward)."""
    if ref_cls in _FUSED_CACHE:
        return _FUSED_CACHE[ref_cls]
    import torch
    from .fused import two_heads_forward

    class TwoScaleLoss(ref_cls):
        def forward(self, logits_interm, logits_final, target):
            if not logits_final.is_cuda:
                return super().forward(logits_interm, logits_final, target)
            if getattr(self, "_b200_side", None) is None:
                self._b200_side = torch.cuda.Stream(logits_final.device)
            return two_heads_forward(self.loss_final, self.loss_interm, logits_interm, logits_final, target,
                                     self.w_final, self.w_interm, self._b200_side)

    TwoScaleLoss.__qualname__ = TwoScaleLoss.__name__ = "TwoScaleLoss"
    TwoScaleLoss._b200_fused = True
    _FUSED_CACHE[ref_cls] = TwoScaleLoss
    return TwoScaleLoss


def install(packages=("losses", "utils", "managers"), verbose: bool = False, fuse_ce: bool = False,
            two_stream_heads: bool = False, async_iou: bool = False):
    """Returns {module_name: [rebound names]}.  ``fuse_ce=True`` additionally replaces ``LossWrapper`` wherever it is
    bound by a subclass that evaluates its CrossEntropyLoss + LovaszSoftmax pair in one fused pass;
    ``two_stream_heads=True`` replaces ``TwoScaleLoss`` by a subclass that runs its two heads on two CUDA streams;
    ``async_iou=True`` rebinds ``to_numpy`` inside the ``managers.*`` modules to ``fused.AsyncToNumpy``: the adaptive
    sampler's per-step read of the per-class IoU vector (managers/OCRNet_Manager.py:114-117) then no longer blocks on the
    GPU (it sees the previous step's vector); every other ``to_numpy`` call keeps the reference's behaviour."""
    replaced = {}
    table = dict(_LOSS_NAMES)
    table.update(_METRIC_NAMES)
    if two_stream_heads:
        ref_ts = getattr(sys.modules.get("losses.TwoScaleLoss"), "TwoScaleLoss", None)
        if ref_ts is not None and not getattr(ref_ts, "_b200_fused", False):
            table["TwoScaleLoss"] = _two_stream_two_scale(ref_ts)
    if fuse_ce:
        ref_lw = getattr(sys.modules.get("losses.LossWrapper"), "LossWrapper", None)
        if ref_lw is not None and not getattr(ref_lw, "_b200_fused", False):
            table["LossWrapper"] = _fused_loss_wrapper(ref_lw)
    if async_iou:
        from .fused import AsyncToNumpy
        for mod_name, mod in list(sys.modules.items()):
            if mod is None or not mod_name.startswith("managers."):
                continue
            cur = mod.__dict__.get("to_numpy")
            if cur is not None and callable(cur) and not isinstance(cur, AsyncToNumpy):
                mod.__dict__["to_numpy"] = AsyncToNumpy(cur)
                replaced.setdefault(mod_name, []).append("to_numpy")
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not any(mod_name == p or mod_name.startswith(p + ".") for p in packages):
            continue
        # the defining modules keep their originals so the reference stays inspectable next to the drop-in
        if mod_name in ("losses.LovaszSoftmax", "losses.OhemCrossEntropy", "utils.torch_utils", "utils.metrics"):
            continue
        hits = []
        for name, obj in table.items():
            if (name, mod_name) in (("LossWrapper", "losses.LossWrapper"), ("TwoScaleLoss", "losses.TwoScaleLoss")):
                continue                                   # the defining module keeps the reference class
            cur = mod.__dict__.get(name)
            if cur is not None and cur is not obj and callable(cur):
                mod.__dict__[name] = obj
                hits.append(name)
        if hits:
            replaced[mod_name] = replaced.get(mod_name, []) + hits
            if verbose:
                print(f"[b200seg.install] {mod_name}: {', '.join(hits)}")
    return replaced
