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

_LOSS_NAMES = {"LovaszSoftmax": _lovasz.LovaszSoftmax}
_METRIC_NAMES = {name: getattr(_metrics, name) for name in (
    "t_get_confusion_matrix", "t_normalise_confusion_matrix", "t_get_pixel_accuracy", "t_get_mean_iou",
    "t_get_miou", "t_get_single_class_iou", "get_confusion_matrix", "normalise_confusion_matrix",
    "get_pixel_accuracy", "get_mean_iou", "get_single_class_iou")}


def install(packages=("losses", "utils", "managers"), verbose: bool = False):
    """Returns {module_name: [rebound names]}."""
    replaced = {}
    table = dict(_LOSS_NAMES)
    table.update(_METRIC_NAMES)
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not any(mod_name == p or mod_name.startswith(p + ".") for p in packages):
            continue
        # the defining modules keep their originals so the reference stays inspectable next to the drop-in
        if mod_name in ("losses.LovaszSoftmax", "utils.torch_utils", "utils.metrics"):
            continue
        hits = []
        for name, obj in table.items():
            cur = mod.__dict__.get(name)
            if cur is not None and cur is not obj and callable(cur):
                mod.__dict__[name] = obj
                hits.append(name)
        if hits:
            replaced[mod_name] = hits
            if verbose:
                print(f"[b200seg.install] {mod_name}: {', '.join(hits)}")
    return replaced
