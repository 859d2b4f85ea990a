"""Fused loss + metrics front end: the per-step work of the reference's training loops
(managers/OCRNet_Manager.py:86-117: loss -> backward -> t_get_confusion_matrix -> pixel accuracy -> mIoU)
in one pass over the logits for the forward side, with every result left on the device.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native
from .class_info import CLASS_INFO
from .lovasz import _PRESENT, _resolve_classes, lovasz_softmax
from .metrics import (accumulate_confusion_matrix, confusion_drop_label, metrics_summary,
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
        self.cm = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.drop_label = confusion_drop_label(self.num_classes, no_ignore_class)

    def reset(self):
        self.cm.zero_()
        self.status.zero_()

    def update(self, prediction: torch.Tensor, target: torch.Tensor):
        accumulate_confusion_matrix(prediction, target, self.cm, self.status, self.drop_label)

    def all_reduce(self, group=None):
        from .dist import all_reduce_confusion_matrix
        all_reduce_confusion_matrix(self.cm, group=group, status=self.status)
        return self.cm

    def summary(self):
        """(iou[C], [mIoU, PA, PAC, mIoU_instruments, mIoU_anatomies, mIoU_rare]) as device tensors."""
        return metrics_summary(self.cm, self.experiment)

    def check(self):
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

    def forward(self, prediction: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.meter is None:
            self.meter = SegmentationMeter(self.experiment, prediction.shape[1], prediction.device)
        keep_absent, mask = _resolve_classes(self.classes_to_consider, prediction.shape[1])
        return lovasz_softmax(prediction, target, self.per_image, self.classes_to_ignore, keep_absent, mask,
                              confusion=self.meter.cm, confusion_drop_label=self.meter.drop_label,
                              status=self.meter.status)
