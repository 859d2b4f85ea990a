"""B200-native (sm_100a) Lovasz-Softmax loss and confusion-matrix mIoU metrics: drop-ins for the hot path of
RViMLab/MICCAI2021_Cataract_semantic_segmentation (losses/LovaszSoftmax.py, utils/torch_utils.py:221-346,
utils/metrics.py, losses/iou.py).  Host side is PyTorch; compute goes through the C ABI in include/b200seg.h."""
from .class_info import CATEGORIES, CLASS_INFO, NUM_CLASSES
from .fused import (AsyncToNumpy, BestModelTracker, GraphedValidationStep, IoUTracker, LossWrapper, LovaszSoftmaxCE, LovaszSoftmaxWithMetrics, SegmentationMeter,
                    TwoScaleLoss)
from .install import install
from .lovasz import LovaszSoftmax, lovasz_softmax, lovasz_softmax_ce
from .upsampled import LovaszSoftmaxUpsampled, lovasz_softmax_upsampled
from .ohem import OhemCrossEntropy, ohem_cross_entropy
from .metrics import (IoU, accumulate_confusion_matrix, accumulate_confusion_matrix_upsampled, get_confusion_matrix, get_mean_iou, get_pixel_accuracy,
                      get_single_class_iou, metrics_summary, normalise_confusion_matrix, set_confusion_dtype,
                      sliding_miou, t_get_confusion_matrix, t_get_mean_iou, t_get_miou, t_get_pixel_accuracy,
                      t_get_single_class_iou, t_normalise_confusion_matrix)

__all__ = [
    "CATEGORIES", "CLASS_INFO", "NUM_CLASSES", "LovaszSoftmax", "LovaszSoftmaxWithMetrics", "SegmentationMeter",
    "lovasz_softmax", "lovasz_softmax_ce", "LovaszSoftmaxUpsampled", "lovasz_softmax_upsampled", "OhemCrossEntropy", "ohem_cross_entropy", "LovaszSoftmaxCE", "LossWrapper", "TwoScaleLoss", "IoUTracker", "AsyncToNumpy", "BestModelTracker", "GraphedValidationStep", "install", "IoU", "accumulate_confusion_matrix", "accumulate_confusion_matrix_upsampled", "metrics_summary", "set_confusion_dtype",
    "sliding_miou", "t_get_confusion_matrix", "t_get_mean_iou", "t_get_miou", "t_get_pixel_accuracy", "t_get_single_class_iou",
    "t_normalise_confusion_matrix", "get_confusion_matrix", "get_mean_iou", "get_pixel_accuracy",
    "get_single_class_iou", "normalise_confusion_matrix",
]
