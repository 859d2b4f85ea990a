"""ctypes binding of oracle/lovasz_cm_ref.c.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_ref.so")
_lib = None
INT64_MIN = -(2 ** 63)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lovasz_cm_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_lovasz.restype = ctypes.c_int
        _lib.oracle_lovasz.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]
        _lib.oracle_confmat.restype = ctypes.c_int
        _lib.oracle_confmat.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def lovasz(logits: np.ndarray, labels: np.ndarray, per_image=False, classes_to_ignore=None,
           classes_to_consider="present", present_only=None, want_grad=True):
    """Returns (loss float32, grad float32 [N,C,H,W] or None).  Arguments mean what the reference config keys mean."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    n, c = logits.shape[:2]
    hw = int(np.prod(logits.shape[2:]))
    if isinstance(classes_to_consider, str):
        mask = np.ones(c, dtype=np.uint8)
        if present_only is None:
            present_only = classes_to_consider == "present"
    else:
        mask = np.zeros(c, dtype=np.uint8)
        mask[[k for k in classes_to_consider if 0 <= k < c]] = 1
        present_only = False
    loss = np.zeros(1, dtype=np.float32)
    grad = np.zeros_like(logits) if want_grad else None
    rc = lib().oracle_lovasz(logits.ctypes.data, labels.ctypes.data, n, c, hw, int(bool(per_image)),
                             INT64_MIN if classes_to_ignore is None else int(classes_to_ignore),
                             0 if present_only else 1, mask.ctypes.data, loss.ctypes.data,
                             grad.ctypes.data if want_grad else None)
    if rc != 0:
        raise MemoryError("oracle_lovasz failed")
    return loss[0], grad


def confmat(pred: np.ndarray, labels: np.ndarray, no_ignore_class=True, existing=None):
    """int64 [C,C], cm[pred, gt]; raises like the reference on out-of-range labels."""
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    n, c = pred.shape[:2]
    hw = int(np.prod(pred.shape[2:]))
    cm = np.zeros((c, c), dtype=np.int64) if existing is None else np.array(existing, dtype=np.int64)
    drop = c if (no_ignore_class and c in (17, 25)) else INT64_MIN
    oob = ctypes.c_int(0)
    lib().oracle_confmat(pred.ctypes.data, labels.ctypes.data, n, c, hw, drop, cm.ctypes.data, ctypes.byref(oob))
    if oob.value:
        raise RuntimeError("Class values must be smaller than num_classes.")
    return cm
