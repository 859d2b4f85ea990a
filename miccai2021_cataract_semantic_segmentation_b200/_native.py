"""ctypes binding of libb200seg.so (the C ABI declared in include/b200seg.h).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but every
compute entry point requires CUDA tensors, and a missing library is a hard error.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import build as _build

LABEL_U8, LABEL_I32, LABEL_I64 = 0, 1, 2
NO_LABEL = -(2 ** 63)
STATUS_LABEL_OOB = 1
STATUS_SPIN_TIMEOUT = 2

_c = ctypes
_vp, _i32, _i64, _u32, _sz = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_uint32, _c.c_size_t

# name -> (restype, argtypes): one entry per symbol include/b200seg.h declares
SIGNATURES = {
    "b200seg_version": (_c.c_int, []),
    "b200seg_last_error": (_c.c_char_p, []),
    "b200seg_lovasz_workspace_bytes": (_c.c_int, [_i32, _i32, _i64, _i32, _c.POINTER(_sz)]),
    "b200seg_lovasz_forward": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i32, _i64, _i32, _u32, _i32, _vp, _sz,
                                           _vp, _vp, _i64, _vp, _vp]),
    "b200seg_lovasz_backward": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i32, _i64, _i32, _u32, _vp, _sz,
                                            _vp, _vp, _vp]),
    "b200seg_lovasz_ce_supported": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _vp]),
    "b200seg_lovasz_ce_forward": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i32, _i64, _i32, _u32, _i32, _vp, _sz,
                                              _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "b200seg_lovasz_ce_backward": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i32, _i64, _i32, _u32, _vp, _sz,
                                               _vp, _i64, _vp, _vp, _vp]),
    "b200seg_lovasz_up_supported": (_c.c_int, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "b200seg_lovasz_up_forward": (_c.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i32, _u32, _i32,
                                              _vp, _sz, _vp, _i32, _i64, _vp, _vp, _i64, _vp, _vp]),
    "b200seg_lovasz_up_backward": (_c.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i32, _u32,
                                               _vp, _sz, _vp, _i32, _i64, _vp, _vp, _vp]),
    "b200seg_confmat_up_supported": (_c.c_int, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "b200seg_confmat_up_accumulate": (_c.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp]),
    "b200seg_confmat_accumulate": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp]),
    "b200seg_metrics_from_confmat": (_c.c_int, [_vp, _i32, _u32, _c.POINTER(_u32), _i32, _vp, _vp, _vp]),
    "b200seg_sliding_miou_scratch_bytes": (_c.c_int, [_i32, _i64, _i64, _c.POINTER(_sz)]),
    "b200seg_sliding_miou": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "b200seg_ohem_workspace_bytes": (_c.c_int, [_i32, _i64, _c.POINTER(_sz)]),
    "b200seg_ohem_ce_forward": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _c.c_float, _i64, _vp, _sz, _vp, _vp, _vp]),
    "b200seg_ohem_ce_backward": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _vp, _sz, _vp, _vp, _vp]),
    "b200seg_set_stage_events": (_c.c_int, [_c.POINTER(_vp), _i32]),
    "b200seg_set_confmat_event": (_c.c_int, [_vp]),
    "b200seg_set_tuning": (_c.c_int, [_c.c_char_p, _i32]),
    "b200seg_debug_exp_mismatches": (_c.c_int, [_vp, _i32, _vp, _vp]),
    "b200seg_debug_upsample": (_c.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "b200seg_debug_layout": (_c.c_int, [_i32, _i32, _i64, _i32, _c.POINTER(_sz), _i32]),
    "b200seg_sort_scratch_bytes": (_c.c_int, [_i32, _i64, _c.POINTER(_sz)]),
    "b200seg_sort_segments": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _sz, _vp, _vp]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (never build) the shared library; raise loudly if it is not there."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m miccai2021_cataract_semantic_segmentation_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().b200seg_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def set_tuning(**knobs: int):
    """Kernel-selection knobs of the library (see include/b200seg.h); every setting computes the same results."""
    lib = load()
    for k, v in knobs.items():
        check(lib.b200seg_set_tuning(k.encode(), int(v)), f"b200seg_set_tuning({k})")


def label_code(t: torch.Tensor) -> int:
    return {torch.uint8: LABEL_U8, torch.int32: LABEL_I32, torch.int64: LABEL_I64}[t.dtype]


def as_label_tensor(target: torch.Tensor) -> torch.Tensor:
    """Labels are consumed as uint8 / int32 / int64 without a copy; anything else is widened to int64."""
    if target.dtype not in (torch.uint8, torch.int32, torch.int64):
        target = target.to(torch.int64)
    return target.contiguous()


def require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("b200seg kernels run on CUDA tensors only (sm_100a); got a %s tensor. "
                               "There is no CPU fallback." % t.device.type)


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
