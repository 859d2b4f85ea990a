"""Drop-in ``OhemCrossEntropy`` (reference: losses/OhemCrossEntropy.py:8-40) backed by the sm_100a kernels.

Same constructor (config keys ``thresh``, ``min_kept``, ``experiment``), same ``forward(score, target)``,
differentiable w.r.t. ``score``.  The reference sorts all N*H*W label probabilities to read one order statistic;
here that is a radix select inside the library (csrc/ohem.cu).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native
from .class_info import CLASS_INFO


class _OhemFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_label, thresh, min_kept, status):
        with torch.cuda.device(logits.device):
            lib = _native.load()
            n, c, h, w = logits.shape
            nbytes = _native._sz(0)
            _native.check(lib.b200seg_ohem_workspace_bytes(n, h * w, nbytes), "b200seg_ohem_workspace_bytes")
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=logits.device)
            loss = torch.empty((), dtype=torch.float32, device=logits.device)
            _native.check(lib.b200seg_ohem_ce_forward(
                logits.data_ptr(), target.data_ptr(), _native.label_code(target), n, c, h * w, ignore_label,
                float(thresh), int(min_kept), ws.data_ptr(), ws.numel(), loss.data_ptr(), status.data_ptr(),
                _native.stream_ptr(logits.device)), "b200seg_ohem_ce_forward")
            if ctx.needs_input_grad[0]:
                ctx.save_for_backward(logits, target, ws)
                ctx.ignore_label = ignore_label
            return loss

    @staticmethod
    def backward(ctx, grad_out):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            logits, target, ws = ctx.saved_tensors
            n, c, h, w = logits.shape
            go = grad_out.detach().to(torch.float32).contiguous()
            dlogits = torch.empty_like(logits)
            _native.check(_native.load().b200seg_ohem_ce_backward(
                logits.data_ptr(), target.data_ptr(), _native.label_code(target), n, c, h * w, ctx.ignore_label,
                ws.data_ptr(), ws.numel(), go.data_ptr(), dlogits.data_ptr(), _native.stream_ptr(logits.device)),
                "b200seg_ohem_ce_backward")
            return dlogits, None, None, None, None, None


def ohem_cross_entropy(score: torch.Tensor, target: torch.Tensor, thresh: float = 0.7, min_kept: int = 100000,
                       ignore_label: int = -100, validate: bool = False) -> torch.Tensor:
    """Functional form of OhemCrossEntropy.forward (losses/OhemCrossEntropy.py:22-40)."""
    _native.require_cuda(score, target)
    if score.dim() != 4 or target.dim() != 3:
        raise ValueError("score must be [N, C, H, W] and target [N, H, W]")
    h, w = target.size(1), target.size(2)
    if score.size(2) != h or score.size(3) != w:            # :23-26 (F.upsample == interpolate, align_corners=False)
        score = F.interpolate(score, size=(h, w), mode='bilinear')
    logits = (score if score.dtype == torch.float32 else score.float()).contiguous()
    tgt = _native.as_label_tensor(target.detach())
    status = torch.zeros(1, dtype=torch.int32, device=logits.device)
    loss = _OhemFunction.apply(logits, tgt, int(ignore_label), float(thresh), int(min_kept), status)
    if validate:                                            # synchronises; the reference's CE asserts on the device
        if int(status.item()) & _native.STATUS_LABEL_OOB:
            raise RuntimeError("Target out of bounds for OhemCrossEntropy")
    return loss


class OhemCrossEntropy(nn.Module):
    def __init__(self, config):
        super().__init__()
        # same keys and defaults as OhemCrossEntropy.py:11-18: thresh 0.7, min_kept 100000 (at least 1), and the ignore
        # label is the last entry of the experiment's class table for experiments 2 / 3, otherwise nothing is ignored
        self.thresh = config.get('thresh', 0.7)
        self.min_kept = max(1, config['min_kept']) if 'min_kept' in config else 100000
        experiment = config.get('experiment')
        self.ignore_label = len(CLASS_INFO[experiment][1]) - 1 if experiment in (2, 3) else -100
        self.validate = bool(config.get('validate_labels', False))

    def forward(self, score, target, **kwargs):
        return ohem_cross_entropy(score, target, self.thresh, self.min_kept, self.ignore_label, self.validate)
