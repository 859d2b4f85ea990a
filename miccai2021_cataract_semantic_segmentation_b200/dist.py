"""Data-parallel plumbing: one process per GPU, images sharded across ranks (the reference itself is single-GPU,
managers/BaseManager.py:83-86; this layer is new).

* per-image Lovasz and the confusion matrix shard by image with no data-path collective;
* flat (batch-level) Lovasz stays rank-local, exactly what per-process DDP would compute with the reference;
* the C x C int64 matrices are summed with ONE all-reduce (NCCL over NVLink on GPUs, gloo on CPU in tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of items for this rank (first ``n_items % world`` ranks get one extra)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class _Pending:
    """Handle of the asynchronous all-reduces of one meter: ``wait()`` makes the current stream wait for them."""

    def __init__(self, works):
        self.works = [w for w in works if w is not None]

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


def all_reduce_confusion_matrix(cm: torch.Tensor, group=None, status: torch.Tensor | None = None, async_op=False):
    """In-place SUM of the int64 C x C matrix over ranks (<= 5 KB at C = 25: latency-bound, one call per step or
    per sweep).  ``status`` (the sticky label-range flag) is MAX-reduced so every rank raises together.
    ``async_op=True`` returns a handle at once: the collectives run on the backend's stream, ordered after the work
    already queued on the current stream (the forward pass that fills the matrix), so they overlap whatever is queued
    next (the backward pass); call ``.wait()`` before reading the matrix."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return _Pending([]) if async_op else None
    assert cm.dtype == torch.int64
    works = [dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=group, async_op=async_op)]
    if status is not None:
        works.append(dist.all_reduce(status, op=dist.ReduceOp.MAX, group=group, async_op=async_op))
    return _Pending(works) if async_op else None


def all_reduce_packed(buf: torch.Tensor, group=None, async_op=False):
    """ONE in-place SUM all-reduce of a meter's packed int64 buffer: the C x C matrix followed by the word that holds the
    sticky status flags (SegmentationMeter lays them out that way).  After the sum the status word is non-zero on every
    rank iff some rank flagged an out-of-range label, so all ranks raise together -- one collective per step instead of a
    SUM for the matrix plus a MAX for the flag."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return _Pending([]) if async_op else None
    assert buf.dtype == torch.int64 and buf.is_contiguous()
    work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return _Pending([work]) if async_op else None


def all_reduce_mean(value: torch.Tensor, group=None):
    """Mean of a per-rank scalar (e.g. the rank-local loss, for logging).  Not needed for training: DDP averages
    the model gradients, and with equal shards the mean of rank-local per-image losses is the global mean."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    out = value.detach().clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out / dist.get_world_size(group)
