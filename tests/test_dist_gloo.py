"""Multi-process host logic on CPU (gloo, world_size 2): image sharding + the C x C int64 all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from miccai2021_cataract_semantic_segmentation_b200 import dist as bd
    from oracle import port as oracle
    r, w, _ = bd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(11)
    x = torch.randn((n_images, 17, 12, 20), generator=g)
    y = torch.randint(0, 18, (n_images, 12, 20), generator=g)
    lo, hi = bd.shard_range(n_images, rank, world)
    # each rank accumulates the matrix of its own images (the CUDA kernel's role on the GPU box) ...
    cm = oracle.confusion_matrix(x[lo:hi], y[lo:hi]).to(torch.int64) if hi > lo else torch.zeros(17, 17, dtype=torch.int64)
    status = torch.tensor([1 if rank == 1 else 0], dtype=torch.int32)
    if n_images % 2:                                          # both forms of the call
        bd.all_reduce_confusion_matrix(cm, status=status)
    else:
        pending = bd.all_reduce_confusion_matrix(cm, status=status, async_op=True)
        pending.wait()
    # ... and the sum over ranks must be the single-process matrix of the concatenated data
    full = oracle.confusion_matrix(x, y).to(torch.int64)
    assert torch.equal(cm, full)
    assert int(status) == 1                                   # sticky error flag reaches every rank
    # per-image Lovasz shards with no collective: mean of equal shards' means == global per-image mean
    local = oracle.lovasz_softmax(x[lo:hi], y[lo:hi], 2, per_image=True)
    mean = bd.all_reduce_mean(torch.as_tensor(float(local)))
    if n_images % world == 0:
        ref = float(oracle.lovasz_softmax(x, y, 2, per_image=True))
        assert abs(float(mean) - ref) < 1e-6
    # the meter's packed form: matrix + status word in ONE collective
    buf = torch.zeros(17 * 17 + 1, dtype=torch.int64)
    buf[:-1] = oracle.confusion_matrix(x[lo:hi], y[lo:hi]).to(torch.int64).view(-1) if hi > lo else 0
    buf[-1:].view(torch.int32)[0] = 1 if rank == 1 else 0
    if n_images % 2:
        bd.all_reduce_packed(buf)
    else:
        bd.all_reduce_packed(buf, async_op=True).wait()
    assert torch.equal(buf[:-1].view(17, 17), full) and int(buf[-1]) == 1
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), cm.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5])
def test_confusion_matrix_allreduce_world2(tmp_path, n_images):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_images, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "ok0.npy"), np.load(tmp_path / "ok1.npy")
    assert np.array_equal(a, b) and a.sum() > 0


def test_shard_range_partitions():
    from miccai2021_cataract_semantic_segmentation_b200.dist import shard_range
    for n in (0, 1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
