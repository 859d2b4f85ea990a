"""Windowed mean IoU (reference utils/torch_utils.py:189-218): the oracle against vectors produced by the
unmodified reference (tests/golden/make_golden_sliding.py), and the CUDA path against both."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

sys.path.insert(0, GOLDEN_DIR)
from make_golden_sliding import sliding_inputs  # noqa: E402

from oracle import port  # noqa: E402

with open(os.path.join(GOLDEN_DIR, "sliding_manifest.json")) as _f:
    CASES = json.load(_f)["cases"]
TOL = 1e-6          # class mean summed in a different order than torch.mean; integer counts are exact


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "sliding.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_vectors(golden, case):
    x, y = sliding_inputs(case["seed"], case["n"], case["c"], case["h"], case["w"], case["style"])
    for key, full in (("windows", False), ("full", True)):
        got = port.sliding_miou(x, y, case["kernel"], case["stride"], original_size=full).numpy()
        assert np.array_equal(got, golden[f"{case['name']}/{key}"])


def test_oracle_argument_errors():
    x = torch.zeros(1, 8, 9, 9)
    with pytest.raises(AssertionError):
        port.sliding_miou(x, torch.zeros(1, 9, 9, dtype=torch.int64), 4, 2)
    with pytest.raises(RuntimeError):
        port.sliding_miou(x, torch.full((1, 9, 9), 8), 3, 2)


# ---- CUDA path -------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def b200():
    import miccai2021_cataract_semantic_segmentation_b200 as pkg
    return pkg


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpu_matches_reference_vectors(b200, golden, case):
    x, y = sliding_inputs(case["seed"], case["n"], case["c"], case["h"], case["w"], case["style"])
    for ldt in (torch.int64, torch.int32, torch.uint8):
        for key, full in (("windows", False), ("full", True)):
            got = b200.sliding_miou(x.cuda(), y.cuda().to(ldt), case["kernel"], case["stride"], original_size=full)
            ref = golden[f"{case['name']}/{key}"]
            assert tuple(got.shape) == ref.shape and got.dtype == torch.float32
            assert float(np.abs(got.cpu().numpy() - ref).max()) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("n,c,h,w,k,s", [(2, 25, 64, 96, 7, 4), (1, 17, 135, 240, 7, 4), (2, 8, 40, 64, 11, 3),
                                         (1, 12, 31, 45, 5, 2), (3, 25, 33, 47, 7, 4), (1, 32, 16, 16, 15, 1)])
def test_gpu_matches_oracle(b200, n, c, h, w, k, s):
    g = torch.Generator().manual_seed(n * 1000 + c * 10 + k)
    coarse = torch.randint(0, c, (n, (h + 3) // 4, (w + 3) // 4), generator=g)
    y = coarse.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :h, :w].contiguous()
    x = 3.0 * torch.nn.functional.one_hot(y, c).permute(0, 3, 1, 2).float() + 2.0 * torch.randn((n, c, h, w), generator=g)
    ref = port.sliding_miou(x, y, k, s, original_size=False)
    got = b200.sliding_miou(x.cuda(), y.cuda(), k, s, original_size=False)
    assert float((got.cpu() - ref).abs().max()) <= TOL
    # offset views (pointer alignment falls back to the scalar class-map kernel) give the same map
    xo = torch.zeros(n * c * h * w + 1).cuda()[1:].view(n, c, h, w).copy_(x)
    got2 = b200.sliding_miou(xo, y.cuda(), k, s, original_size=False)
    assert torch.equal(got, got2)


@pytest.mark.gpu
def test_gpu_full_size_frame_properties(b200):
    # BASELINE-size frames: a perfect prediction scores 1 everywhere; shifting the labels of one window's worth of
    # pixels changes only the windows that overlap them
    n, c, h, w, k, s = 2, 25, 544, 960, 7, 4
    g = torch.Generator().manual_seed(5)
    y = torch.randint(0, c, (n, h // 8, w // 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2).cuda()
    x = torch.nn.functional.one_hot(y, c).permute(0, 3, 1, 2).float().contiguous()
    m = b200.sliding_miou(x, y, k, s, original_size=False)
    assert tuple(m.shape) == (n, (h - k) // s + 1, (w - k) // s + 1)
    assert bool((m == 1.0).all())
    y2 = y.clone()
    y2[0, 100:107, 200:207] = (y2[0, 100:107, 200:207] + 1) % c
    m2 = b200.sliding_miou(x, y2, k, s, original_size=False)
    changed = (m2 != m).nonzero()
    assert changed.numel() > 0 and bool((changed[:, 0] == 0).all())
    assert int(changed[:, 1].min()) * s + k > 100 and int(changed[:, 1].max()) * s < 107
    assert int(changed[:, 2].min()) * s + k > 200 and int(changed[:, 2].max()) * s < 207
    full = b200.sliding_miou(x, y2, k, s)
    assert tuple(full.shape) == (n, h, w)
    assert float(full[:, :k // 2].abs().max()) == 0.0 and torch.equal(full[0, 3:7, 3:7], m2[0, 0, 0].expand(4, 4))


@pytest.mark.gpu
def test_gpu_argument_errors(b200):
    x = torch.zeros(1, 8, 9, 9, device="cuda")
    y = torch.zeros(1, 9, 9, dtype=torch.int64, device="cuda")
    with pytest.raises(AssertionError):
        b200.sliding_miou(x, y, 4, 2)
    with pytest.raises(RuntimeError, match="Class values must be smaller"):
        b200.sliding_miou(x, torch.full((1, 9, 9), 8, device="cuda"), 3, 2)
    with pytest.raises(RuntimeError):
        b200.sliding_miou(x, y, 11, 2)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        b200.sliding_miou(x.cpu(), y.cpu(), 3, 2)
