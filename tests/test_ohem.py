"""OhemCrossEntropy (reference losses/OhemCrossEntropy.py:8-40): the oracle against vectors produced by
the unmodified reference (tests/golden/make_golden_ohem.py), and the CUDA path against both.
Gates: loss within 1e-5 relative, gradient within 1e-5 of its maximum (the reference's fp32 mean / softmax round
differently on CPU, CUDA and here; the kept set itself is a rank / threshold decision and must agree)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

sys.path.insert(0, GOLDEN_DIR)
from make_golden_ohem import ohem_inputs  # noqa: E402

from oracle import port  # noqa: E402

with open(os.path.join(GOLDEN_DIR, "ohem_manifest.json")) as _f:
    CASES = json.load(_f)["cases"]
RTOL = 1e-5


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "ohem.npz"))


def _inputs(case):
    return ohem_inputs(case["seed"], case["n"], case["c"], case["h"], case["w"], case["config"], case["style"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_vectors(golden, case):
    x, y = _inputs(case)
    loss, grad = port.ohem_with_grad(x, y, thresh=case["thresh"], min_kept=case["min_kept"],
                                     ignore_label=case["ignore_label"])
    ref_l, ref_g = float(golden[case["name"] + "/loss"]), golden[case["name"] + "/grad"]
    assert abs(float(loss) - ref_l) <= 1e-6 * abs(ref_l)
    assert float(np.abs(grad.numpy() - ref_g).max()) <= 1e-7 * float(np.abs(ref_g).max())
    assert int((grad.abs().sum(1) > 0).sum()) == case["kept_pixels"]


def test_host_constructor_mirrors_reference():
    import miccai2021_cataract_semantic_segmentation_b200 as b200
    m = b200.OhemCrossEntropy({"experiment": 3, "min_kept": 0, "thresh": 0.5})
    assert (m.thresh, m.min_kept, m.ignore_label) == (0.5, 1, 25)
    m = b200.OhemCrossEntropy({})
    assert (m.thresh, m.min_kept, m.ignore_label) == (0.7, 100000, -100)
    assert b200.OhemCrossEntropy({"experiment": 1}).ignore_label == -100
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m(torch.zeros(1, 8, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))


# ---- CUDA path -------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def b200():
    import miccai2021_cataract_semantic_segmentation_b200 as pkg
    return pkg


def _check(loss, grad, ref_loss, ref_grad):
    ref_loss = float(ref_loss)
    assert abs(float(loss) - ref_loss) <= RTOL * abs(ref_loss), (float(loss), ref_loss)
    gmax = float(ref_grad.abs().max())
    assert float((grad - ref_grad).abs().max()) <= RTOL * gmax
    assert torch.equal(grad.abs().sum(1) > 0, ref_grad.abs().sum(1) > 0)          # same kept set


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpu_matches_reference_vectors(b200, golden, case):
    x, y = _inputs(case)
    for ldt in (torch.int64, torch.int32):
        xd = x.cuda().requires_grad_(True)
        loss = b200.OhemCrossEntropy(case["config"])(xd, y.cuda().to(ldt))
        loss.backward()
        _check(loss.detach().cpu(), xd.grad.cpu(), golden[case["name"] + "/loss"],
               torch.from_numpy(golden[case["name"] + "/grad"]))


@pytest.mark.gpu
@pytest.mark.parametrize("n,c,h,w,cfg", [
    (2, 25, 136, 240, {"experiment": 3, "min_kept": 20000}),                 # pipelined kernels, thresh decides
    (2, 25, 136, 240, {"experiment": 3, "min_kept": 30000, "thresh": 0.05}),  # order statistic decides
    (4, 17, 64, 96, {"experiment": 2}),                                      # default min_kept >= n_valid: rank n-1
    (1, 8, 135, 240, {"experiment": 1, "min_kept": 5000, "thresh": 0.3}),
    (3, 12, 33, 47, {"min_kept": 1000, "thresh": 0.2}),                      # scalar kernels (class count, odd plane)
    (2, 25, 33, 47, {"experiment": 3, "min_kept": 700, "thresh": 0.1}),
])
def test_gpu_matches_oracle(b200, n, c, h, w, cfg):
    g = torch.Generator().manual_seed(n * 100 + c + h)
    hi = c + 1 if cfg.get("experiment") in (2, 3) else c
    y = torch.randint(0, hi, (n, h, w), generator=g)
    x = torch.randn((n, c, h, w), generator=g) + 3.0 * torch.nn.functional.one_hot(y.clamp(max=c - 1), c) \
        .permute(0, 3, 1, 2).float() * (torch.rand((n, 1, h, w), generator=g) < 0.6).float()
    mod = b200.OhemCrossEntropy(cfg)
    xd = x.cuda().requires_grad_(True)
    loss = mod(xd, y.cuda())
    (2.5 * loss).backward()
    ref_l, ref_g = port.ohem_with_grad(x.cuda(), y.cuda(), thresh=mod.thresh, min_kept=mod.min_kept,
                                       ignore_label=mod.ignore_label)
    _check(loss.detach(), xd.grad, ref_l, 2.5 * ref_g)
    # uint8 labels and a second backward through the same graph
    xd2 = x.cuda().requires_grad_(True)
    l2 = mod(xd2, y.cuda().to(torch.uint8))
    l2.backward(retain_graph=True)
    g1 = xd2.grad.clone()
    xd2.grad = None
    l2.backward()
    assert torch.equal(l2.detach(), loss.detach()) and torch.equal(g1, xd2.grad)
    assert float((g1 * 2.5 - xd.grad).abs().max()) <= 1e-6 * float(xd.grad.abs().max())


@pytest.mark.gpu
def test_gpu_full_size_properties(b200):
    # BASELINE-size batch: the kept set is exactly {p_label < threshold}, and the loss is its mean -log p_label
    n, c, h, w = 8, 25, 544, 960
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda")
    x = torch.randn((n, c, h, w), generator=g, device="cuda")
    x += 4.0 * torch.nn.functional.one_hot(y.clamp(max=c - 1), c).permute(0, 3, 1, 2) * \
        (torch.rand((n, 1, h, w), generator=g, device="cuda") < 0.8)
    x.requires_grad_(True)
    for cfg in ({"experiment": 3}, {"experiment": 3, "min_kept": 2000000, "thresh": 0.1}):
        mod = b200.OhemCrossEntropy(cfg)
        x.grad = None
        loss = mod(x, y)
        loss.backward()
        with torch.no_grad():
            valid = y != c
            p = torch.softmax(x, 1).gather(1, y.clamp(max=c - 1).unsqueeze(1)).squeeze(1)
            pv = p[valid]
            kth = torch.kthvalue(pv, min(mod.min_kept, pv.numel() - 1) + 1).values
            thr = max(float(kth), mod.thresh)
            kept = valid & (p < thr)
            got_kept = x.grad.abs().sum(1) > 0
            assert int((kept ^ got_kept).sum()) <= 2          # softmax bits may differ from ATen's at the boundary
            ref = (-torch.log(p[kept].double())).mean()
            assert abs(float(loss) - float(ref)) <= 1e-5 * float(ref)
            assert float(x.grad.sum(1).abs().max()) < 1e-9    # softmax - onehot sums to zero per pixel


@pytest.mark.gpu
def test_gpu_edge_cases(b200):
    c = 17
    x = torch.randn(1, c, 8, 8, device="cuda", requires_grad=True)
    mod = b200.OhemCrossEntropy({"experiment": 2, "min_kept": 5})
    # every pixel ignored: the reference raises (index -1 of an empty tensor); here the mean over nothing is NaN
    loss = mod(x, torch.full((1, 8, 8), c, device="cuda"))
    loss.backward()
    assert bool(torch.isnan(loss)) and float(x.grad.abs().max()) == 0.0
    # nothing below the threshold -> NaN like torch's mean of an empty selection, zero gradient
    y = torch.randint(0, c, (1, 8, 8), device="cuda")
    conf = (20.0 * torch.nn.functional.one_hot(y, c).permute(0, 3, 1, 2).float()).requires_grad_(True)
    l2 = b200.OhemCrossEntropy({"experiment": 2, "min_kept": 5, "thresh": 0.5})(conf, y)
    l2.backward()
    ref = port.ohem_cross_entropy(conf.detach(), y, thresh=0.5, min_kept=5, ignore_label=c)
    assert bool(torch.isnan(l2)) == bool(torch.isnan(ref))
    # out-of-range label: flagged when validation is on
    bad = y.clone()
    bad[0, 0, 0] = c + 3
    with pytest.raises(RuntimeError, match="out of bounds"):
        b200.OhemCrossEntropy({"experiment": 2, "validate_labels": True})(x, bad)
    # score at a lower resolution than the target: resized like the reference (:23-26)
    xs = torch.randn(2, c, 16, 24, device="cuda")
    yt = torch.randint(0, c + 1, (2, 32, 48), device="cuda")
    got = b200.OhemCrossEntropy({"experiment": 2, "min_kept": 300})(xs, yt)
    up = torch.nn.functional.interpolate(xs, size=(32, 48), mode="bilinear")
    ref = port.ohem_cross_entropy(up, yt, thresh=0.7, min_kept=300, ignore_label=c)
    assert abs(float(got) - float(ref)) <= RTOL * float(ref)
    # empty batch
    l0 = mod(torch.zeros(0, c, 4, 4, device="cuda"), torch.zeros(0, 4, 4, dtype=torch.int64, device="cuda"))
    assert bool(torch.isnan(l0))
