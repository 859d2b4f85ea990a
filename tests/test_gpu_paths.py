"""GPU tests of the kernel-selection paths (``-m gpu``): every variant the library can take for the same call
-- record-driven vs streaming emission, pipelined vs register-tile stats kernel, interleaved vs contiguous tiles,
ballot vs MATCH.ANY ranking -- must produce the same bits, and the per-pixel candidate records the stats kernel
leaves behind must agree with a plain softmax / top-k.  Parity against the oracle is in test_gpu_parity.py; here the
default path (already held to the oracle there) is the reference.
"""
import ctypes

import numpy as np
import pytest
import torch

from test_gpu_parity import CASES, _blocky, _d1

pytestmark = pytest.mark.gpu

DEFAULTS = dict(interleave=1, stats_variant=0, emit_path=0, sort_match=2, sort_path=0, dbg=0, pdl=1)
VARIANTS = [
    ("emit_records", dict(emit_path=1)),
    ("emit_stream", dict(emit_path=2)),
    ("stats_register_tile", dict(stats_variant=1)),
    ("stats_128x2", dict(stats_variant=2)),
    ("stats_128x4", dict(stats_variant=3)),
    ("stats_128x3", dict(stats_variant=5)),
    ("stats_512x2", dict(stats_variant=6)),
    ("stats_tma", dict(stats_variant=7)),
    ("contiguous_tiles", dict(interleave=0)),
    ("rank_ballots", dict(sort_match=0)),
    ("rank_match", dict(sort_match=1)),
    ("sort_lsd", dict(sort_path=1)),
    ("sort_lsd_ballots", dict(sort_path=1, sort_match=0)),
    ("sort_lsd_stream", dict(sort_path=1, emit_path=2)),
    ("plain_launches", dict(pdl=0)),
]
PATH_CASES = [c for c in CASES if c[0] in ("d1_c8_flat", "d1_c17_per_image", "d1_c25_flat", "d1_c25_flat_ignore",
                                           "d1_c25_all", "d2_c25_flat", "d2_c17_per_image_ignore", "d2_c25_list")]


@pytest.fixture(scope="module")
def b200():
    assert torch.cuda.is_available()
    import miccai2021_cataract_semantic_segmentation_b200 as pkg
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    _native.load()
    return pkg


@pytest.fixture(autouse=True)
def _reset_tuning():
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    _native.set_tuning(**DEFAULTS)
    yield
    _native.set_tuning(**DEFAULTS)


def _run(b200, x, y, cfg, exp, c):
    meter = b200.SegmentationMeter(exp, c)
    xd = x.cuda().requires_grad_(True)
    loss = b200.LovaszSoftmaxWithMetrics(cfg, meter)(xd, y.cuda())
    loss.backward()
    meter.check()
    return float(loss), xd.grad.clone(), meter.cm.clone()


@pytest.mark.parametrize("case", PATH_CASES, ids=[c[0] for c in PATH_CASES])
def test_every_kernel_path_gives_the_same_bits(b200, case):
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    name, builder, (n, c, h, w), exp, extra = case
    x, y = builder(n, c, h, w, seed=77 + n * c, with_ignore=exp != 1)
    cfg = {"experiment": exp, **extra}
    loss0, grad0, cm0 = _run(b200, x, y, cfg, exp, c)
    assert np.isfinite(loss0)
    for tag, knobs in VARIANTS:
        _native.set_tuning(**DEFAULTS)
        _native.set_tuning(**knobs)
        loss, grad, cm = _run(b200, x, y, cfg, exp, c)
        assert loss == loss0, f"{tag}: loss {loss} vs {loss0}"
        assert torch.equal(grad, grad0), f"{tag}: gradients differ, max {float((grad - grad0).abs().max())}"
        assert torch.equal(cm, cm0), f"{tag}: confusion matrix differs"


def test_record_emission_with_many_candidates_per_tile(b200):
    """All-equal logits: every (pixel, class) pair is a candidate (25 per pixel, far beyond the 4096-entry stage of the
    record kernel, so every tile takes several passes over the class ranges) and every key ties."""
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    from oracle import port
    n, c, h, w = 1, 25, 64, 96
    g = torch.Generator().manual_seed(5)
    x = torch.zeros((n, c, h, w))
    y = torch.randint(0, c + 1, (n, h, w), generator=g)
    ref_loss, ref_grad = port.lovasz_softmax_with_grad(x.cuda(), y.cuda(), 3)
    outs = []
    for path in (1, 2):
        _native.set_tuning(emit_path=path)
        xd = x.cuda().requires_grad_(True)
        loss = b200.LovaszSoftmax({"experiment": 3})(xd, y.cuda())
        loss.backward()
        outs.append((float(loss), xd.grad.clone()))
        assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
        assert float((xd.grad - ref_grad).abs().max()) <= 1e-5 * float(ref_grad.abs().max())
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("c,exp", [(25, 3), (17, 2), (8, 1)])
def test_candidate_records_match_softmax_topk(b200, c, exp):
    """rec16 = {key of 1 - p_label, p of the two most probable other classes, upper bound of the third}."""
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    lib = _native.load()
    n, h, w = 2, 96, 160
    hw, P = h * w, n * h * w
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((n, c, h, w), generator=g, device="cuda")
    y = torch.randint(0, c + (exp != 1), (n, h, w), generator=g, device="cuda")
    nb = _native._sz(0)
    _native.check(lib.b200seg_lovasz_workspace_bytes(n, c, hw, 0, nb), "workspace")
    ws = torch.zeros(nb.value, dtype=torch.uint8, device="cuda")
    loss = torch.empty((), device="cuda")
    _native.check(lib.b200seg_lovasz_forward(x.data_ptr(), y.data_ptr(), _native.LABEL_I64, n, c, hw, 0, _native.NO_LABEL,
                                             0, (1 << c) - 1, 1, ws.data_ptr(), ws.numel(), loss.data_ptr(), None,
                                             _native.NO_LABEL, None, torch.cuda.current_stream().cuda_stream), "forward")
    torch.cuda.synchronize()
    offs = (ctypes.c_size_t * 8)()
    _native.check(lib.b200seg_debug_layout(n, c, hw, 0, offs, 8), "layout")

    def view(off, nbytes, dtype):
        return ws[off:off + nbytes].view(dtype)

    rec16 = view(offs[4], 16 * P, torch.int32).view(P, 4)
    rec4 = view(offs[5], 4 * P, torch.int32)
    lab8 = view(offs[2], P, torch.uint8)
    cmask = view(offs[3], 4 * P, torch.int32)
    thr = view(offs[6], 4 * c, torch.float32)
    p1, p2, p3 = (rec16[:, i].view(torch.float32) for i in (1, 2, 3))
    prob = torch.softmax(x, 1).permute(0, 2, 3, 1).reshape(P, c)
    yy = y.view(-1)
    valid = yy < c
    assert torch.equal(lab8.long(), torch.where(valid, yy, torch.full_like(yy, 255)))
    assert torch.equal(rec4 & 255, lab8.int())
    # own-class key: 0x3F800000 - bits(1 - p_label)
    own = prob[valid.nonzero().squeeze(1), yy[valid]]
    key = 0x3F800000 - (1.0 - own).view(torch.int32)
    assert int((rec16[valid, 0] - key).abs().max()) <= 2          # ATen's softmax and ours agree to the last ulp or two
    other = prob.clone()
    other[valid.nonzero().squeeze(1), yy[valid]] = -1.0
    # recorded classes carry their exact probability; the choice among near-ties (within 255 ulp of the exponential,
    # the class index rides in the low byte during the top-3 search) is free, the guard covers whatever was not chosen
    ar = torch.arange(P, device="cuda")
    c1 = ((rec4 >> 8) & 31).long()
    c2 = ((rec4 >> 16) & 31).long()
    assert bool((c1 != c2).all()) and bool((c1[valid] != yy[valid]).all()) and bool((c2[valid] != yy[valid]).all())
    assert float((p1 - other[ar, c1]).abs().max()) <= 2e-7
    assert float((p2 - other[ar, c2]).abs().max()) <= 2e-7
    top = other.topk(3, 1)
    assert float((top.values[:, 0] - p1).max()) <= 2e-5 and float((top.values[:, 1] - p2).max()) <= 2e-5
    rest = other.clone()
    rest[ar, c1] = -1.0
    rest[ar, c2] = -1.0
    third = rest.max(1).values
    assert bool((p3 >= third - 1e-7).all())                         # a bound: never below any unrecorded class
    assert float((p3 - third).max()) <= 2e-5                        # ... and a tight one
    # the candidate mask the emission left for the backward pass: background classes with p >= threshold
    expect = (other >= thr.view(1, c)).int()
    bits = torch.stack([(cmask >> k) & 1 for k in range(c)], 1)
    near = ((other - thr.view(1, c)).abs() <= 2e-7).any(1)          # one-ulp disagreements with ATen at the threshold
    assert torch.equal(bits[~near], expect[~near])


def test_register_constant_exponential_is_expf_bit_for_bit(b200):
    """stats_kernel_async forms exp(z - max) with expf()'s own instruction sequence but register-held constants."""
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    lib = _native.load()
    g = torch.Generator(device="cuda").manual_seed(11)
    parts = [-torch.rand(1 << 22, generator=g, device="cuda") * 110.0,            # the softmax range, past underflow
             -torch.rand(1 << 20, generator=g, device="cuda") * 1e-3,               # just below zero
             torch.randn(1 << 20, generator=g, device="cuda") * 50.0,               # both signs, overflow included
             torch.tensor([0.0, -0.0, -87.3, -88.8, -103.9, -104.1, -1e30, 88.7, 89.0, float("-inf"), float("inf"),
                           float("nan"), -1e-45, 1e-45], device="cuda")]
    x = torch.cat(parts).contiguous()
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    _native.check(lib.b200seg_debug_exp_mismatches(x.data_ptr(), x.numel(), bad.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream), "exp check")
    assert int(bad.item()) == 0
