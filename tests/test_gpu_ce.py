"""GPU tests (``-m gpu``) of the fused Lovasz-Softmax + cross-entropy pass (reference
losses/LossWrapper.py:17-24,43-73): total loss and logit gradient against the oracle, the pair against the two
separate evaluations, the fallback for shapes the pipelined kernels do not cover."""
import numpy as np
import pytest
import torch

from conftest import grad_err, rel_err
from test_gpu_parity import _blocky, _d1

pytestmark = pytest.mark.gpu

CASES = [
    # name, builder, (n, c, h, w), experiment, lovasz config, (w_ce, w_lovasz)
    ("d1_c25_flat", _d1, (2, 25, 128, 240), 3, {}, (1.0, 1.0)),
    ("d1_c17_per_image", _d1, (3, 17, 128, 240), 2, {"per_image": True}, (0.7, 1.3)),
    ("d1_c8_flat", _d1, (2, 8, 128, 240), 1, {}, (1.0, 0.5)),                     # ignore_index = -100: every pixel counts
    ("d1_c25_filter_ignore", _d1, (2, 25, 96, 192), 3, {"classes_to_ignore": 25}, (1.0, 1.0)),
    ("d2_c25_flat", _blocky, (2, 25, 128, 240), 3, {}, (2.0, 1.0)),
    ("d2_c25_list", _blocky, (2, 25, 96, 160), 3, {"classes_to_consider": [0, 1, 2, 7]}, (1.0, 1.0)),   # CE-only classes
    ("d1_c5_generic", _d1, (2, 5, 64, 96), 1, {}, (1.0, 1.0)),                    # no templated kernel: torch CE on the side
    ("d1_c25_odd_plane", _d1, (1, 25, 61, 97), 3, {}, (1.0, 1.0)),                # plane % 16 != 0: same fallback
]


@pytest.fixture(scope="module")
def b200():
    assert torch.cuda.is_available()
    import miccai2021_cataract_semantic_segmentation_b200 as pkg
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    _native.load()
    return pkg


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_loss_wrapper_pair_matches_oracle(b200, case):
    from oracle import port
    name, builder, (n, c, h, w), exp, extra, (w_ce, w_lov) = case
    x, y = builder(n, c, h, w, seed=4321 + n * c, with_ignore=exp != 1)
    cfg = {"losses": {"CrossEntropyLoss": w_ce, "LovaszSoftmax": w_lov}, "experiment": exp, "device": "cuda", **extra}
    wrapper = b200.LossWrapper(cfg)
    xd = x.cuda().requires_grad_(True)
    total = wrapper(None, xd, y.cuda())
    total.backward()
    kw = dict(per_image=extra.get("per_image", False), classes_to_ignore=extra.get("classes_to_ignore"),
              classes_to_consider=extra.get("classes_to_consider", "present"))
    xr = x.cuda().requires_grad_(True)
    ref = port.loss_wrapper_pair(xr, y.cuda(), exp, w_ce, w_lov, **kw)
    ref.backward()
    assert rel_err(float(total), float(ref)) <= 1e-5
    assert grad_err(xd.grad.cpu().numpy(), xr.grad.cpu().numpy()) <= 1e-5
    gmax = float(xr.grad.abs().max())
    assert np.allclose(xd.grad.cpu().numpy(), xr.grad.cpu().numpy(), rtol=1e-5, atol=1e-5 * gmax)
    # the weighted parts the reference keeps for logging (LossWrapper.py:71-72)
    ce_ref = float(port.cross_entropy(x.cuda(), y.cuda(), exp)) * w_ce
    assert rel_err(float(wrapper.loss_vals["CrossEntropyLoss"]), ce_ref) <= 1e-5
    assert set(wrapper.loss_vals) == {"CrossEntropyLoss", "LovaszSoftmax"}


def test_pair_equals_the_two_separate_evaluations(b200):
    n, c, h, w, exp = 2, 25, 128, 240, 3
    x, y = _d1(n, c, h, w, seed=99, with_ignore=True)
    meter = b200.SegmentationMeter(exp, c)
    pair = b200.LovaszSoftmaxCE({"experiment": exp}, meter)
    xd = x.cuda().requires_grad_(True)
    lov, ce = pair(xd, y.cuda())
    # Lovasz part alone: same bits as the plain module, forward and backward
    (g_lov,) = torch.autograd.grad(lov, xd, retain_graph=True)
    x2 = x.cuda().requires_grad_(True)
    lov2 = b200.LovaszSoftmax({"experiment": exp})(x2, y.cuda())
    lov2.backward()
    assert float(lov) == float(lov2) and torch.equal(g_lov, x2.grad)
    # cross-entropy part alone against torch
    (g_ce,) = torch.autograd.grad(ce, xd)
    x3 = x.cuda().requires_grad_(True)
    ce3 = torch.nn.functional.cross_entropy(x3, y.cuda(), ignore_index=25)
    ce3.backward()
    assert rel_err(float(ce), float(ce3)) <= 1e-6
    assert float((g_ce - x3.grad).abs().max()) <= 1e-5 * float(x3.grad.abs().max())
    # the confusion matrix rode along
    assert torch.equal(meter.cm, b200.t_get_confusion_matrix(x.cuda(), y.cuda()))
    meter.check()


def test_cross_entropy_edge_cases(b200):
    c, exp = 25, 3
    g = torch.Generator().manual_seed(1)
    x = torch.randn((1, c, 32, 64), generator=g)
    pair = b200.LovaszSoftmaxCE({"experiment": exp})
    # every pixel ignored: torch's mean over nothing is NaN; the Lovasz term has no class present and is 0
    y = torch.full((1, 32, 64), 25)
    lov, ce = pair(x.cuda(), y.cuda())
    assert float(lov) == 0.0 and np.isnan(float(ce))
    # a label that is neither a class nor the ignore index: torch raises, here the status word is set
    y = torch.randint(0, c, (1, 32, 64), generator=g)
    y[0, 3, 5] = 77
    meter = b200.SegmentationMeter(exp, c)
    pair_m = b200.LovaszSoftmaxCE({"experiment": exp}, meter)
    pair_m(x.cuda(), y.cuda())
    with pytest.raises(RuntimeError):
        meter.check()


def test_two_scale_loss_two_streams_matches_oracle(b200):
    """TwoScaleLoss (losses/TwoScaleLoss.py:43-52): the low-resolution intermediate head is upsampled bilinearly, both
    heads get the same target; here they run on two CUDA streams."""
    from oracle import port
    n, c, exp = 2, 25, 3
    g = torch.Generator().manual_seed(8)
    x_final = torch.randn((n, c, 64, 128), generator=g)
    x_interm = torch.randn((n, c, 32, 64), generator=g)                    # upsampled inside forward
    y = torch.randint(0, c + 1, (n, 64, 128), generator=g)
    cfg = {"interm": {"name": "LovaszSoftmax", "args": []}, "final": {"name": "LovaszSoftmax", "args": [], "weight": 1.0},
           "experiment": exp}
    mod = b200.TwoScaleLoss(cfg)
    assert mod.w_interm == 0.4 and mod.w_final == 1.0 and mod.ignore_label == 25
    xf, xi = x_final.cuda().requires_grad_(True), x_interm.cuda().requires_grad_(True)
    for _ in range(2):                                                      # second round reuses the side stream
        xf.grad = xi.grad = None
        loss = mod(xi, xf, y.cuda())
        loss.backward()
    rf, ri = x_final.cuda().requires_grad_(True), x_interm.cuda().requires_grad_(True)
    up = torch.nn.functional.interpolate(ri, size=(64, 128), mode="bilinear")
    ref = port.lovasz_softmax(rf, y.cuda(), exp) * 1.0 + port.lovasz_softmax(up, y.cuda(), exp) * 0.4
    ref.backward()
    assert rel_err(float(loss), float(ref)) <= 1e-5
    assert grad_err(xf.grad.cpu().numpy(), rf.grad.cpu().numpy()) <= 1e-5
    assert grad_err(xi.grad.cpu().numpy(), ri.grad.cpu().numpy()) <= 1e-5
    with pytest.raises(NotImplementedError):
        b200.TwoScaleLoss({"interm": {"name": "LovaszSoftmax", "args": []}, "final": {"name": "CrossEntropyLoss", "args": []},
                           "experiment": exp})


def test_two_streams_with_big_sort_segments(b200):
    """Two heads on two streams while both sorts have segments above the local-scan limit (confident logits: the grid
    barriers of the scatter kernels run concurrently): results must equal the one-stream evaluation bit for bit."""
    from test_gpu_parity import _blocky
    n, c, h, w, exp = 2, 25, 540, 960, 3
    xa, y = _blocky(n, c, h, w, seed=21, with_ignore=True)
    xb, _ = _blocky(n, c, h, w, seed=22, with_ignore=True)
    cfg = lambda: {"interm": {"name": "LovaszSoftmax", "args": [], "weight": 0.4},
                   "final": {"name": "LovaszSoftmax", "args": [], "weight": 1.0}, "experiment": exp}
    two = b200.TwoScaleLoss(cfg())
    la, lb = b200.LovaszSoftmax({"experiment": exp}), b200.LovaszSoftmax({"experiment": exp})
    a1, b1 = xa.cuda().requires_grad_(True), xb.cuda().requires_grad_(True)
    yd = y.cuda()
    ref = lb(b1, yd) * 1.0 + la(a1, yd) * 0.4
    ref.backward()
    for _ in range(3):
        a2, b2 = xa.cuda().requires_grad_(True), xb.cuda().requires_grad_(True)
        out = two(a2, b2, yd)
        out.backward()
        torch.cuda.synchronize()
        assert float(out) == float(ref)
        assert torch.equal(a2.grad, a1.grad) and torch.equal(b2.grad, b1.grad)


def test_iou_tracker_mirrors_without_blocking(b200):
    """Device-side EMA of the per-class IoU with an asynchronous pinned mirror."""
    c, exp = 25, 3
    gen = torch.Generator(device="cuda").manual_seed(2)
    meter = b200.SegmentationMeter(exp, c)
    tr = b200.IoUTracker(c, alpha=0.25, device="cuda")
    ref = np.zeros(c, np.float32)
    for _ in range(4):
        x = torch.randn((1, c, 64, 96), generator=gen, device="cuda")
        y = torch.randint(0, c + 1, (1, 64, 96), generator=gen, device="cuda")
        meter.reset()
        meter.update(x, y)
        iou, _ = meter.summary()
        tr.update(iou)
        ref = 0.75 * ref + 0.25 * iou.cpu().numpy()
        stale = tr.host_values()                               # never blocks: this or the previous step's value
        assert stale.shape == (c,)
    assert np.allclose(tr.host_values(wait=True), ref, rtol=1e-6, atol=1e-7)


def test_pair_outputs_survive_in_place_scaling_and_anomaly_mode(b200):
    """The reference scales losses in place (`loss *= w`, LossWrapper.py:71) and runs under
    torch.autograd.set_detect_anomaly(True) (main.py:8)."""
    n, c, exp = 1, 25, 3
    x, y = _d1(n, c, 64, 96, seed=3, with_ignore=True)
    with torch.autograd.detect_anomaly():
        xd = x.cuda().requires_grad_(True)
        lov, ce = b200.LovaszSoftmaxCE({"experiment": exp})(xd, y.cuda())
        lov *= 0.5
        ce *= 2.0
        total = lov + ce
        total.backward()
    xr = x.cuda().requires_grad_(True)
    lov2, ce2 = b200.LovaszSoftmaxCE({"experiment": exp})(xr, y.cuda())
    (lov2 * 0.5 + ce2 * 2.0).backward()
    assert torch.equal(xd.grad, xr.grad) and torch.isfinite(xd.grad).all()


def test_fused_pair_raises_lazily_on_an_out_of_range_label(b200):
    """nn.CrossEntropyLoss raises on a target outside [0, C) that is not ignore_index; the fused pass flags it and the
    module raises at its next call (or at check()), never synchronising on the step it has just queued."""
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((1, 8, 64, 96), generator=g, device="cuda")
    y = torch.randint(0, 8, (1, 64, 96), generator=g, device="cuda")
    mod = b200.LovaszSoftmaxCE({"experiment": 1})
    mod(x, y)
    mod.check()                                                    # clean labels: nothing to report
    bad = y.clone()
    bad[0, 3, 5] = 255                                             # experiment 1 ignores -100 only: 255 is out of range
    mod(x, bad)
    with pytest.raises(IndexError):
        mod(x, y)                                                  # the flag of the previous call
    mod(x, y)
    mod.check()                                                    # cleared after raising


def test_async_to_numpy_hands_out_the_previous_step_without_blocking(b200):
    """install(async_iou=True) semantics: small float CUDA vectors come back one call late from a pinned mirror; everything
    else takes the blocking path."""
    calls = []

    def blocking(t):
        calls.append(tuple(t.shape))
        return t.detach().cpu().numpy()

    f = b200.AsyncToNumpy(blocking)
    vecs = [torch.full((25,), float(i), device="cuda") for i in range(5)]
    got = []
    for v in vecs:
        got.append(f(v).copy())
        torch.cuda.synchronize()                                    # (so that "the previous copy has completed" is deterministic here)
    assert [float(g[0]) for g in got] == [0.0, 0.0, 1.0, 2.0, 3.0] and calls == []
    big = torch.zeros((3, 8, 8), device="cuda")
    assert f(big).shape == (3, 8, 8) and calls == [(3, 8, 8)]       # not a small vector: the reference's own path


def test_best_model_tracker_on_device(b200):
    tr = b200.BestModelTracker()
    assert tr.poll(wait=True) == (False, [0.0, 0.0, 0.0, 0.0])
    tr.update(torch.tensor(0.71946, device="cuda"), torch.tensor(0.8, device="cuda"), torch.tensor(0.6, device="cuda"),
              torch.tensor(0.3, device="cuda"))
    flag, vals = tr.poll(wait=True)
    assert flag and vals[0] == round(float(np.float32(0.71946)), 4) == 0.7195
    tr.update(torch.tensor(0.71949, device="cuda"), torch.tensor(0.9, device="cuda"), torch.tensor(0.9, device="cuda"),
              torch.tensor(0.9, device="cuda"))
    assert tr.poll(wait=True)[0] is False and float(tr.best[1]) == 0.8
