"""GPU tests (``-m gpu``) of the fused bilinear upsampling + Lovasz-Softmax path (reference
models/OCR.py:126-131, models/DeepLabv3Plus.py:65-68 followed by losses/LovaszSoftmax.py): the in-kernel interpolation
against ATen bit for bit, then loss / confusion matrix / low-resolution gradient against F.interpolate + the full-resolution
path and against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import grad_err, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    assert torch.cuda.is_available()
    import miccai2021_cataract_semantic_segmentation_b200 as pkg
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    _native.load()
    return pkg


def _up(low, size):
    return F.interpolate(low, size=size, mode="bilinear", align_corners=True)


@pytest.mark.parametrize("shape", [(2, 25, 68, 120, 544, 960), (2, 17, 136, 240, 544, 960), (1, 8, 68, 120, 540, 960),
                                   (3, 5, 7, 9, 33, 64), (1, 3, 1, 5, 4, 32), (1, 2, 5, 1, 9, 32), (1, 4, 13, 17, 13, 17),
                                   (1, 4, 40, 50, 30, 32)])
def test_interpolation_is_atens_bit_for_bit(b200, shape):
    """the arithmetic of csrc/upsample.cuh (what every fused kernel evaluates) == upsample_bilinear2d on the same device"""
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    n, c, h, w, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    low = torch.randn((n, c, h, w), generator=g, device="cuda") * 4
    out = torch.empty((n, c, H, W), device="cuda")
    _native.check(_native.load().b200seg_debug_upsample(low.data_ptr(), n * c, h, w, H, W, out.data_ptr(), -1,
                                                        torch.cuda.current_stream().cuda_stream), "debug_upsample")
    assert torch.equal(out.view(torch.int32), _up(low, (H, W)).view(torch.int32))


def _inputs(n, c, h, w, H, W, seed, dist, with_ignore):
    g = torch.Generator().manual_seed(seed)
    if dist == "d1":
        low = torch.randn((n, c, h, w), generator=g) * 2
        y = torch.randint(0, c + 1 if with_ignore else c, (n, H, W), generator=g)
    else:   # trained-like: blocky labels, the low-resolution logits confident about a noisy copy of them
        coarse = torch.randint(0, c + 1 if with_ignore else c, (n, h, w), generator=g)
        coarse[coarse >= c // 2 + 2] = c if with_ignore else 0
        y = F.interpolate(coarse[:, None].float(), size=(H, W), mode="nearest")[:, 0].long()
        noisy = coarse.clone()
        flips = torch.rand((n, h, w), generator=g) < 0.10
        noisy[flips] = torch.randint(0, c, (int(flips.sum()),), generator=g)
        low = 6.0 * F.one_hot(noisy.clamp(max=c - 1), c).permute(0, 3, 1, 2).float() + torch.randn((n, c, h, w), generator=g)
    # NCHW-contiguous like a convolution's output (models/OCR.py:125): the sum above comes out channels-last, and ATen
    # interpolates channels-last tensors with another kernel whose rounding differs (up to 1e-6 on a logit)
    return low.contiguous(), y


CASES = [
    # name, (n, c, h, w, H, W), dist, experiment-like options
    ("ocr_stride8_c25", (2, 25, 34, 60, 272, 480), "d1", dict(ignore=True)),
    ("deeplab_stride4_c17", (2, 17, 68, 120, 272, 480), "d1", dict(ignore=True)),
    ("c8_no_ignore", (3, 8, 17, 30, 136, 224), "d1", dict(ignore=False)),
    ("per_image_c25", (3, 25, 20, 28, 160, 224), "d1", dict(ignore=True, per_image=True)),
    ("trained_like_c25", (2, 25, 34, 60, 272, 480), "d2", dict(ignore=True)),
    ("trained_like_c17_per_image", (2, 17, 34, 60, 270, 480), "d2", dict(ignore=True, per_image=True)),
    ("filter_ignore_c25", (2, 25, 17, 30, 136, 256), "d1", dict(ignore=True, classes_to_ignore=25)),
    ("all_classes_c25", (1, 25, 17, 30, 136, 256), "d2", dict(ignore=True, classes="all")),
    ("odd_sizes_c8", (2, 8, 9, 11, 75, 96), "d1", dict(ignore=False)),
    ("single_source_row", (1, 8, 1, 10, 16, 64), "d1", dict(ignore=False)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_equals_interpolate_then_loss(b200, case):
    from miccai2021_cataract_semantic_segmentation_b200 import upsampled
    name, (n, c, h, w, H, W), dist, opt = case
    low, y = _inputs(n, c, h, w, H, W, 99 + n * c + h, dist, opt.get("ignore", False))
    yd = y.cuda()
    kw = dict(per_image=opt.get("per_image", False), classes_to_ignore=opt.get("classes_to_ignore"),
              keep_absent=1 if opt.get("classes") == "all" else 0)
    before = dict(upsampled.FALLBACK_COUNTS)
    # fused
    cm_f = torch.zeros((c, c), dtype=torch.int64, device="cuda")
    st_f = torch.zeros(1, dtype=torch.int32, device="cuda")
    lf = low.cuda().requires_grad_(True)
    loss_f = b200.lovasz_softmax_upsampled(lf, yd, confusion=cm_f, confusion_drop_label=c if opt.get("ignore") else None,
                                           status=st_f, **kw)
    (loss_f * 1.7).backward()
    assert upsampled.FALLBACK_COUNTS == before, "the fused kernels must have taken this shape"
    # F.interpolate + the full-resolution kernels
    cm_u = torch.zeros((c, c), dtype=torch.int64, device="cuda")
    st_u = torch.zeros(1, dtype=torch.int32, device="cuda")
    lu = low.cuda().requires_grad_(True)
    loss_u = b200.lovasz_softmax(_up(lu, (H, W)), yd, confusion=cm_u, confusion_drop_label=c if opt.get("ignore") else None,
                                 status=st_u, **kw)
    (loss_u * 1.7).backward()
    assert torch.equal(cm_f, cm_u), "confusion matrix of the fused path differs"
    assert int(st_f) == int(st_u)
    assert rel_err(float(loss_f), float(loss_u)) <= 1e-6
    gmax = float(lu.grad.abs().max())
    assert gmax > 0
    assert grad_err(lf.grad.cpu().numpy(), lu.grad.cpu().numpy()) <= 1e-5
    assert np.allclose(lf.grad.cpu().numpy(), lu.grad.cpu().numpy(), rtol=1e-4, atol=2e-6 * gmax)


@pytest.mark.parametrize("case", CASES[:6], ids=[c[0] for c in CASES[:6]])
def test_fused_matches_oracle(b200, case):
    """against oracle/port.py (the reference's algorithm, stable sort) applied to the ATen-upsampled logits on the device"""
    from oracle import port
    name, (n, c, h, w, H, W), dist, opt = case
    low, y = _inputs(n, c, h, w, H, W, 7 + n * c + h, dist, opt.get("ignore", False))
    yd = y.cuda()
    lf = low.cuda().requires_grad_(True)
    loss_f = b200.lovasz_softmax_upsampled(lf, yd, per_image=opt.get("per_image", False))
    loss_f.backward()
    lr = low.cuda().requires_grad_(True)
    ref = port.lovasz_softmax(_up(lr, (H, W)), yd, {8: 1, 17: 2, 25: 3}[c], per_image=opt.get("per_image", False))
    ref.backward()
    assert rel_err(float(loss_f), float(ref)) <= 1e-5
    assert grad_err(lf.grad.cpu().numpy(), lr.grad.cpu().numpy()) <= 1e-5


def test_fused_cross_entropy_pair(b200):
    """(Lovasz, cross entropy) of the upsampled logits in one pass == the two torch-side evaluations (LossWrapper.py:17-24)"""
    n, c, h, w, H, W = 2, 25, 34, 60, 272, 480
    low, y = _inputs(n, c, h, w, H, W, 5, "d2", True)
    yd = y.cuda()
    lf = low.cuda().requires_grad_(True)
    lov, ce = b200.lovasz_softmax_upsampled(lf, yd, ce_ignore_index=c)
    (0.8 * lov + 1.3 * ce).backward()
    lu = low.cuda().requires_grad_(True)
    full = _up(lu, (H, W))
    lov_u = b200.lovasz_softmax(full, yd)
    ce_u = F.cross_entropy(full, yd, ignore_index=c)
    (0.8 * lov_u + 1.3 * ce_u).backward()
    assert rel_err(float(lov), float(lov_u)) <= 1e-6
    assert rel_err(float(ce), float(ce_u)) <= 1e-5
    assert grad_err(lf.grad.cpu().numpy(), lu.grad.cpu().numpy()) <= 1e-5


def test_module_and_label_dtypes(b200):
    n, c, h, w, H, W = 2, 17, 17, 30, 136, 224
    low, y = _inputs(n, c, h, w, H, W, 21, "d1", True)
    mod = b200.LovaszSoftmaxUpsampled({"experiment": 2})
    ref = b200.LovaszSoftmax({"experiment": 2})
    want = float(ref(_up(low.cuda(), (H, W)), y.cuda()))
    for dt in (torch.uint8, torch.int32, torch.int64):
        got = float(mod(low.cuda(), y.to(dt).cuda()))
        assert rel_err(got, want) <= 1e-6, dt


def test_full_size_training_shape(b200):
    """BASELINE configs[2] shape reached through stride-8 logits: 8 x 25 x 68 x 120 -> 544 x 960"""
    n, c, h, w, H, W = 8, 25, 68, 120, 544, 960
    g = torch.Generator(device="cuda").manual_seed(3)
    low = (torch.randn((n, c, h, w), generator=g, device="cuda") * 2).requires_grad_(True)
    y = torch.randint(0, c + 1, (n, H, W), generator=g, device="cuda")
    cm_f = torch.zeros((c, c), dtype=torch.int64, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    loss = b200.lovasz_softmax_upsampled(low, y, confusion=cm_f, confusion_drop_label=c, status=st)
    loss.backward()
    low2 = low.detach().clone().requires_grad_(True)
    full = _up(low2, (H, W))
    loss2 = b200.lovasz_softmax(full, y)
    loss2.backward()
    assert rel_err(float(loss), float(loss2)) <= 1e-6
    assert grad_err(low.grad.cpu().numpy(), low2.grad.cpu().numpy()) <= 1e-5
    ref_cm = torch.zeros((c, c), dtype=torch.int64, device="cuda")
    keep = y != c
    idx = torch.argmax(full, 1)[keep] * c + y[keep]
    ref_cm += torch.bincount(idx, minlength=c * c).view(c, c)
    assert torch.equal(cm_f, ref_cm)


@pytest.mark.parametrize("shape", [(2, 25, 68, 120, 544, 960), (2, 17, 34, 60, 136, 224), (3, 8, 9, 11, 75, 96), (1, 12, 5, 7, 33, 64),
                                   (1, 25, 1, 1, 8, 32), (2, 25, 136, 240, 272, 480), (1, 3, 40, 50, 30, 32)])
@pytest.mark.parametrize("ldt", [torch.int64, torch.int32, torch.uint8])
def test_confusion_matrix_from_low_resolution_logits(b200, shape, ldt):
    """SegmentationMeter.update_upsampled == update(F.interpolate(...)) bit for bit (torch_utils.py:221-241 after OCR.py:126),
    any class count and scale, near-ties from exactly equal low-resolution logits, the ignore label dropped, NaN pixels."""
    n, c, h, w, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(5 + c + h)
    low = torch.randn((n, c, h, w), generator=g, device="cuda")
    low[:, : c // 2] = (low[:, : c // 2] * 2).round() / 2               # plateaus: exact ties after interpolation
    if h > 2 and w > 2:
        low[0, 1, 1, 1] = float("nan")
    exp = {8: 1, 17: 2, 25: 3}.get(c, 1)
    y = torch.randint(0, c + (1 if c in (17, 25) else 0), (n, H, W), generator=g, device="cuda").to(ldt)
    m1 = b200.SegmentationMeter(exp, c)
    m2 = b200.SegmentationMeter(exp, c)
    m1.update_upsampled(low, y)
    m1.update_upsampled(low, y)                                          # accumulates
    full = _up(low, (H, W))
    m2.update(full, y)
    m2.update(full, y)
    assert torch.equal(m1.cm, m2.cm) and int(m1.cm.sum()) > 0
    pred = full.argmax(1)
    keep = y.long() < c
    ref = 2 * torch.bincount((pred[keep] * c + y.long()[keep]), minlength=c * c).view(c, c)
    assert torch.equal(m1.cm, ref)
    m1.check()


def test_confusion_matrix_upsampled_falls_back_for_odd_widths(b200):
    from miccai2021_cataract_semantic_segmentation_b200 import metrics
    low = torch.randn((1, 8, 6, 7), device="cuda")
    y = torch.randint(0, 8, (1, 24, 40), device="cuda")
    before = metrics.UPSAMPLED_FALLBACK_COUNTS["interpolate_torch"]
    m1, m2 = b200.SegmentationMeter(1, 8), b200.SegmentationMeter(1, 8)
    m1.update_upsampled(low, y)
    m2.update(_up(low, (24, 40)), y)
    assert metrics.UPSAMPLED_FALLBACK_COUNTS["interpolate_torch"] == before + 1
    assert torch.equal(m1.cm, m2.cm)
    y[0, 0, 0] = 9                                                        # out-of-range label: status word, raised by check()
    m3 = b200.SegmentationMeter(1, 8)
    m3.update_upsampled(torch.randn((1, 8, 6, 8), device="cuda"), torch.full((1, 24, 64), 9, device="cuda"))
    with pytest.raises(RuntimeError):
        m3.check()


def test_unsupported_shapes_fall_back_and_count(b200):
    from miccai2021_cataract_semantic_segmentation_b200 import upsampled
    n, c, h, w, H, W = 1, 5, 8, 8, 32, 40            # C = 5 and W % 32 != 0: outside the fused kernels
    low, y = _inputs(n, c, h, w, H, W, 2, "d1", False)
    before = upsampled.FALLBACK_COUNTS["interpolate_torch"]
    with pytest.warns(UserWarning) if before == 0 else _nullcontext():
        got = b200.lovasz_softmax_upsampled(low.cuda(), y.cuda())
    assert upsampled.FALLBACK_COUNTS["interpolate_torch"] == before + 1
    want = b200.lovasz_softmax(_up(low.cuda(), (H, W)), y.cuda())
    assert float(got) == float(want)


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
