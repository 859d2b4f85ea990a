"""Pins the oracle (oracle/port.py, oracle/lovasz_cm_ref.c) to vectors produced by the reference itself."""
import numpy as np
import pytest
import torch

from conftest import (confmat_case_ids, confmat_entry, golden, grad_err, lovasz_case_ids, lovasz_entry, rel_err)
from oracle import cref, port

# North-star tolerances: loss and gradient within 1e-5 relative (gradient: max error / max |g_ref|).
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-5


def _cfg(entry):
    cfg = dict(entry["config"])
    kw = dict(per_image=cfg.get("per_image", False), classes_to_ignore=cfg.get("classes_to_ignore"),
              classes_to_consider=cfg.get("classes_to_consider", "present"))
    if "present_only" in entry:
        kw["present_only"] = entry["present_only"]
    return cfg["experiment"], kw


@pytest.mark.parametrize("name", lovasz_case_ids())
def test_port_matches_reference(name):
    g = golden()
    e = lovasz_entry(name)
    x, y = g.inputs(e)
    exp, kw = _cfg(e)
    loss, grad = port.lovasz_softmax_with_grad(torch.from_numpy(x), torch.from_numpy(y), exp, **kw)
    # same ATen ops in the same order on the same torch build: expect equality, gate at the north-star tolerance
    assert rel_err(loss, g.get("lovasz", name, "loss")) <= LOSS_RTOL
    assert grad_err(grad.numpy(), g.get("lovasz", name, "grad")) <= GRAD_RTOL
    assert float(loss) == float(g.get("lovasz", name, "loss"))


@pytest.mark.parametrize("name", lovasz_case_ids())
def test_c_restatement_matches_reference(name):
    g = golden()
    e = lovasz_entry(name)
    x, y = g.inputs(e)
    _, kw = _cfg(e)
    loss, grad = cref.lovasz(x, y, **kw)
    assert rel_err(loss, g.get("lovasz", name, "loss")) <= LOSS_RTOL
    assert grad_err(grad, g.get("lovasz", name, "grad")) <= GRAD_RTOL


def test_known_answer_vector_digits():
    """loss 0.4214904308 and the recorded gradient digits."""
    g = golden()
    assert abs(float(g.get("lovasz", "kat_c3", "loss")) - 0.4214904308) < 5e-8
    grad = g.get("lovasz", "kat_c3", "grad")
    assert np.allclose(grad[0, 0], [[-0.00878701, 0.02238557, 0.00100839], [0.01851852, -0.06342983, -0.02777778]],
                       atol=5e-8)
    assert np.array_equal(g.get("confmat", "kat_c3", "cm"), [[3, 1, 0], [0, 1, 0], [0, 0, 1]])


@pytest.mark.parametrize("name", confmat_case_ids())
def test_confusion_matrix_and_metrics(name):
    g = golden()
    e = confmat_entry(name)
    x, y = g.inputs(e)
    existing = g.get("confmat", name, "existing") if e["has_existing"] else None
    ref_cm = g.get("confmat", name, "cm")
    tdt = getattr(torch, e["target_dtype"])
    cm = port.confusion_matrix(torch.from_numpy(x), torch.from_numpy(y).to(tdt),
                               None if existing is None else torch.from_numpy(existing), e["no_ignore_class"])
    assert cm.dtype == torch.int32 and np.array_equal(cm.numpy(), ref_cm)
    ccm = cref.confmat(x, y, e["no_ignore_class"], existing)
    assert np.array_equal(ccm, ref_cm.astype(np.int64))
    # invariants of utils/metrics.py:17-21
    if existing is None:
        c = x.shape[1]
        assert int(cm.sum()) == int((y < c).sum())
    pa, pac = port.pixel_accuracy(cm)
    assert np.array_equal(np.array([pa.item(), pac.item()], np.float32), g.get("confmat", name, "pixel_accuracy"))
    if not e["metrics"]:
        return
    exp = e["experiment"]
    assert np.float32(port.mean_iou(cm, exp).item()) == g.get("confmat", name, "miou")
    four = port.mean_iou(cm, exp, True, rare=True)
    assert np.array_equal(np.array([v.item() for v in four], np.float32), g.get("confmat", name, "miou_categories_rare"))
    three = port.mean_iou(cm, exp, True)
    assert np.array_equal(np.array([v.item() for v in three], np.float32), g.get("confmat", name, "miou_categories"))
    vecs = port.mean_iou(cm, exp, True, calculate_mean=False, rare=True)
    for tag, v in zip(("all", "instruments", "anatomies", "rare"), vecs):
        assert np.array_equal(v.numpy(), g.get("confmat", name, f"iou_vec_{tag}"))
    assert np.array_equal(port.normalise_confusion_matrix(cm, "row").numpy(), g.get("confmat", name, "norm_row"))
    assert np.array_equal(port.normalise_confusion_matrix(cm, "col").numpy(), g.get("confmat", name, "norm_col"))
    sc = np.array([float(port.single_class_iou(cm, exp, k)) for k in range(x.shape[1])], np.float32)
    assert np.array_equal(sc, g.get("confmat", name, "single_class_iou"))
    if g.has("confmat", name, "np_cm"):
        ncm = port.np_confusion_matrix(x, y)
        assert np.array_equal(ncm, g.get("confmat", name, "np_cm"))
        assert np.allclose(np.array(port.np_mean_iou(ncm, 1, True)), g.get("confmat", name, "np_miou_categories"),
                           rtol=0, atol=1e-15)
        assert np.allclose(np.array(port.np_pixel_accuracy(ncm.copy())), g.get("confmat", name, "np_pixel_accuracy"),
                           rtol=0, atol=1e-15)


def test_class_tables_match_reference():
    g = golden()
    for exp in (1, 2, 3):
        info = g.manifest["class_info"][str(exp)]
        assert port.class_keys(exp) == info["keys"]
        assert port.CATEGORIES[exp] == info["categories"]


def test_soft_iou():
    g = golden()
    out = port.soft_iou(torch.from_numpy(g.arrays["softiou/x"]), torch.from_numpy(g.arrays["softiou/t"]))
    assert np.array_equal(out.numpy(), g.arrays["softiou/out"])


def test_label_range_errors_mirror_one_hot():
    x = torch.zeros(1, 8, 2, 2)
    with pytest.raises(RuntimeError):
        port.confusion_matrix(x, torch.full((1, 2, 2), 8))
    with pytest.raises(RuntimeError):
        cref.confmat(x.numpy(), np.full((1, 2, 2), 8))
    x17 = torch.zeros(1, 17, 2, 2)
    assert int(port.confusion_matrix(x17, torch.full((1, 2, 2), 17)).sum()) == 0      # ignore label dropped
    with pytest.raises(RuntimeError):
        port.confusion_matrix(x17, torch.full((1, 2, 2), 18))
    with pytest.raises(RuntimeError):
        port.confusion_matrix(x17, torch.full((1, 2, 2), 17), no_ignore_class=False)
