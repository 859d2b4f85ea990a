#!/usr/bin/env python
"""Generate tests/golden/cases.npz + manifest.json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported from where it lies (nothing is copied into this repo).  Two
process-local shims, both described in SURVEY.md §8(c):
  * matplotlib is absent here, so utils/utils.py:10-13's imports are satisfied with MagicMock;
  * the ``torch`` name inside losses/LovaszSoftmax.py is wrapped so torch.sort runs with
    stable=True (canonical tie order: descending error, ascending flattened pixel index).
Everything else (softmax, cumsum, dot, autograd, one_hot GEMM, IoU formulas) is the reference's code.
"""
import json
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("B200SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for _m in ("matplotlib", "matplotlib.colors", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
    sys.modules.setdefault(_m, MagicMock())
sys.path.insert(0, REF)
warnings.simplefilter("ignore")

import torch  # noqa: E402
import losses  # noqa: E402  (reference)
import utils  # noqa: E402   (reference)

_LS = sys.modules["losses.LovaszSoftmax"]


class _StableTorch:
    def __getattr__(self, k):
        return getattr(torch, k)

    @staticmethod
    def sort(input, dim=-1, descending=False):
        return torch.sort(input, dim=dim, descending=descending, stable=True)


_LS.torch = _StableTorch()

C_OF = {1: 8, 2: 17, 3: 25}
OUT = {}
MANIFEST = {"torch": torch.__version__, "lovasz": [], "confmat": [], "class_info": {}}


def _labels(gen, n, h, w, c, with_ignore):
    hi = c + 1 if with_ignore else c
    return torch.randint(0, hi, (n, h, w), generator=gen)


_INPUTS = []


def _store_input(x, y):
    """Inputs are shared between cases; store each distinct (logits, labels) pair once (labels as int16)."""
    xn = x.detach().numpy().astype(np.float32)
    yn = y.numpy().astype(np.int16)
    for i, (a, b) in enumerate(_INPUTS):
        if a.shape == xn.shape and np.array_equal(a, xn) and np.array_equal(b, yn):
            return i
    _INPUTS.append((xn, yn))
    i = len(_INPUTS) - 1
    OUT[f"input/{i}/logits"] = xn
    OUT[f"input/{i}/target"] = yn
    return i


def add_lovasz(name, logits, target, cfg):
    """cfg: the dict handed to reference LovaszSoftmax (experiment, per_image, classes_to_ignore, classes_to_consider)."""
    x = logits.clone().requires_grad_(True)
    loss = losses.LovaszSoftmax(dict(cfg))(x, target)
    if torch.is_tensor(loss) and loss.requires_grad and loss.numel() == 1:
        loss.backward()
        g = x.grad.detach().numpy()
        lv = np.float32(loss.item())
    else:   # python int 0 (nothing kept) or an empty tensor
        g = np.zeros_like(logits.numpy())
        lv = np.float32(float(loss) if not torch.is_tensor(loss) else float(loss.sum()))
    inp = _store_input(logits, target)
    OUT[f"lovasz/{name}/loss"] = lv
    OUT[f"lovasz/{name}/grad"] = g.astype(np.float32)
    jcfg = {k: (v if not isinstance(v, (np.integer,)) else int(v)) for k, v in cfg.items()}
    MANIFEST["lovasz"].append({"name": name, "config": jcfg, "input": inp})
    print(f"lovasz  {name:34s} loss={lv:.10f} |g|max={np.abs(g).max():.3e}")


def add_confmat(name, prediction, target, experiment, existing=None, no_ignore_class=True, metrics=True):
    cm = utils.t_get_confusion_matrix(prediction, target, existing, no_ignore_class)
    inp = _store_input(prediction, target)
    if existing is not None:
        OUT[f"confmat/{name}/existing"] = existing.numpy()
    OUT[f"confmat/{name}/cm"] = cm.numpy()
    assert cm.dtype == torch.int32
    pa, pac = utils.t_get_pixel_accuracy(cm)
    OUT[f"confmat/{name}/pixel_accuracy"] = np.array([pa.item(), pac.item()], dtype=np.float32)
    MANIFEST["confmat"].append({"name": name, "experiment": experiment, "no_ignore_class": no_ignore_class,
                                "has_existing": existing is not None, "metrics": metrics, "input": inp,
                                "target_dtype": str(target.dtype).replace("torch.", "")})
    if not metrics:     # class count differs from the experiment's table (known-answer vector)
        print(f"confmat {name:34s} sum={int(cm.sum())}")
        return
    OUT[f"confmat/{name}/miou"] = np.float32(utils.t_get_mean_iou(cm, experiment).item())
    cat = utils.t_get_mean_iou(cm, experiment, True, rare=True)
    OUT[f"confmat/{name}/miou_categories_rare"] = np.array([v.item() for v in cat], dtype=np.float32)
    cat3 = utils.t_get_mean_iou(cm, experiment, True)
    OUT[f"confmat/{name}/miou_categories"] = np.array([v.item() for v in cat3], dtype=np.float32)
    vec = utils.t_get_mean_iou(cm, experiment, True, calculate_mean=False, rare=True)
    for tag, v in zip(("all", "instruments", "anatomies", "rare"), vec):
        OUT[f"confmat/{name}/iou_vec_{tag}"] = v.numpy().astype(np.float32)
    OUT[f"confmat/{name}/norm_row"] = utils.t_normalise_confusion_matrix(cm, "row").numpy()
    OUT[f"confmat/{name}/norm_col"] = utils.t_normalise_confusion_matrix(cm, "col").numpy()
    c = prediction.shape[1]
    OUT[f"confmat/{name}/single_class_iou"] = np.array(
        [float(utils.t_get_single_class_iou(cm, experiment, k)) for k in range(c)], dtype=np.float32)
    if experiment == 1:   # numpy twins only work without an ignore label (utils/metrics.py:5-25)
        ncm = utils.get_confusion_matrix(prediction, target.long())
        OUT[f"confmat/{name}/np_cm"] = ncm
        OUT[f"confmat/{name}/np_miou_categories"] = np.array(utils.get_mean_iou(ncm, 1, categories=True), np.float64)
        OUT[f"confmat/{name}/np_pixel_accuracy"] = np.array(utils.get_pixel_accuracy(ncm.copy()), np.float64)
        OUT[f"confmat/{name}/np_miou"] = np.array(utils.get_mean_iou(ncm, 1), np.float64)
        OUT[f"confmat/{name}/np_norm_row"] = np.asarray(utils.normalise_confusion_matrix(ncm.copy(), "row"), np.float64)
        OUT[f"confmat/{name}/np_norm_col"] = np.asarray(utils.normalise_confusion_matrix(ncm.copy(), "col"), np.float64)
        OUT[f"confmat/{name}/np_single_class_iou"] = np.array(
            [utils.get_single_class_iou(ncm, 1, k) for k in range(prediction.shape[1])], np.float64)
    print(f"confmat {name:34s} sum={int(cm.sum())} miou={OUT[f'confmat/{name}/miou']:.8f}")


def real_mask_crop(experiment, n, h, w, seed):
    """Blocky labels from the reference's relabelled/*.png (raw ids 0..35) remapped like the dataset does
    (utils/utils.py:23-47, to_network=True), subsampled to h x w."""
    import cv2
    files = sorted(os.listdir(os.path.join(REF, "relabelled")))
    rng = np.random.RandomState(seed)
    out = []
    for f in rng.choice(files, n, replace=False):
        m = cv2.imread(os.path.join(REF, "relabelled", f), 0)
        m = utils.remap_mask(m, utils.CLASS_INFO[experiment][0], to_network=True)
        ys = np.linspace(0, m.shape[0] - 1, h).astype(int)
        xs = np.linspace(0, m.shape[1] - 1, w).astype(int)
        out.append(m[np.ix_(ys, xs)])
    return torch.from_numpy(np.stack(out).astype(np.int64))


def trained_like_logits(labels, c, gen, flip=0.10, gain=6.0):
    n, h, w = labels.shape
    noisy = labels.clone()
    flips = torch.rand((n, h, w), generator=gen) < flip
    noisy[flips] = torch.randint(0, c, (int(flips.sum()),), generator=gen)
    noisy = noisy.clamp(max=c - 1)
    onehot = torch.nn.functional.one_hot(noisy, c).permute(0, 3, 1, 2).float()
    return gain * onehot + torch.randn((n, c, h, w), generator=gen)


def main():
    g = torch.Generator().manual_seed(20211001)

    # ---- known-answer vector of SURVEY.md §8(c) ---------------------------------------------------
    kat_logits = torch.tensor([[[[2, .5, -1], [0, 1, .25]],
                                [[0, 1.5, .5], [0, -2, .25]],
                                [[-1, .5, 3], [0, .5, .25]]]], dtype=torch.float32)
    kat_target = torch.tensor([[[0, 1, 2], [1, 0, 0]]])
    add_lovasz("kat_c3", kat_logits, kat_target, {"experiment": 1})
    add_confmat("kat_c3", kat_logits, kat_target.int(), 1, metrics=False)

    H, W = 15, 28
    for exp in (1, 2, 3):
        c = C_OF[exp]
        x = torch.randn((2, c, H, W), generator=g)
        y = _labels(g, 2, H, W, c, exp != 1)
        add_lovasz(f"d1_exp{exp}_flat", x, y, {"experiment": exp})
        add_lovasz(f"d1_exp{exp}_per_image", x, y, {"experiment": exp, "per_image": True})
        if exp != 1:
            add_lovasz(f"d1_exp{exp}_flat_ignore", x, y, {"experiment": exp, "classes_to_ignore": c})
            add_lovasz(f"d1_exp{exp}_per_image_ignore", x, y,
                       {"experiment": exp, "per_image": True, "classes_to_ignore": c})
        add_confmat(f"d1_exp{exp}", x, y.int(), exp)
        prev = torch.randint(0, 1000, (c, c), generator=g).int()
        add_confmat(f"d1_exp{exp}_accumulate", x, y.int(), exp, existing=prev)

    # ignore value 255 (BASELINE config 1 wording) through classes_to_ignore
    x = torch.randn((2, 8, H, W), generator=g)
    y = _labels(g, 2, H, W, 8, False)
    y[torch.rand((2, H, W), generator=g) < 0.05] = 255
    add_lovasz("d1_255_exp1_flat_ignore255", x, y, {"experiment": 1, "classes_to_ignore": 255})
    add_lovasz("d1_255_exp1_flat_keep255", x, y, {"experiment": 1})   # 255 pixels act as background

    # class selection modes with absent classes
    c = 17
    x = torch.randn((2, c, H, W), generator=g)
    y = _labels(g, 2, H, W, c, True)
    y[y == 3] = 4
    y[y == 11] = 0          # classes 3 and 11 absent
    add_lovasz("modes_exp2_present", x, y, {"experiment": 2})
    add_lovasz("modes_exp2_all", x, y, {"experiment": 2, "classes_to_consider": "all"})
    add_lovasz("modes_exp2_list", x, y, {"experiment": 2, "classes_to_consider": [0, 3, 5, 11, 16]})
    add_lovasz("modes_exp2_all_per_image_ignore", x, y,
               {"experiment": 2, "classes_to_consider": "all", "per_image": True, "classes_to_ignore": 17})
    json_present = json.loads('"present"')       # not interned: `is 'present'` is False at LovaszSoftmax.py:53
    add_lovasz("modes_exp2_json_present", x, y, {"experiment": 2, "classes_to_consider": json_present})
    MANIFEST["lovasz"][-1]["present_only"] = False

    # adversarial ties: logits on a 0.5 grid, +-60 spikes, one class with a single pixel, one image all-ignore
    c = 8
    x = (torch.randn((3, c, H, W), generator=g) * 2).round() / 2
    spikes = torch.rand((3, c, H, W), generator=g)
    x[spikes < 0.02] = 60.0
    x[spikes > 0.98] = -60.0
    y = _labels(g, 3, H, W, c, False)
    y[y == 6] = 5
    y[0, 0, 0] = 6           # class 6: exactly one pixel
    y[y == 2] = 1            # class 2 absent
    add_lovasz("d3_ties_exp1_flat", x, y, {"experiment": 1})
    add_lovasz("d3_ties_exp1_per_image", x, y, {"experiment": 1, "per_image": True})
    add_confmat("d3_ties_exp1", x, y.int(), 1)
    c = 25
    x = (torch.randn((3, c, H, W), generator=g) * 2).round() / 2
    y = _labels(g, 3, H, W, c, True)
    y[1] = c                 # image 1: ignore everywhere
    add_lovasz("d3_ties_exp3_per_image_allignore", x, y, {"experiment": 3, "per_image": True})
    # NB: the same input with classes_to_ignore=25 makes the reference itself raise (an empty [0, C] tensor meets
    # a 0-dim accumulator in mean(), LovaszSoftmax.py:117) -- no golden vector exists for that case.
    add_confmat("d3_ties_exp3", x, y.int(), 3)
    # odd plane size (vector-tail paths), N=1
    x = torch.randn((1, 17, 7, 13), generator=g)
    y = _labels(g, 1, 7, 13, 17, True)
    add_lovasz("odd_exp2_flat", x, y, {"experiment": 2})
    add_confmat("odd_exp2", x, y.to(torch.uint8), 2)
    add_confmat("odd_exp2_no_drop", x, y.clamp(max=16).long(), 2, no_ignore_class=False)

    # trained-like logits over real CaDIS masks (D2)
    for exp in (2, 3):
        c = C_OF[exp]
        y = real_mask_crop(exp, 2, 27, 48, seed=exp)
        x = trained_like_logits(y, c, g)
        add_lovasz(f"d2_exp{exp}_flat", x, y, {"experiment": exp})
        add_lovasz(f"d2_exp{exp}_per_image", x, y, {"experiment": exp, "per_image": True})
        add_confmat(f"d2_exp{exp}", x, y.int(), exp)
    # probabilities instead of logits (Ensemble path, models/Ensemble.py:66)
    add_confmat("d2_exp3_probabilities", torch.softmax(x, 1), y.int(), 3)

    # losses/iou.py:31-35 (function IoU); the module-level name is shadowed by nothing once imported directly
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_iou", os.path.join(REF, "losses", "iou.py"))
    ref_iou = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_iou)
    a = torch.rand((2, 3, 9, 11), generator=g)
    b = (torch.rand((2, 3, 9, 11), generator=g) > 0.5).float()
    OUT["softiou/x"] = a.numpy()
    OUT["softiou/t"] = b.numpy()
    OUT["softiou/out"] = ref_iou.IoU(a, b).numpy()

    for exp in (1, 2, 3):
        MANIFEST["class_info"][str(exp)] = {
            "keys": [int(k) for k in utils.CLASS_INFO[exp][1].keys()],
            "categories": {k: [int(i) for i in v] for k, v in utils.CLASS_INFO[exp][2].items()},
        }

    np.savez_compressed(os.path.join(HERE, "cases.npz"), **OUT)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(MANIFEST, f, indent=1, sort_keys=True)
    sz = os.path.getsize(os.path.join(HERE, "cases.npz"))
    print(f"wrote {len(OUT)} arrays, {sz / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
