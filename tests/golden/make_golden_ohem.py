#!/usr/bin/env python
"""Generate tests/golden/ohem.npz by running the UNMODIFIED reference's OhemCrossEntropy
(losses/OhemCrossEntropy.py) forward + autograd backward on seeded inputs.  Build container only:

    python tests/golden/make_golden_ohem.py

Same import shim as make_golden.py (matplotlib is absent here).  Inputs are rebuilt by tests from the recorded
seeds with `ohem_inputs` below; losses and gradients are stored."""
import json
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("B200SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, seed, n, c, h, w, config, style
    ("o_exp1_default", 21, 2, 8, 24, 32, {"experiment": 1}, "iid"),                       # min_kept > n: all below max kept
    ("o_exp2_kept50", 22, 2, 17, 24, 32, {"experiment": 2, "min_kept": 50}, "confident"),  # thresh 0.7 decides
    ("o_exp3_kept400", 23, 1, 25, 32, 48, {"experiment": 3, "min_kept": 400, "thresh": 0.2}, "confident"),
    ("o_exp3_rank", 24, 2, 25, 20, 28, {"experiment": 3, "min_kept": 300, "thresh": 0.01}, "iid"),   # order statistic decides
    ("o_noexp", 25, 1, 5, 9, 11, {"min_kept": 10, "thresh": 0.5}, "confident"),
    ("o_exp2_ignore_heavy", 26, 2, 17, 16, 16, {"experiment": 2, "min_kept": 20, "thresh": 0.3}, "mostly_ignored"),
]


def ohem_inputs(seed, n, c, h, w, config, style):
    import torch
    g = torch.Generator().manual_seed(seed)
    with_ignore = config.get("experiment") in (2, 3)
    hi = c + 1 if with_ignore else c
    y = torch.randint(0, hi, (n, h, w), generator=g)
    if style == "mostly_ignored":
        y[torch.rand((n, h, w), generator=g) < 0.8] = c
    x = torch.randn((n, c, h, w), generator=g)
    if style != "iid":
        x = x + 4.0 * torch.nn.functional.one_hot(y.clamp(max=c - 1), c).permute(0, 3, 1, 2).float() \
            * (torch.rand((n, 1, h, w), generator=g) < 0.7).float()
    return x, y


if __name__ == "__main__":
    for _m in ("matplotlib", "matplotlib.colors", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        sys.modules.setdefault(_m, MagicMock())
    sys.path.insert(0, REF)
    warnings.simplefilter("ignore")
    import torch
    import losses  # noqa: F401  (reference)
    ref_cls = sys.modules["losses.OhemCrossEntropy"].OhemCrossEntropy
    out, manifest = {}, {"torch": torch.__version__, "cases": []}
    for name, seed, n, c, h, w, config, style in CASES:
        x, y = ohem_inputs(seed, n, c, h, w, config, style)
        x.requires_grad_(True)
        mod = ref_cls(config)
        loss = mod(x, y)
        loss.backward()
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/grad"] = x.grad.numpy()
        manifest["cases"].append(dict(name=name, seed=seed, n=n, c=c, h=h, w=w, config=config, style=style,
                                      ignore_label=mod.ignore_label, thresh=mod.thresh, min_kept=mod.min_kept,
                                      kept_pixels=int((x.grad.abs().sum(1) > 0).sum())))
    np.savez_compressed(os.path.join(HERE, "ohem.npz"), **out)
    with open(os.path.join(HERE, "ohem_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(CASES), "cases;", os.path.getsize(os.path.join(HERE, "ohem.npz")), "bytes")
    for cse in manifest["cases"]:
        print(cse["name"], "kept", cse["kept_pixels"], "of", cse["n"] * cse["h"] * cse["w"])
