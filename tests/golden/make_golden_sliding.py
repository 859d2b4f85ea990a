#!/usr/bin/env python
"""Generate tests/golden/sliding.npz by running the UNMODIFIED reference's sliding_miou
(utils/torch_utils.py:189-218) on seeded inputs.  Build container only (needs /root/reference):

    python tests/golden/make_golden_sliding.py

Same import shim as make_golden.py (matplotlib is absent here).  Inputs are not stored: tests rebuild them from
the recorded seeds with `sliding_inputs` below."""
import json
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("B200SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, seed, n, c, h, w, kernel, stride, style
    ("d_k7s4_c8", 11, 2, 8, 37, 53, 7, 4, "iid"),
    ("d_k7s4_c17", 12, 1, 17, 64, 96, 7, 4, "blocky"),
    ("d_k7s4_c25", 13, 2, 25, 48, 80, 7, 4, "blocky"),
    ("d_k3s1_c5", 14, 1, 5, 9, 11, 3, 1, "iid"),
    ("d_k5s5_c25", 15, 1, 25, 33, 47, 5, 5, "agree"),
    ("d_k9s2_c17", 16, 3, 17, 20, 24, 9, 2, "blocky"),
    ("d_k1s1_c8", 17, 1, 8, 6, 7, 1, 1, "iid"),
]


def sliding_inputs(seed, n, c, h, w, style):
    import torch
    g = torch.Generator().manual_seed(seed)
    if style == "iid":
        x = torch.randn((n, c, h, w), generator=g)
        y = torch.randint(0, c, (n, h, w), generator=g)
    else:
        coarse = torch.randint(0, c, (n, (h + 5) // 6, (w + 5) // 6), generator=g)
        y = coarse.repeat_interleave(6, 1).repeat_interleave(6, 2)[:, :h, :w].contiguous()
        noise = 2.5 if style == "blocky" else 0.0
        x = 4.0 * torch.nn.functional.one_hot(y, c).permute(0, 3, 1, 2).float() + noise * torch.randn((n, c, h, w), generator=g)
    return x, y


if __name__ == "__main__":
    for _m in ("matplotlib", "matplotlib.colors", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        sys.modules.setdefault(_m, MagicMock())
    sys.path.insert(0, REF)
    warnings.simplefilter("ignore")
    import torch
    import utils  # noqa: F401  (reference)
    ref = sys.modules["utils.torch_utils"].sliding_miou
    out, manifest = {}, {"torch": torch.__version__, "cases": []}
    for name, seed, n, c, h, w, k, s, style in CASES:
        x, y = sliding_inputs(seed, n, c, h, w, style)
        out[name + "/windows"] = ref(x, y, k, s, original_size=False).numpy()
        out[name + "/full"] = ref(x, y, k, s, original_size=True).numpy()
        manifest["cases"].append(dict(name=name, seed=seed, n=n, c=c, h=h, w=w, kernel=k, stride=s, style=style))
    np.savez_compressed(os.path.join(HERE, "sliding.npz"), **out)
    with open(os.path.join(HERE, "sliding_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(CASES), "cases;", os.path.getsize(os.path.join(HERE, "sliding.npz")), "bytes")
