"""The raw C-ABI stubs printed in INTEGRATION.md section 2 are executed as written (only the library path is
substituted) and compared with the package's own wrappers."""
import os
import re

import pytest
import torch

from conftest import ROOT


def _section2_code():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 2."):text.index("## 3.")]
    return "\n".join(re.findall(r"```python\n(.*?)```", sec, flags=re.S))


def test_section2_snippets_compile():
    code = _section2_code()
    assert "b200seg_lovasz_forward" in code and "b200seg_ohem_ce_forward" in code and "b200seg_sliding_miou" in code
    compile(code, "INTEGRATION.md#2", "exec")


@pytest.mark.gpu
def test_section2_snippets_run_and_agree_with_the_package():
    import miccai2021_cataract_semantic_segmentation_b200 as b200
    from miccai2021_cataract_semantic_segmentation_b200 import _native
    code = _section2_code().replace('ctypes.CDLL("libb200seg.so")', f'ctypes.CDLL("{_native.lib_path()}")')
    ns = {}
    exec(compile(code, "INTEGRATION.md#2", "exec"), ns)
    g = torch.Generator().manual_seed(1)
    x = torch.randn((2, 25, 32, 48), generator=g).cuda()
    y = torch.randint(0, 26, (2, 32, 48), generator=g).cuda()
    loss, _ws = ns["lovasz_forward"](x, y)
    assert float(loss) == float(b200.LovaszSoftmax({"experiment": 3})(x, y))
    cm = torch.zeros((25, 25), dtype=torch.int64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ns["confusion_matrix"](x, y, cm, status)
    assert torch.equal(cm, b200.t_get_confusion_matrix(x, y))
    lo, _ws2 = ns["ohem_forward"](x, y, 0.7, 500, 25)
    assert float(lo) == float(b200.OhemCrossEntropy({"experiment": 3, "min_kept": 500})(x, y))
    yy = y.clamp(max=24)
    assert torch.equal(ns["sliding_miou_windows"](x, yy), b200.sliding_miou(x, yy, 7, 4, original_size=False))
