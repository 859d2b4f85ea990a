"""BASELINE configs[3] parity half (``-m gpu``): an OCRNet-R50 (random init) step through the UNMODIFIED reference's own
LossWrapper / TwoScaleLoss -- constructed by the reference's name lookup -- before and after install().  Needs the
reference packages (baseline/_ref, staged by tools/stage_reference.sh where /root/reference exists); skipped otherwise."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _have_reference():
    return any(c and os.path.isdir(os.path.join(c, "losses")) and os.path.isdir(os.path.join(c, "models"))
               for c in (os.environ.get("B200SEG_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")))


@pytest.mark.parametrize("kind", ["wrapper", "twoscale"])
def test_ocrnet_step_through_the_reference_compositors(kind):
    if not _have_reference():
        pytest.skip("no reference tree on this box")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_step_ocrnet.py"), "--batch", "2", "--height", "192",
                          "--width", "320", "--steps", "1", "--loss", kind], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert "unavailable" not in out
    assert out["loss_rel_err"] <= 1e-5
    assert out["dlogits_rel_err"] <= 1e-5                      # north_star gate, at the loss boundary
    # parameter gradients: the same dlogits pushed through ~100 layers whose backward uses atomics; held to the noise floor
    # the reference shows against itself (same loss evaluated twice)
    assert out["param_grad_rel_err"] <= max(1e-4, 4.0 * out["param_grad_noise_floor"])
    assert out["confusion_matrix_equal"] and out["iou_max_abs_err"] <= 1e-6
    assert "managers.OCRNet_Manager" in out["config"]["rebound_modules"]
