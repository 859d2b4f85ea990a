"""bench.py prints ONE JSON line with the keys the driver reads (task statement, section 4): the reference arm on the
CPU here, the GPU arm (value, e2e, roofline, cpu_baseline, clocks, gpu_launches) on the B200 box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--height", "48", "--width", "64", "--batch", "2"], 300)
    assert BASE_KEYS <= set(out) and out["impl"] == "reference"
    assert out["unit"] == "Mpixel/s" and out["higher_is_better"] is True and out["vs_baseline"] is None
    assert out["value"] > 0 and out["e2e"]["value"] == out["value"]
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["e2e"]["d2h_bytes_per_step"] == 0
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1 and "workload" in out["config"]


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    out = _run(["--steps", "3", "--warmup", "3", "--batch", "2"], 600)
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "roofline_step", "stage_ms"} <= set(out)
    assert out["n_gpus"] == 1 and out["dtype"] == "f32" and out["data"] == "synthetic" and out["scaling"] == "weak"
    assert out["value"] > 0 and out["gpu_launches"] == 11 * out["steps"] and out["sort_path"] == "hybrid"
    e2e = out["e2e"]
    assert 0 < e2e["value"] < out["value"]                       # host buffers: PCIe-bound
    assert e2e["h2d_bytes_per_step"] == 2 * 540 * 960 * (25 * 4 + 8) and e2e["d2h_bytes_per_step"] > 0
    roof = out["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert roof["frac"] > 0 and roof["kernel"] in out["stage_ms"]
    cpu = out["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["value"] > 0 and cpu["cores"] >= 1 and cpu["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(out["clocks"])
    names = [cfg["name"] for cfg in out["configs"]]
    assert any("d2_blocky_c25_flat" in n for n in names) and any("configs[1]" in n for n in names) and any("configs[4]" in n for n in names)
    assert all(cfg["value"] > 0 for cfg in out["configs"])


@pytest.mark.gpu
def test_gpu_sweep_mode_prints_the_contract_line():
    out = _run(["--sweep-frames", "48", "--sweep-batch", "16"], 600)
    assert out["value"] > 0 and out["n_gpus"] == 1 and "configs[4]" in out["config"]["workload"] and out["roofline"]["frac"] > 0
