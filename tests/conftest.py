import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """tests/golden/cases.npz + manifest.json: vectors produced by the unmodified reference (make_golden.py)."""

    def __init__(self):
        self.arrays = np.load(os.path.join(GOLDEN_DIR, "cases.npz"))
        with open(os.path.join(GOLDEN_DIR, "manifest.json")) as f:
            self.manifest = json.load(f)

    def inputs(self, entry):
        i = entry["input"]
        return self.arrays[f"input/{i}/logits"], self.arrays[f"input/{i}/target"].astype(np.int64)

    def get(self, kind, name, field):
        return self.arrays[f"{kind}/{name}/{field}"]

    def has(self, kind, name, field):
        return f"{kind}/{name}/{field}" in self.arrays.files


_GOLDEN = None


def golden():
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = Golden()
    return _GOLDEN


def lovasz_case_ids():
    return [e["name"] for e in golden().manifest["lovasz"]]


def confmat_case_ids():
    return [e["name"] for e in golden().manifest["confmat"]]


def lovasz_entry(name):
    return next(e for e in golden().manifest["lovasz"] if e["name"] == name)


def confmat_entry(name):
    return next(e for e in golden().manifest["confmat"] if e["name"] == name)


def rel_err(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-30)


def grad_err(g, gref):
    """norm-relative max error, SURVEY.md §8(d) parity gate."""
    den = max(float(np.abs(gref).max()), 1e-30)
    return float(np.abs(np.asarray(g, dtype=np.float64) - np.asarray(gref, dtype=np.float64)).max()) / den


@pytest.fixture(scope="session")
def gold():
    return golden()
