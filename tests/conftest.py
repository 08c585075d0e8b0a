import glob
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _gpu_usable():
    """A CUDA device and the built CUDA library: without both, ``-m gpu`` tests are skipped, not failed."""
    try:
        import torch
        from gym_treasure_game_b200 import _lib
        return torch.cuda.is_available() and os.path.exists(_lib.SO)
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _gpu_usable():
        skip = pytest.mark.skip(reason="no CUDA device (or libtreasure_b200.so not built)")
        for it in gpu_items:
            it.add_marker(skip)


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))


def load_golden(path):
    with gzip.open(path, "rt") as f:
        return json.load(f)


def golden_level(rec):
    import py_oracle as po
    lv = rec["level"]
    return po.LevelText.from_strings(lv["domain"], lv["objects"], lv["interactions"])


def norm_snap(s):
    """Golden JSON turns tuples into lists; normalise for comparison."""
    s = dict(s)
    s["items"] = [tuple(i) for i in s["items"]]
    return s


@pytest.fixture(scope="session")
def goldens():
    return [load_golden(p) for p in golden_files()]
