import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_keys(name):
    with open(os.path.join(GOLDEN, f"keys_{name}.json")) as f:
        return [(k, tuple(s)) for k, s in json.load(f)]


def golden_npz(fname):
    return np.load(os.path.join(GOLDEN, fname))


def golden_json(fname):
    with open(os.path.join(GOLDEN, fname)) as f:
        return json.load(f)
