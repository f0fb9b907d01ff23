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


def same_up_to_score_ties(a, b):
    """NMS outputs are equal, except that rows with bit-identical confidence may appear in either order: above max_nms
    = 30000 candidates the reference orders them with an UNSTABLE `argsort(descending=True)` (nms.py:90-91), so the rank
    of tied rows is unspecified there (the CUDA path and the oracle use the stable order)."""
    if a.shape != b.shape or not np.array_equal(a[:, 4], b[:, 4]):
        return False
    i = 0
    while i < a.shape[0]:
        j = i + 1
        while j < a.shape[0] and a[j, 4] == a[i, 4]:
            j += 1
        if sorted(map(tuple, a[i:j].tolist())) != sorted(map(tuple, b[i:j].tolist())):
            return False
        i = j
    return True
