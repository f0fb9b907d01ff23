"""GPU parity of the batched evaluation post-processing (yv6_eval_boxes + evalpost.convert_to_coco_format) against the
reference's rows (tests/golden/evalpost.json) and the oracle: identical json rows (decimal strings included)."""
import pytest
import torch

from conftest import golden_json
from oracle import evalpost as oe
from test_oracle_evalpost import IDS

pytestmark = pytest.mark.gpu


def pack(outs, max_det, dev):
    out = torch.zeros(len(outs), max_det, 6)
    count = torch.zeros(len(outs), dtype=torch.int32)
    for i, o in enumerate(outs):
        out[i, :o.shape[0]] = o
        count[i] = o.shape[0]
    return out.to(dev), count.to(dev)


def test_rows_equal_reference_and_oracle():
    from yolov6_b200.evalpost import convert_to_coco_format, to_end2end
    g = golden_json("evalpost.json")
    dev = torch.device("cuda:0")
    for seed in (0, 1, 5):
        outs, paths, shapes = oe.synthetic_batch(seed=seed)
        out, count = pack(outs, 64, dev)
        rows = convert_to_coco_format(out, count, paths, shapes, IDS)
        assert rows == oe.convert_to_coco_format(outs, paths, shapes, IDS)
        if f"seed{seed}" in g:
            assert rows == g[f"seed{seed}"]
    n, boxes, scores, classes = to_end2end(out, count)
    assert n.shape == (4, 1) and boxes.shape == (4, 64, 4) and scores.shape == (4, 64) and classes.dtype == torch.int32
