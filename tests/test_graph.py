"""Host-side logic without a GPU: the layer graph reproduces the reference's parameter names/shapes
(checkpoint compatibility, SURVEY.md section 5 'Checkpoint / resume'), and the BN-fold / RepVGG
re-parameterisation (fold.py) reproduces the oracle's train-form network when applied on CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_keys, golden_npz
from oracle import fabricate as fab
from oracle import model as om
from yolov6_b200 import arch, configs
from yolov6_b200.fold import fold_op
from yolov6_b200.model import build_model

MODELS = ["yolov6n", "yolov6s", "yolov6m", "yolov6l6"]


@pytest.mark.parametrize("name", MODELS)
def test_state_dict_matches_reference_layout(name):
    ref = dict(golden_keys(name))
    m = build_model(name, 80, torch.device("cpu"))
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == ref
    m.load_state_dict(fab.fabricate_state_dict(list(ref.items()), 0), strict=True)
    assert m.detect.nc == 80 and m.detect.no == 85 and len(m.detect.stems) == m.detect.nl
    assert torch.equal(m.stride, torch.tensor(configs.CONFIGS[name]["head"]["strides"]))


def test_conv_counts_match_survey():
    # 71 conv launches for YOLOv6-S in deploy form (69 Conv2d + 2 ConvTranspose2d)
    g = arch.build_graph(configs.get_config("yolov6s"), 80)
    assert sum(1 for o in g.ops if o.kind in ("stem", "conv", "pred", "convT")) == 71


def _run_graph_cpu(g, sd, x):
    """Execute the folded graph with plain fp64 torch ops on CPU (test-only executor)."""
    N, _, H, W = x.shape
    bufs = [torch.zeros(N, b.c_total, H >> b.level, W >> b.level, dtype=torch.float64) for b in g.bufs]
    heads = {"cls": [None] * len(g.strides), "reg": [None] * len(g.strides)}

    def act(y, a):
        return {"relu": torch.relu, "silu": lambda t: t * torch.sigmoid(t), "sigmoid": torch.sigmoid, None: lambda t: t}[a](y)

    def rd(t):
        return bufs[t.buf][:, t.c_off:t.c_off + t.c]

    for op in g.ops:
        if op.kind == "pool":
            s = rd(op.src)
            y1 = F.max_pool2d(s, 5, 1, 2); y2 = F.max_pool2d(y1, 5, 1, 2); y3 = F.max_pool2d(y2, 5, 1, 2)
            bufs[op.dst.buf][:, op.cin:4 * op.cin] = torch.cat([y1, y2, y3], 1)
            continue
        w, b = fold_op(sd, op)
        if op.kind == "convT":
            s = rd(op.src)
            out = rd(op.dst)
            for q, wq in enumerate(w):
                out[:, :, q // 2::2, q % 2::2] = F.conv2d(s, wq.permute(0, 3, 1, 2), b)
            continue
        src = x.double() if op.kind == "stem" else rd(op.src)
        y = act(F.conv2d(src, w.permute(0, 3, 1, 2), b, stride=op.s, padding=op.k // 2), op.act)
        if op.res is not None:
            y = y + float(sd[op.alpha]) * rd(op.res)
        if op.kind == "pred":
            heads[op.head[0]][op.head[1]] = y.flatten(2).permute(0, 2, 1)
        else:
            rd(op.dst).copy_(y)
    return torch.cat(heads["cls"], 1), torch.cat(heads["reg"], 1)


@pytest.mark.parametrize("name", MODELS)
def test_folded_graph_equals_oracle(name):
    size = 64 if name != "yolov6l6" else 128
    keys = golden_keys(name)
    sd = fab.fabricate_state_dict(keys, 0)
    x = fab.synthetic_images(1, size, size, seed=3)
    g = arch.build_graph(configs.get_config(name), 80)
    with torch.no_grad():
        cls, reg = _run_graph_cpu(g, sd, x)
        ocls, oreg, _ = om.forward(sd, om.CONFIGS[name], x.double(), train_outputs=True)
    assert float((cls - ocls).abs().max()) < 1e-9
    assert float((reg - oreg).abs().max()) < 1e-8
