"""Pins oracle/model.py against golden vectors minted from the live reference
(tests/golden/make_golden.py): eval outputs, train-form head outputs and the reference's own
deploy-form (fuse_model + switch_to_deploy) outputs for YOLOv6-N/S/M/L6."""
import numpy as np
import pytest
import torch

from conftest import golden_keys, golden_npz
from oracle import fabricate as fab
from oracle import model as om

MODELS = {"yolov6n": 64, "yolov6s": 64, "yolov6m": 64, "yolov6l6": 128}


def rel_err(a, b):
    """max |a-b| / (1 + |b|): absolute for scores in [0,1], relative for pixel coordinates."""
    return float((np.abs(a - b) / (1.0 + np.abs(b))).max())


@pytest.mark.parametrize("name", list(MODELS))
def test_oracle_matches_reference(name):
    g = golden_npz(f"model_{name}.npz")
    keys = golden_keys(name)
    sd = fab.fabricate_state_dict(keys, seed=0)
    size = MODELS[name]
    x = fab.synthetic_images(2, size, size, seed=0)
    assert abs(fab.checksum(x) - float(g["x_checksum"])) < 1e-6 * abs(float(g["x_checksum"])), "input RNG drift"
    wsum = sum(fab.checksum(v) for v in sd.values())
    assert abs(wsum - float(g["w_checksum"])) < 1e-6 * abs(float(g["w_checksum"])), "weight RNG drift"
    cfg = om.CONFIGS[name]
    with torch.no_grad():
        out = om.forward(sd, cfg, x).numpy()
        cls, reg, _ = om.forward(sd, cfg, x, train_outputs=True)
        out64 = om.forward(sd, cfg, x.double()).numpy()
    # fp32 oracle vs fp32 reference: same math, different op order -> 1e-5 (north_star bar is 1e-4)
    assert rel_err(out, g["eval_out"]) < 1e-5
    assert rel_err(cls.numpy(), g["cls_train"]) < 1e-5
    assert rel_err(reg.numpy(), g["reg_train"]) < 1e-5
    # fp64 oracle brackets the reference's own rounding
    assert rel_err(out64, g["eval_out"]) < 1e-5
    # the reference's re-parameterised deploy form drifts from its train form by <= ~1.2e-4
    assert rel_err(g["deploy_out"], g["eval_out"]) < 2e-4
    assert rel_err(out64, g["deploy_out"]) < 2e-4


def test_config_scaling_matches_reference_shapes():
    # widths / repeats restated from yolo.py:66-67 must reproduce the reference's parameter shapes
    for name in MODELS:
        keys = dict(golden_keys(name))
        reps, chans = om.scaled_lists(om.CONFIGS[name])
        assert keys["backbone.stem.rbr_dense.conv.weight" if om.CONFIGS[name]["mode"] == "repvgg"
                    else "backbone.stem.block.conv.weight"][0] == chans[0]
        assert keys["detect.cls_preds.0.weight"][0] == 80
        assert keys["detect.reg_preds.0.weight"][0] == 4 * (om.CONFIGS[name]["reg_max"] + 1)
