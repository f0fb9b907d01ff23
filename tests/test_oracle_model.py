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


@pytest.mark.parametrize("name,batch,size", [("yolov6n", 4, 64), ("yolov6m", 2, 64)])
def test_oracle_train_mode_matches_reference(name, batch, size):
    """Train mode (batch-statistics BatchNorm, train branch of Detect, BottleRep alpha) of oracle/model.py against
    the reference model run in .train() mode in float64 (tests/golden/make_golden_train.py): head outputs, the
    scalar L = sum(cls*wc) + sum(reg*wr), the gradient norm of EVERY parameter and selected full gradients."""
    g = golden_npz(f"train_{name}.npz")
    keys = golden_keys(name)
    sd = fab.fabricate_state_dict(keys, seed=0)
    for k in sd:
        if (".cls_preds." in k or ".reg_preds." in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
        if k.endswith(".alpha"):
            sd[k] = sd[k] * 0.75
    x = fab.synthetic_images(batch, size, size, seed=7)
    assert abs(fab.checksum(x) - float(g["x_checksum"])) < 1e-6 * abs(float(g["x_checksum"])), "input RNG drift"
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    with om.train_mode():
        cls, reg, _ = om.forward(sd64, om.CONFIGS[name], x.double(), train_outputs=True)
    assert rel_err(cls.detach().numpy(), g["cls"]) < 1e-9
    assert rel_err(reg.detach().numpy(), g["reg"]) < 1e-9
    gen = torch.Generator().manual_seed(11)
    wc = torch.randn(cls.shape, generator=gen).double()
    wr = torch.randn(reg.shape, generator=gen).double()
    L = (cls * wc).sum() + (reg * wr).sum()
    assert abs(L.item() - float(g["L"])) < 1e-8 * max(1.0, abs(float(g["L"])))
    L.backward()
    names, norms = [str(n) for n in g["grad_names"]], g["grad_norms"]
    assert len(names) > 300
    for n, ref in zip(names, norms):
        assert sd64[n].grad is not None, n
        got = float(sd64[n].grad.norm())
        assert abs(got - ref) <= 1e-7 * max(1.0, ref), (n, got, ref)
    for k in g.files:
        if k.startswith("grad::"):
            n = k[6:]
            np.testing.assert_allclose(sd64[n].grad.numpy().reshape(g[k].shape), g[k], rtol=1e-7, atol=1e-9 * (1 + np.abs(g[k]).max()))


@pytest.mark.parametrize("name,B,size,step", [("yolov6s", 4, 640, 16), ("yolov6l6", 1, 1280, 32)])
def test_oracle_matches_reference_at_benchmark_size(name, B, size, step):
    """BASELINE.json configs 2 / 5: 640x640 (A = 8400) and 1280x1280 (A = 34000) goldens of make_golden_configs.py."""
    g = golden_npz("configs.npz")
    sd = fab.fabricate_state_dict(golden_keys(name), seed=0)
    x = fab.synthetic_images(B, size, size, seed=40)
    assert abs(fab.checksum(x) - float(g[f"{name}_x_checksum"])) < 1e-9 * abs(float(g[f"{name}_x_checksum"])), "input RNG drift"
    with torch.no_grad():
        out = om.forward(sd, om.CONFIGS[name], x).double().numpy()
    assert rel_err(out[:, ::step], g[f"{name}_rows"].astype(np.float64)) < 1e-5
    A = out.shape[1]
    assert float((np.abs(out.sum(1) - g[f"{name}_colsum"]) / (A + g[f"{name}_abs_colsum"])).max()) < 1e-5
