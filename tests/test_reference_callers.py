"""The drop-in boundary under the reference's OWN callers (build container only: needs /root/reference).  Applies the
swap of INTEGRATION.md -- `yolov6_b200.build_model` where the reference calls `yolov6.models.yolo.build_model` -- and
drives the reference's unmodified `Config.fromfile`, `build_optimizer` (solver/build.py:10-33), `ModelEMA`
(utils/ema.py) and pickled-module checkpoints (utils/checkpoint.py:22-32) over the result.  CPU only: nothing here runs
a network forward (that needs the CUDA engine; see tests/test_gpu_*.py)."""
import copy
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "yolov6")), reason="reference checkout not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    sys.path[:0] = [os.path.join(HERE, "golden", "refshim"), REF]
    import torch.nn as nn
    cuda_was = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    from yolov6.models.yolo import build_model as ref_build_model
    from yolov6.solver.build import build_optimizer
    from yolov6.utils.config import Config
    from yolov6.utils.ema import ModelEMA
    yield dict(build_model=ref_build_model, build_optimizer=build_optimizer, Config=Config, ModelEMA=ModelEMA, nn=nn)
    torch.cuda.is_available = cuda_was


def load_cfg(ref, name):
    cfg = ref["Config"].fromfile(f"{REF}/configs/{name}.py")
    if not hasattr(cfg, "training_mode"):
        setattr(cfg, "training_mode", "repvgg")          # tools/train.py:99-100
    return cfg


@pytest.mark.parametrize("name", ["yolov6n", "yolov6s", "yolov6m", "yolov6l6"])
def test_build_model_from_reference_config_and_optimizer_groups(ref, name):
    from yolov6_b200.model import build_model
    cfg = load_cfg(ref, name)
    ours = build_model(cfg, 80, torch.device("cpu"))
    theirs = ref["build_model"](cfg, 80, torch.device("cpu"))
    sd_o, sd_t = ours.state_dict(), theirs.state_dict()
    assert sorted(sd_o) == sorted(sd_t)                  # same keys (registration order differs, loading is by key)
    assert all(sd_o[k].shape == sd_t[k].shape and sd_o[k].dtype == sd_t[k].dtype for k in sd_t)
    theirs.load_state_dict(sd_o, strict=True)            # and the state round-trips in both directions
    ours.load_state_dict(theirs.state_dict(), strict=True)
    assert torch.equal(ours.stride.float(), theirs.stride.float())
    # the reference's optimizer builder sees the same three parameter groups on both models
    o1, o2 = ref["build_optimizer"](cfg, ours), ref["build_optimizer"](cfg, theirs)
    sizes = lambda o: [sorted(tuple(p.shape) for p in g["params"]) for g in o.param_groups]   # noqa: E731
    assert sizes(o1) == sizes(o2)
    assert [g.get("weight_decay", 0) for g in o1.param_groups] == [g.get("weight_decay", 0) for g in o2.param_groups]
    assert o1.param_groups[0]["nesterov"] and o1.param_groups[0]["momentum"] == cfg.solver.momentum
    if name == "yolov6s":
        print("yolov6s parameter groups (bn weights, weights, biases):", [len(g["params"]) for g in o1.param_groups])
    # an optimizer step over zero gradients is a weight-decay-only update and runs through the views of the flat state
    for p in ours.parameters():
        if p.requires_grad:
            p.grad = torch.zeros_like(p)
    before = ours.state_dict()["backbone.ERBlock_2.0." + ("rbr_dense" if name != "yolov6l6" else "block") + ".conv.weight"].clone()
    o1.step()
    after = ours.state_dict()["backbone.ERBlock_2.0." + ("rbr_dense" if name != "yolov6l6" else "block") + ".conv.weight"]
    # nesterov, first step: g = wd*p, buf = g, p -= lr * (g + momentum * buf)
    assert torch.allclose(after, before * (1 - cfg.solver.lr0 * cfg.solver.weight_decay * (1 + cfg.solver.momentum)), rtol=1e-5, atol=1e-9)


def test_model_ema_over_the_drop_in_model(ref):
    from yolov6_b200.model import build_model
    cfg = load_cfg(ref, "yolov6n")
    m = build_model(cfg, 80, torch.device("cpu"))
    ema = ref["ModelEMA"](m)                              # deepcopy(model).eval()
    assert type(ema.ema) is type(m) and not ema.ema.training
    with torch.no_grad():
        for p in m.parameters():
            p.add_(1.0)
    ema.update(m)
    d = ema.decay(1)
    k = "backbone.stem.rbr_dense.conv.weight"
    want = (m.state_dict()[k] - 1.0) * d + (1 - d) * m.state_dict()[k]
    assert torch.allclose(ema.ema.state_dict()[k], want, rtol=1e-6, atol=1e-8)
    ema.update_attr(m, include=['nc', 'names', 'stride'])  # core/engine.py:186


def test_pickled_reference_checkpoint_converts(ref, tmp_path):
    """checkpoint.py:22-32 unpickles `yolov6.models.yolo.Model`; `yolov6_b200.checkpoint` turns it into the kernel-backed model."""
    from yolov6_b200.checkpoint import from_reference, load_checkpoint
    from yolov6_b200.model import Model
    cfg = load_cfg(ref, "yolov6s")
    theirs = ref["build_model"](cfg, 80, torch.device("cpu"))
    with torch.no_grad():
        for p in theirs.parameters():
            p.mul_(1.01)
    path = tmp_path / "last_ckpt.pt"
    torch.save({"model": copy.deepcopy(theirs).half(), "ema": None, "epoch": 3}, path)   # Trainer saves half (engine.py:185)
    m = load_checkpoint(str(path), map_location="cpu")
    assert isinstance(m, Model) and not m.training
    sd_t = theirs.state_dict()
    for k, a in m.state_dict().items():
        assert torch.allclose(a.float(), sd_t[k].half().float()), k
    m2 = from_reference(theirs.train())
    assert m2.training and m2.detect.nc == 80
    from yolov6.utils.torch_utils import fuse_model
    theirs = fuse_model(theirs.eval())                     # the reference's deploy order: fuse BN, then re-parameterise
    with pytest.raises(RuntimeError):
        for layer in theirs.modules():
            if hasattr(layer, "switch_to_deploy"):
                layer.switch_to_deploy()
        from_reference(theirs)
