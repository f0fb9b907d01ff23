"""tests/golden/make_golden_train.py -- train-mode golden vectors from the UNMODIFIED reference.

For YOLOv6-N (RepVGG blocks) and YOLOv6-M (BepC3 / BottleRep with learnable alpha, DFL head) the reference
model is run in `.train()` mode (batch-statistics BatchNorm, train branch of Detect.forward,
models/yolo.py:33-41, effidehead.py:72-92) in float64 on CPU with fabricated weights; a scalar
L = sum(cls * wc) + sum(reg * wr) is back-propagated with torch autograd.  Stored: the head outputs, L, the
gradient norm of every parameter and full gradients of a few representative tensors, and the running
statistics of one BatchNorm after the step.  `tests/test_oracle_model.py::test_oracle_train_mode_matches_reference`
holds oracle/model.py's train mode (forward + autograd) to them.

    PYTHONPATH=tests/golden/refshim:/root/reference:. python tests/golden/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]

torch.cuda.is_available = lambda: False
nn.Module.cuda = lambda self, *a, **k: self

from yolov6.models.yolo import build_model  # noqa: E402
from yolov6.utils.config import Config  # noqa: E402

from oracle import fabricate as fab  # noqa: E402

CASES = {"yolov6n": (4, 64), "yolov6m": (2, 64)}     # name -> (batch, size)
FULL = {  # gradients stored in full
    "yolov6n": ["backbone.stem.rbr_dense.conv.weight", "backbone.ERBlock_3.1.block.0.rbr_identity.weight",
                "neck.Bifusion0.upsample.upsample_transpose.weight", "detect.reg_convs.1.block.bn.bias",
                "detect.cls_preds.2.weight", "detect.reg_preds.0.bias"],
    "yolov6m": ["backbone.stem.rbr_dense.conv.weight", "backbone.ERBlock_2.1.m.conv1.alpha",
                "backbone.ERBlock_2.1.cv3.block.conv.weight", "neck.Rep_p4.m.conv1.conv1.rbr_1x1.conv.weight",
                "detect.reg_preds.1.weight"],
}


def load_cfg(name):
    cfg = Config.fromfile(f"/root/reference/configs/{name}.py")
    if not hasattr(cfg, "training_mode"):
        setattr(cfg, "training_mode", "repvgg")
    return cfg


def main():
    for name, (B, size) in CASES.items():
        m = build_model(load_cfg(name), 80, torch.device("cpu"))
        keys = [(k, list(v.shape)) for k, v in m.state_dict().items()]
        sd = fab.fabricate_state_dict(keys, seed=0)
        for k in sd:      # keep the head logits O(1) under batch-statistics BN
            if (".cls_preds." in k or ".reg_preds." in k) and k.endswith("weight"):
                sd[k] = sd[k] * 0.1
            if k.endswith(".alpha"):
                sd[k] = sd[k] * 0.75
        m.load_state_dict(sd, strict=True)
        m = m.double()
        m.train()
        x = fab.synthetic_images(B, size, size, seed=7).double()
        (feats, cls, reg), _ = m(x)
        g = torch.Generator().manual_seed(11)
        wc = torch.randn(cls.shape, generator=g).double()
        wr = torch.randn(reg.shape, generator=g).double()
        L = (cls * wc).sum() + (reg * wr).sum()
        L.backward()
        names = [k for k, p in m.named_parameters() if p.grad is not None]
        store = dict(cls=cls.detach().numpy(), reg=reg.detach().numpy(), L=np.float64(L.item()),
                     grad_names=np.array(names), grad_norms=np.array([float(p.grad.norm()) for k, p in m.named_parameters()
                                                                       if p.grad is not None]),
                     x_checksum=np.float64(fab.checksum(x.float())))
        params = dict(m.named_parameters())
        missing = [k for k in FULL[name] if k not in params]
        assert not missing, missing
        for k in FULL[name]:
            store["grad::" + k] = params[k].grad.numpy()
        bn = "backbone.ERBlock_2.0.rbr_dense.bn" if name == "yolov6n" else "backbone.ERBlock_2.0.rbr_dense.bn"
        bufs = dict(m.named_buffers())
        store["bn_name"] = np.array(bn)
        store["running_mean"] = bufs[bn + ".running_mean"].numpy()
        store["running_var"] = bufs[bn + ".running_var"].numpy()
        np.savez_compressed(os.path.join(HERE, f"train_{name}.npz"), **store)
        print(name, "L", L.item(), "params with grad", len(names), "feat shapes", [tuple(f.shape) for f in feats])


if __name__ == "__main__":
    main()
