"""tests/golden/make_golden_fuseab.py -- golden vectors of the anchor-aided (fuse_ab) branch from the UNMODIFIED reference.

A. Model: YOLOv6-N built with `fuse_ab=True` (yolo.py:122-126 -> heads/effidehead_fuseab.py), `.train()` mode, float64,
   fabricated weights, 2 x 3 x 64 x 64 input: the five training outputs, L = sum(cls_ab * w1) + sum(reg_ab * w2) and its
   gradients w.r.t. the two extra pred convs of every level and a few upstream tensors (autograd).
B. Loss: `yolov6.models.losses.loss_fuseab.ComputeLoss` (as the Trainer builds it, core/engine.py:298-309) on seeded synthetic
   ab-head outputs, 4 images of 320 x 320: loss, loss_items, the foreground mask and the gradients w.r.t. both inputs.

    PYTHONPATH=tests/golden/refshim:/root/reference:. python tests/golden/make_golden_fuseab.py
"""
import json
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

from yolov6.models.losses.loss_fuseab import ComputeLoss as ComputeLossAB  # noqa: E402
from yolov6.models.yolo import build_model  # noqa: E402
from yolov6.utils.config import Config  # noqa: E402

from oracle import fabricate as fab  # noqa: E402
from oracle import loss as oloss  # noqa: E402
from oracle import loss_fuseab as oab  # noqa: E402

FULL = ["detect.cls_preds_ab.0.weight", "detect.cls_preds_ab.2.bias", "detect.reg_preds_ab.1.weight", "detect.reg_preds_ab.0.bias",
        "detect.reg_convs.1.block.bn.bias", "detect.cls_convs.2.block.conv.weight", "neck.Rep_n4.block.0.rbr_1x1.conv.weight"]
LOSS_CASES = [  # name, img, B, seed, iou_type, drop (images whose targets are removed)
    ["ab_giou", 320, 4, 21, "giou", None],
    ["ab_siou_empty_image", 320, 3, 23, "siou", [1]],
]


def main():
    store = {}
    # ---------------- A: model ----------------
    cfg = Config.fromfile("/root/reference/configs/yolov6n.py")
    if not hasattr(cfg, "training_mode"):
        setattr(cfg, "training_mode", "repvgg")
    m = build_model(cfg, 80, torch.device("cpu"), fuse_ab=True)
    keys = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    with open(os.path.join(HERE, "keys_yolov6n_fuseab.json"), "w") as f:
        json.dump(keys, f)
    sd = fab.fabricate_state_dict(keys, seed=0)
    for k in sd:      # keep the head logits O(1) under batch-statistics BN
        if (".cls_preds" in k or ".reg_preds" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    m.load_state_dict(sd, strict=True)
    m = m.double().train()
    x = fab.synthetic_images(2, 64, 64, seed=7).double()
    (feats, cls_ab, reg_ab, cls, reg), _ = m(x)
    g = torch.Generator().manual_seed(13)
    w1 = torch.randn(cls_ab.shape, generator=g).double()
    w2 = torch.randn(reg_ab.shape, generator=g).double()
    L = (cls_ab * w1).sum() + (reg_ab * w2).sum()
    L.backward()
    params = dict(m.named_parameters())
    names = [k for k, p in params.items() if p.grad is not None]
    store.update(m_cls_ab=cls_ab.detach().numpy(), m_reg_ab=reg_ab.detach().numpy(), m_cls=cls.detach().numpy(), m_reg=reg.detach().numpy(),
                 m_L=np.float64(L.item()), m_grad_names=np.array(names), m_grad_norms=np.array([float(params[k].grad.norm()) for k in names]),
                 m_x_checksum=np.float64(fab.checksum(x.float())))
    for k in FULL:
        store["m_grad::" + k] = params[k].grad.numpy()
    print("model: L", L.item(), "cls_ab", tuple(cls_ab.shape), "reg_ab", tuple(reg_ab.shape), "params with grad", len(names))
    # ---------------- A2: the N / S distillation student head (yolo.py:113-120 -> heads/effidehead_distill_ns.py) ----------------
    cfg = Config.fromfile("/root/reference/configs/yolov6n.py")
    if not hasattr(cfg, "training_mode"):
        setattr(cfg, "training_mode", "repvgg")
    cfg.model.head.use_dfl, cfg.model.head.reg_max = True, 16      # "set to True / 16 if you want to further train with distillation"
    m = build_model(cfg, 80, torch.device("cpu"), distill_ns=True)
    keys = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    with open(os.path.join(HERE, "keys_yolov6n_distill_ns.json"), "w") as f:
        json.dump(keys, f)
    sd = fab.fabricate_state_dict(keys, seed=0)
    for k in sd:
        if (".cls_preds" in k or ".reg_preds" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    m.load_state_dict(sd, strict=True)
    m = m.double().train()
    x = fab.synthetic_images(2, 64, 64, seed=7).double()
    (feats, cls, reg_dist, reg_lrtb), _ = m(x)
    g = torch.Generator().manual_seed(17)
    w0, w1, w2 = (torch.randn(t.shape, generator=g).double() for t in (cls, reg_dist, reg_lrtb))
    L = (cls * w0).sum() + (reg_dist * w1).sum() + (reg_lrtb * w2).sum()
    L.backward()
    params = dict(m.named_parameters())
    names = [k for k, p in params.items() if p.grad is not None]
    store.update(ns_cls=cls.detach().numpy(), ns_reg_dist=reg_dist.detach().numpy(), ns_reg_lrtb=reg_lrtb.detach().numpy(), ns_L=np.float64(L.item()),
                 ns_grad_names=np.array(names), ns_grad_norms=np.array([float(params[k].grad.norm()) for k in names]))
    for k in ("detect.reg_preds_dist.1.weight", "detect.reg_preds.0.bias", "detect.reg_convs.2.block.conv.weight"):
        store["ns_grad::" + k] = params[k].grad.numpy()
    m.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}, strict=True)   # the train forward moved the running statistics
    m.eval()
    with torch.no_grad():
        store["ns_eval"] = m(x)[0].numpy()
    print("distill_ns model: L", L.item(), "reg_dist", tuple(reg_dist.shape), "reg_lrtb", tuple(reg_lrtb.shape), "params with grad", len(names))
    # ---------------- B: loss ----------------
    for name, img, B, seed, iou_type, drop in LOSS_CASES:
        strides = [8, 16, 32]
        sizes = [(img // s, img // s) for s in strides]
        ps, pd = oab.synthetic_ab_outputs(B, sizes, 80, seed)
        targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
        cl = ComputeLossAB(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type=iou_type)
        psl, pdl = ps.clone().requires_grad_(True), pd.clone().requires_grad_(True)
        feats = [torch.zeros(B, 8, h, w) for h, w in sizes]
        # the reference adds the anchor points to pred_distri IN PLACE (loss_fuseab.py:72): hand it a non-leaf copy
        loss, items = cl((feats, psl * 1.0, pdl * 1.0), targets.clone(), 0, 1, img, img)
        loss.backward()
        # assignment of the same call, through the reference's own assigner on the same boxes
        with torch.no_grad():
            _, _, a = oab.compute_loss_ab(sizes, ps, pd, targets, strides=strides, ori_img_size=img, iou_type=iou_type, return_assign=True)
        fg = a["fg"].numpy()
        store[f"{name}_loss"] = np.float64(loss.item())
        store[f"{name}_items"] = items.double().numpy()
        store[f"{name}_in_checksum"] = np.float64(fab.checksum(ps) + fab.checksum(pd) + fab.checksum(targets))
        store[f"{name}_grad_scores_abs"] = np.float64(psl.grad.double().abs().sum().item())
        store[f"{name}_grad_scores_head"] = psl.grad.flatten()[:4096].double().numpy()
        store[f"{name}_grad_distri_abs"] = np.float64(pdl.grad.double().abs().sum().item())
        nz = pdl.grad.abs().sum(-1) > 0
        store[f"{name}_fg_from_grad"] = np.packbits(nz.numpy())            # rows with a box gradient = the reference's positives
        store[f"{name}_grad_distri_rows"] = pdl.grad[nz].double().numpy()
        store[f"{name}_grad_scores_rows"] = psl.grad[nz].double().numpy()
        print(name, "loss", loss.item(), "items", items.tolist(), "positives", int(nz.sum()), "oracle positives", int(fg.sum()))
    with open(os.path.join(HERE, "fuseab_cases.json"), "w") as f:
        json.dump(LOSS_CASES, f)
    np.savez_compressed(os.path.join(HERE, "fuseab.npz"), **store)


if __name__ == "__main__":
    main()
