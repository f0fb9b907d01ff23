"""tests/golden/make_golden_distill.py -- golden vectors of the self-distillation loss from the UNMODIFIED reference
(yolov6/models/losses/loss_distill.py, built as in core/engine.py:311-322 for a DFL model, distill_feat=False) on seeded synthetic
student / teacher head outputs: loss, the four loss items and the gradients w.r.t. the student's scores and distributions.

    PYTHONPATH=tests/golden/refshim:/root/reference:. python tests/golden/make_golden_distill.py
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

from yolov6.models.losses.loss_distill import ComputeLoss as ComputeLossDistill  # noqa: E402
from yolov6.models.losses.loss_distill_ns import ComputeLoss as ComputeLossDistillNS  # noqa: E402

from oracle import fabricate as fab  # noqa: E402
from oracle import loss as oloss  # noqa: E402

CASES = [  # name, img, B, seed, iou_type, warmup_epoch, epoch, max_epoch, temperature, drop
    ["distill_tal", 320, 3, 31, "giou", 0, 40, 300, 20.0, None],
    ["distill_atss_empty_image", 320, 3, 33, "giou", 4, 1, 300, 20.0, [2]],
]


def main():
    store = {}
    for name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop in CASES:
        strides = [8, 16, 32]
        sizes = [(img // s, img // s) for s in strides]
        ps, pd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed)
        tps, tpd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed + 100)
        targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
        cl = ComputeLossDistill(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=warm, use_dfl=True, reg_max=16,
                                iou_type=iou_type, distill_weight={"class": 1.0, "dfl": 1.0}, distill_feat=False)
        psl, pdl = ps.clone().requires_grad_(True), pd.clone().requires_grad_(True)
        feats = [torch.zeros(B, 8, h, w) for h, w in sizes]
        loss, items = cl((feats, psl, pdl), (feats, tps, tpd), None, None, targets.clone(), epoch, max_epoch, T, 1, img, img)
        loss.backward()
        store[f"{name}_loss"] = np.float64(loss.item())
        store[f"{name}_items"] = items.double().numpy()
        store[f"{name}_in_checksum"] = np.float64(fab.checksum(ps) + fab.checksum(pd) + fab.checksum(tps) + fab.checksum(tpd) + fab.checksum(targets))
        store[f"{name}_grad_scores_abs"] = np.float64(psl.grad.double().abs().sum().item())
        store[f"{name}_grad_scores_head"] = psl.grad.flatten()[:4096].double().numpy()
        store[f"{name}_grad_distri_abs"] = np.float64(pdl.grad.double().abs().sum().item())
        nz = pdl.grad.abs().sum(-1) > 0
        store[f"{name}_pos"] = np.packbits(nz.numpy())
        store[f"{name}_grad_distri_rows"] = pdl.grad[nz].double().numpy()
        print(name, "loss", loss.item(), "items", items.tolist(), "positives", int(nz.sum()))
    # ---- feature-map term (distill_feat=True, loss_distill.py:223-245) on seeded stand-ins for the three neck outputs
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = ["feat_" + CASES[0][0]] + CASES[0][1:]
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed)
    tps, tpd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed + 100)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
    gf = torch.Generator().manual_seed(seed + 300)
    s_feats = [torch.randn(B, c, h, w, generator=gf) for c, (h, w) in zip((32, 64, 128), sizes)]
    t_feats = [torch.randn(B, c, h, w, generator=gf) * 1.3 for c, (h, w) in zip((32, 64, 128), sizes)]
    cl = ComputeLossDistill(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=warm, use_dfl=True, reg_max=16,
                            iou_type=iou_type, distill_weight={"class": 1.0, "dfl": 1.0}, distill_feat=True)
    psl, pdl = ps.clone().requires_grad_(True), pd.clone().requires_grad_(True)
    sfl = [f.clone().requires_grad_(True) for f in s_feats]
    feats = [torch.zeros(B, 8, h, w) for h, w in sizes]
    loss, items = cl((feats, psl, pdl), (feats, tps, tpd), sfl, t_feats, targets.clone(), epoch, max_epoch, T, 1, img, img)
    loss.backward()
    store[f"{name}_loss"] = np.float64(loss.item())
    store[f"{name}_items"] = items.double().numpy()
    store[f"{name}_feat_checksum"] = np.float64(sum(fab.checksum(f) for f in s_feats + t_feats))
    for l, f in enumerate(sfl):
        store[f"{name}_grad_feat{l}_abs"] = np.float64(f.grad.double().abs().sum().item())
        store[f"{name}_grad_feat{l}_head"] = f.grad.flatten()[:2048].double().numpy()
    print(name, "loss", loss.item(), "items", items.tolist())
    # ---- N / S variant (loss_distill_ns.py): a third student tensor, the lrtb distances of the inference branch
    for name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop in [["ns_" + c[0]] + c[1:] for c in CASES]:
        strides = [8, 16, 32]
        sizes = [(img // s, img // s) for s in strides]
        ps, pd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed)
        _, pl = fab.synthetic_head_outputs(B, sizes, 80, 4, seed + 200)
        tps, tpd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed + 100)
        targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
        cl = ComputeLossDistillNS(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=warm, use_dfl=True, reg_max=16,
                                  iou_type=iou_type, distill_weight={"class": 1.0, "dfl": 1.0}, distill_feat=False)
        psl, pdl, pll = ps.clone().requires_grad_(True), pd.clone().requires_grad_(True), pl.clone().requires_grad_(True)
        feats = [torch.zeros(B, 8, h, w) for h, w in sizes]
        loss, items = cl((feats, psl, pdl, pll), (feats, tps, tpd), None, None, targets.clone(), epoch, max_epoch, T, 1, img, img)
        loss.backward()
        nz = pll.grad.abs().sum(-1) > 0
        store[f"{name}_loss"] = np.float64(loss.item())
        store[f"{name}_items"] = items.double().numpy()
        store[f"{name}_pos"] = np.packbits(nz.numpy())
        store[f"{name}_grad_lrtb_rows"] = pll.grad[nz].double().numpy()
        store[f"{name}_grad_distri_rows"] = pdl.grad[nz].double().numpy()
        store[f"{name}_grad_scores_abs"] = np.float64(psl.grad.double().abs().sum().item())
        print(name, "loss", loss.item(), "items", items.tolist(), "positives", int(nz.sum()))
    with open(os.path.join(HERE, "distill_cases.json"), "w") as f:
        json.dump(CASES, f)
    np.savez_compressed(os.path.join(HERE, "distill.npz"), **store)


if __name__ == "__main__":
    main()
