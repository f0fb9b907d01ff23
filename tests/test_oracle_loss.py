"""Pins oracle/assign.py + oracle/loss.py against the reference's ComputeLoss (TAL / ATSS, VFL,
GIoU/SIoU, DFL) outputs and gradients stored by tests/golden/make_golden.py."""
import numpy as np
import pytest
import torch

from conftest import golden_json, golden_npz
from oracle import fabricate as fab
from oracle import loss as oloss

CASES = golden_json("loss_cases.json")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_loss_and_assignment_match_reference(case):
    name, nc, strides, img, use_dfl, reg_max, iou_type, warm, epoch, B, seed = case[:11]
    g = golden_npz("loss.npz")
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, nc, 4 * (reg_max + 1), seed)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=nc), case[11] if len(case) > 11 else None)
    chk = fab.checksum(ps) + fab.checksum(pd) + fab.checksum(targets)
    assert abs(chk - float(g[f"{name}_in_checksum"])) <= 1e-9 * abs(chk), "RNG drift"
    ps.requires_grad_(True)
    pd.requires_grad_(True)
    loss, items, asg = oloss.compute_loss(sizes, ps, pd, targets, strides=strides, num_classes=nc, ori_img_size=img,
                                          warmup_epoch=warm, epoch_num=epoch, use_dfl=use_dfl, reg_max=reg_max,
                                          iou_type=iou_type, return_assign=True)
    g_ps, g_pd = torch.autograd.grad(loss, [ps, pd], allow_unused=True)
    # integer outputs exact
    fg = asg["fg"].numpy()
    assert np.array_equal(np.packbits(fg), g[f"{name}_fg"])
    ref_labels = g[f"{name}_labels"].astype(np.int64)
    ref_labels_bg = np.where(fg, ref_labels, nc)           # the reference relabels background afterwards
    assert np.array_equal(asg["labels"].numpy(), ref_labels_bg)
    nz = asg["scores"].nonzero().numpy().astype(np.int32)
    assert np.array_equal(nz, g[f"{name}_scores_idx"])
    # float outputs: same fp64 math -> tight.  During the ATSS warm-up epochs the reference itself computes the soft labels
    # and hence the three loss sums in float32 (atss_assigner.py:86-92: `target_scores *= ious` in place), so those cases carry
    # float32 summation-order noise (the result depends on torch's CPU thread partitioning): 2e-6 instead of 1e-10.
    tol = 2e-6 if epoch < warm else 1e-10
    np.testing.assert_allclose(asg["scores"][asg["scores"] != 0].double().numpy(), g[f"{name}_scores_val"], rtol=1e-12, atol=0)
    assert abs(loss.item() - float(g[f"{name}_loss"])) <= tol * abs(loss.item())
    np.testing.assert_allclose(items.double().numpy(), g[f"{name}_items"], rtol=tol, atol=1e-12)
    np.testing.assert_allclose(g_ps[asg["fg"]].double().numpy(), g[f"{name}_grad_scores_fg"], rtol=max(1e-9, 10 * tol), atol=1e-12)
    np.testing.assert_allclose(g_ps.flatten()[:4096].double().numpy(), g[f"{name}_grad_scores_head"], rtol=max(1e-9, 10 * tol), atol=1e-12)
    assert abs(g_ps.double().abs().sum().item() - float(g[f"{name}_grad_scores_abs"])) <= max(1e-9, 10 * tol) * float(g[f"{name}_grad_scores_abs"])
    if g_pd is not None and f"{name}_grad_distri_fg" in g:
        ref_gd = g[f"{name}_grad_distri_fg"]
        np.testing.assert_allclose(g_pd[asg["fg"]].double().numpy(), ref_gd, rtol=max(1e-9, 10 * tol),
                                   atol=1e-12 if tol < 1e-9 else 1e-6 * np.abs(ref_gd).max())


def test_preprocess_pads_and_scales():
    t = torch.tensor([[0, 3, 0.5, 0.5, 0.2, 0.4], [2, 7, 0.25, 0.75, 0.1, 0.1], [0, 1, 0.1, 0.1, 0.1, 0.1]])
    out = oloss.preprocess_targets(t, 3, torch.tensor([640.0] * 4))
    assert out.shape == (3, 2, 5) and out.dtype == torch.float64
    assert out[1].tolist() == [[-1, 0, 0, 0, 0]] * 2
    np.testing.assert_allclose(out[0, 0].numpy(), [3, 256, 192, 384, 448], rtol=1e-6)
    assert oloss.preprocess_targets(torch.zeros(0, 6), 2, torch.tensor([640.0] * 4)).shape == (2, 0, 5)


def test_loss_at_640_batch_32_matches_reference():
    """BASELINE.json config 3's loss: 640x640, batch 32, A = 8400, COCO-shaped targets (make_golden_configs.py)."""
    g = golden_npz("configs.npz")
    B, img, strides, nc = 32, 640, [8, 16, 32], 80
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, nc, 4, seed=60)
    targets = oloss.synthetic_targets(B, seed=61, num_classes=nc)
    chk = fab.checksum(ps) + fab.checksum(pd) + fab.checksum(targets)
    assert abs(chk - float(g["loss640_in_checksum"])) <= 1e-9 * abs(chk), "RNG drift"
    ps.requires_grad_(True)
    pd.requires_grad_(True)
    loss, items, asg = oloss.compute_loss(sizes, ps, pd, targets, strides=strides, num_classes=nc, ori_img_size=img, warmup_epoch=0,
                                          epoch_num=0, use_dfl=False, reg_max=0, iou_type="giou", return_assign=True)
    g_ps, g_pd = torch.autograd.grad(loss, [ps, pd])
    fg = asg["fg"].numpy()
    assert np.array_equal(np.packbits(fg), g["loss640_fg"])
    assert np.array_equal(asg["labels"].numpy()[fg], g["loss640_labels_fg"].astype(np.int64))
    assert abs(loss.item() - float(g["loss640_loss"])) <= 1e-10 * abs(loss.item())
    np.testing.assert_allclose(items.double().numpy(), g["loss640_items"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(g_pd[asg["fg"]].double().numpy(), g["loss640_grad_distri_fg"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g_ps.double().abs().sum(-1)[:, ::64].numpy(), g["loss640_grad_scores_rowabs"], rtol=1e-9, atol=1e-12)


def _small_case(seed=5, B=3, img=256, nc=8, reg_max=16):
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, nc, 4 * (reg_max + 1), seed)
    targets = oloss.synthetic_targets(B, seed=seed + 1, num_classes=nc)
    kw = dict(strides=strides, num_classes=nc, ori_img_size=img, use_dfl=True, reg_max=reg_max, iou_type="giou", epoch_num=5)
    return sizes, ps, pd, targets, kw


def test_loss_size_independent_properties():
    """Properties of ComputeLoss (loss.py:57-182) that hold at every size -- the GPU path is held to the same through its parity
    with this oracle: the order of the target rows does not matter (preprocess groups them per image, loss.py:184-192), the three
    terms scale linearly with their loss weights (:171-177), an image without targets contributes no positives, and removing
    every target leaves the class term alone (background-only VarifocalLoss, :161-169)."""
    sizes, ps, pd, targets, kw = _small_case()
    loss, items, asg = oloss.compute_loss(sizes, ps, pd, targets, return_assign=True, **kw)
    assert loss.item() > 0 and int(asg["fg"].sum()) > 0
    # (a) permuting the target rows (keeping each image's relative order, which decides ties between identical boxes)
    shuffled = torch.cat([targets[targets[:, 0] == b] for b in reversed(range(ps.shape[0]))])
    assert not torch.equal(shuffled, targets) and shuffled.shape == targets.shape
    loss_p, items_p = oloss.compute_loss(sizes, ps, pd, shuffled, **kw)
    assert abs(loss_p.item() - loss.item()) <= 1e-12 * abs(loss.item())
    # (b) linear in the loss weights
    w2 = {"class": 2.0, "iou": 7.5, "dfl": 0.25}
    loss_w, items_w = oloss.compute_loss(sizes, ps, pd, targets, loss_weight=w2, **kw)
    base = {"class": 1.0, "iou": 2.5, "dfl": 0.5}
    scale = torch.tensor([w2["iou"] / base["iou"], w2["dfl"] / base["dfl"], w2["class"] / base["class"]], dtype=items.dtype)
    np.testing.assert_allclose(items_w.numpy(), (items * scale).numpy(), rtol=1e-12)
    assert abs(loss_w.item() - float(items_w.sum())) <= 1e-12 * abs(loss_w.item())
    # (c) an image without targets has no positive anchors
    no_img1 = targets[targets[:, 0] != 1]
    _, _, asg1 = oloss.compute_loss(sizes, ps, pd, no_img1, return_assign=True, **kw)
    assert int(asg1["fg"][1].sum()) == 0 and int(asg1["fg"][0].sum()) == int(asg["fg"][0].sum())
    # (d) no targets at all: IoU and DFL terms vanish, the class term is the background VarifocalLoss, not normalised (sum <= 1)
    loss0, items0 = oloss.compute_loss(sizes, ps, pd, targets[:0], **kw)
    bg = (torch.nn.functional.binary_cross_entropy(ps, torch.zeros_like(ps), reduction="none") * 0.75 * ps.pow(2)).sum()
    assert items0[0].item() == 0 and items0[1].item() == 0
    assert abs(items0[2].item() - bg.item()) <= 1e-12 * bg.item() and abs(loss0.item() - bg.item()) <= 1e-12 * bg.item()
