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
