"""The oracle's self-distillation loss against the reference's goldens (tests/golden/make_golden_distill.py:
losses/loss_distill.py of the unmodified reference, DFL model, distill_feat=False)."""
import numpy as np
import pytest
import torch

from conftest import golden_json, golden_npz
from oracle import fabricate as fab
from oracle import loss as oloss
from oracle import loss_distill as odist


def make_inputs(case):
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = case
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed)
    tps, tpd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed + 100)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
    return strides, sizes, ps, pd, tps, tpd, targets


@pytest.mark.parametrize("case", golden_json("distill_cases.json"), ids=lambda c: c[0])
def test_distill_loss_matches_reference(case):
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = case
    g = golden_npz("distill.npz")
    strides, sizes, ps, pd, tps, tpd, targets = make_inputs(case)
    chk = fab.checksum(ps) + fab.checksum(pd) + fab.checksum(tps) + fab.checksum(tpd) + fab.checksum(targets)
    assert abs(chk - float(g[f"{name}_in_checksum"])) < 1e-6 * abs(chk), "input RNG drift"
    psl, pdl = ps.clone().requires_grad_(True), pd.clone().requires_grad_(True)
    loss, items = odist.compute_loss_distill(sizes, psl, pdl, tps, tpd, targets, strides=strides, epoch_num=epoch, max_epoch=max_epoch,
                                             temperature=T, ori_img_size=img, warmup_epoch=warm, iou_type=iou_type)
    assert abs(loss.item() - float(g[f"{name}_loss"])) <= 1e-5 * abs(float(g[f"{name}_loss"]))
    np.testing.assert_allclose(items.double().numpy(), g[f"{name}_items"], rtol=1e-5, atol=1e-7)
    loss.backward()
    nz = (pdl.grad.abs().sum(-1) > 0)
    assert np.array_equal(np.packbits(nz.numpy()), g[f"{name}_pos"])
    np.testing.assert_allclose(pdl.grad[nz].double().numpy(), g[f"{name}_grad_distri_rows"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(psl.grad.flatten()[:4096].double().numpy(), g[f"{name}_grad_scores_head"], rtol=2e-4, atol=1e-7)   # fp32 softmax at T = 20: differences of ~4e-8 absolute
    assert abs(float(psl.grad.double().abs().sum()) - float(g[f"{name}_grad_scores_abs"])) <= 1e-4 * float(g[f"{name}_grad_scores_abs"])


@pytest.mark.parametrize("case", golden_json("distill_cases.json"), ids=lambda c: "ns_" + c[0])
def test_distill_ns_loss_matches_reference(case):
    """loss_distill_ns.ComputeLoss: TaskAlignedAssigner at every epoch, DFL distillation branch + IoU of the lrtb branch."""
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = case
    name = "ns_" + name
    g = golden_npz("distill.npz")
    strides, sizes, ps, pd, tps, tpd, targets = make_inputs(case)
    _, pl = fab.synthetic_head_outputs(B, sizes, 80, 4, seed + 200)
    psl, pdl, pll = ps.clone().requires_grad_(True), pd.clone().requires_grad_(True), pl.clone().requires_grad_(True)
    loss, items = odist.compute_loss_distill_ns(sizes, psl, pdl, pll, tps, tpd, targets, strides=strides, epoch_num=epoch, max_epoch=max_epoch,
                                                temperature=T, ori_img_size=img, iou_type=iou_type)
    assert abs(loss.item() - float(g[f"{name}_loss"])) <= 1e-5 * abs(float(g[f"{name}_loss"]))
    np.testing.assert_allclose(items.double().numpy(), g[f"{name}_items"], rtol=1e-5, atol=1e-7)
    loss.backward()
    nz = (pll.grad.abs().sum(-1) > 0)
    assert np.array_equal(np.packbits(nz.numpy()), g[f"{name}_pos"])
    np.testing.assert_allclose(pll.grad[nz].double().numpy(), g[f"{name}_grad_lrtb_rows"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(pdl.grad[nz].double().numpy(), g[f"{name}_grad_distri_rows"], rtol=2e-4, atol=1e-7)
    assert abs(float(psl.grad.double().abs().sum()) - float(g[f"{name}_grad_scores_abs"])) <= 1e-4 * float(g[f"{name}_grad_scores_abs"])


def _feat_inputs(case):
    name, img, B, seed = case[0], case[1], case[2], case[3]
    sizes = [(img // s, img // s) for s in (8, 16, 32)]
    gf = torch.Generator().manual_seed(seed + 300)
    s_feats = [torch.randn(B, c, h, w, generator=gf) for c, (h, w) in zip((32, 64, 128), sizes)]
    t_feats = [torch.randn(B, c, h, w, generator=gf) * 1.3 for c, (h, w) in zip((32, 64, 128), sizes)]
    return s_feats, t_feats


def test_distill_feature_term_matches_reference():
    """distill_feat=True (loss_distill.py:223-245): loss, the four items and the gradient w.r.t. the student's feature maps."""
    case = golden_json("distill_cases.json")[0]
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = case
    g = golden_npz("distill.npz")
    strides, sizes, ps, pd, tps, tpd, targets = make_inputs(case)
    s_feats, t_feats = _feat_inputs(case)
    assert abs(sum(fab.checksum(f) for f in s_feats + t_feats) - float(g[f"feat_{name}_feat_checksum"])) < 1e-6 * abs(float(g[f"feat_{name}_feat_checksum"]))
    sfl = [f.clone().requires_grad_(True) for f in s_feats]
    loss, items = odist.compute_loss_distill(sizes, ps, pd, tps, tpd, targets, strides=strides, epoch_num=epoch, max_epoch=max_epoch, temperature=T,
                                             ori_img_size=img, warmup_epoch=warm, iou_type=iou_type, s_feats=sfl, t_feats=t_feats)
    assert abs(loss.item() - float(g[f"feat_{name}_loss"])) <= 1e-5 * abs(float(g[f"feat_{name}_loss"]))
    np.testing.assert_allclose(items.double().numpy(), g[f"feat_{name}_items"], rtol=1e-5, atol=1e-7)
    loss.backward()
    for l, f in enumerate(sfl):
        np.testing.assert_allclose(f.grad.flatten()[:2048].double().numpy(), g[f"feat_{name}_grad_feat{l}_head"], rtol=2e-4, atol=1e-9)
        assert abs(float(f.grad.double().abs().sum()) - float(g[f"feat_{name}_grad_feat{l}_abs"])) <= 1e-4 * float(g[f"feat_{name}_grad_feat{l}_abs"])
