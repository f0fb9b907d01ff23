"""The oracle's anchor-aided (fuse_ab) head and loss against the reference's goldens (tests/golden/make_golden_fuseab.py:
heads/effidehead_fuseab.py + losses/loss_fuseab.py of the unmodified reference, run on CPU)."""
import numpy as np
import pytest
import torch

from conftest import golden_json, golden_npz
from oracle import fabricate as fab
from oracle import loss as oloss
from oracle import loss_fuseab as oab
from oracle import model as om


def rel_err(a, b):
    return float((np.abs(a - b) / (1.0 + np.abs(b))).max())


def fuseab_state_dict():
    keys = [(k, tuple(s)) for k, s in golden_json("keys_yolov6n_fuseab.json")]
    sd = fab.fabricate_state_dict(keys, seed=0)
    for k in sd:
        if (".cls_preds" in k or ".reg_preds" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    return sd


def test_fuseab_head_train_mode_matches_reference():
    """Five training outputs of the fuse_ab model, L = sum(cls_ab w1) + sum(reg_ab w2), the gradient norm of every
    parameter that L reaches and selected full gradients (float64)."""
    g = golden_npz("fuseab.npz")
    sd = fuseab_state_dict()
    x = fab.synthetic_images(2, 64, 64, seed=7)
    assert abs(fab.checksum(x) - float(g["m_x_checksum"])) < 1e-6 * abs(float(g["m_x_checksum"])), "input RNG drift"
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    with om.train_mode():
        cls, reg, _, cls_ab, reg_ab = om.forward(sd64, om.CONFIGS["yolov6n"], x.double(), train_outputs=True, fuse_ab=True)
    for got, key in ((cls, "m_cls"), (reg, "m_reg"), (cls_ab, "m_cls_ab"), (reg_ab, "m_reg_ab")):
        assert got.shape == g[key].shape, key
        assert rel_err(got.detach().numpy(), g[key]) < 1e-9, key
    gen = torch.Generator().manual_seed(13)
    w1 = torch.randn(cls_ab.shape, generator=gen).double()
    w2 = torch.randn(reg_ab.shape, generator=gen).double()
    L = (cls_ab * w1).sum() + (reg_ab * w2).sum()
    assert abs(L.item() - float(g["m_L"])) < 1e-8 * max(1.0, abs(float(g["m_L"])))
    L.backward()
    names, norms = [str(n) for n in g["m_grad_names"]], g["m_grad_norms"]
    assert len(names) == 370
    for n, ref in zip(names, norms):
        assert sd64[n].grad is not None, n
        assert abs(float(sd64[n].grad.norm()) - ref) <= 1e-7 * max(1.0, ref), n
    for k in g.files:
        if k.startswith("m_grad::"):
            n = k[8:]
            np.testing.assert_allclose(sd64[n].grad.numpy().reshape(g[k].shape), g[k], rtol=1e-7, atol=1e-9 * (1 + np.abs(g[k]).max()))


@pytest.mark.parametrize("case", golden_json("fuseab_cases.json"), ids=lambda c: c[0])
def test_fuseab_loss_matches_reference(case):
    """loss_fuseab.ComputeLoss: loss, loss_items (fp32 reference: 1e-5), the positives and the gradients w.r.t. both inputs."""
    name, img, B, seed, iou_type, drop = case
    g = golden_npz("fuseab.npz")
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = oab.synthetic_ab_outputs(B, sizes, 80, seed)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
    chk = fab.checksum(ps) + fab.checksum(pd) + fab.checksum(targets)
    assert abs(chk - float(g[f"{name}_in_checksum"])) < 1e-6 * abs(chk), "input RNG drift"
    psl, pdl = ps.clone().requires_grad_(True), pd.clone().requires_grad_(True)
    loss, items, a = oab.compute_loss_ab(sizes, psl, pdl, targets, strides=strides, ori_img_size=img, iou_type=iou_type, return_assign=True)
    assert abs(loss.item() - float(g[f"{name}_loss"])) <= 1e-5 * abs(float(g[f"{name}_loss"]))
    np.testing.assert_allclose(items.double().numpy(), g[f"{name}_items"], rtol=1e-5, atol=1e-7)
    loss.backward()
    nz = (pdl.grad.abs().sum(-1) > 0).numpy()
    assert np.array_equal(np.packbits(nz), g[f"{name}_fg_from_grad"]), "positives differ from the reference"
    assert np.array_equal(nz, a["fg"].numpy())
    np.testing.assert_allclose(pdl.grad[torch.from_numpy(nz)].double().numpy(), g[f"{name}_grad_distri_rows"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(psl.grad[torch.from_numpy(nz)].double().numpy(), g[f"{name}_grad_scores_rows"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(psl.grad.flatten()[:4096].double().numpy(), g[f"{name}_grad_scores_head"], rtol=2e-4, atol=1e-8)
    assert abs(float(psl.grad.double().abs().sum()) - float(g[f"{name}_grad_scores_abs"])) <= 1e-4 * float(g[f"{name}_grad_scores_abs"])


def test_fuseab_graph_matches_reference_state_dict_layout():
    """Model(fuse_ab=True) exposes exactly the reference's state_dict keys and shapes (yolo.py:122-126, effidehead_fuseab.py:44-55)."""
    from yolov6_b200.model import Model
    m = Model("yolov6n", num_classes=80, fuse_ab=True)
    have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = {k: tuple(s) for k, s in golden_json("keys_yolov6n_fuseab.json")}
    assert have == want
    assert tuple(m.detect.anchors_init.shape) == (3, 3, 2) and m.detect.na == 3


def test_distill_ns_head_matches_reference():
    """N / S distillation student (heads/effidehead_distill_ns.py): training outputs (cls, DFL logits, lrtb distances), gradients,
    the eval-mode prediction (lrtb branch, no DFL) and the state_dict layout of Model(distill_ns=True)."""
    from yolov6_b200 import configs
    from yolov6_b200.model import Model
    g = golden_npz("fuseab.npz")
    keys = [(k, tuple(s)) for k, s in golden_json("keys_yolov6n_distill_ns.json")]
    cfg = configs.get_config("yolov6n")
    cfg["head"]["use_dfl"], cfg["head"]["reg_max"] = True, 16
    m = Model(cfg, num_classes=80, distill_ns=True)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(keys)
    sd = fab.fabricate_state_dict(keys, seed=0)
    for k in sd:
        if (".cls_preds" in k or ".reg_preds" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    x = fab.synthetic_images(2, 64, 64, seed=7)
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    ocfg = dict(om.CONFIGS["yolov6n"], use_dfl=True, reg_max=16)
    with om.train_mode():
        cls, reg, _, reg_dist = om.forward(sd64, ocfg, x.double(), train_outputs=True, distill_ns=True)
    for got, key in ((cls, "ns_cls"), (reg, "ns_reg_lrtb"), (reg_dist, "ns_reg_dist")):
        assert got.shape == g[key].shape and rel_err(got.detach().numpy(), g[key]) < 1e-9, key
    gen = torch.Generator().manual_seed(17)
    w0, w1, w2 = (torch.randn(t.shape, generator=gen).double() for t in (cls, reg_dist, reg))
    L = (cls * w0).sum() + (reg_dist * w1).sum() + (reg * w2).sum()
    assert abs(L.item() - float(g["ns_L"])) < 1e-8 * max(1.0, abs(float(g["ns_L"])))
    L.backward()
    for n, ref in zip([str(n) for n in g["ns_grad_names"]], g["ns_grad_norms"]):
        assert abs(float(sd64[n].grad.norm()) - ref) <= 1e-7 * max(1.0, ref), n
    for k in g.files:
        if k.startswith("ns_grad::"):
            np.testing.assert_allclose(sd64[k[9:]].grad.numpy().reshape(g[k].shape), g[k], rtol=1e-7, atol=1e-9 * (1 + np.abs(g[k]).max()))
    with torch.no_grad():
        ev = om.forward({k: v.detach() for k, v in sd64.items()}, ocfg, x.double(), distill_ns=True)
    assert rel_err(ev.numpy(), g["ns_eval"]) < 1e-9
