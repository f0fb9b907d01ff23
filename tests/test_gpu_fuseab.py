"""GPU parity of the anchor-aided (fuse_ab) loss -- yv6_ab_boxes + yv6_tal_assign(topk 26) + yv6_det_loss + yv6_ab_boxes_bwd behind
`yolov6_b200.loss_fuseab.ComputeLoss` -- against the reference's goldens (tests/golden/make_golden_fuseab.py) and the oracle,
and of the whole fuse_ab training step through the drop-in API.  (The head kernels are checked op by op in
tests/test_gpu_train.py::test_train_step_matches_reference_op_by_op[yolov6n-128-2-True].)

Bars as in test_gpu_loss.py: positives exact; loss / loss_items 1e-5 against the fp32 reference; gradients rtol 2e-4."""
import numpy as np
import pytest
import torch

from conftest import golden_json, golden_keys, golden_npz
from oracle import fabricate as fab
from oracle import loss as oloss
from oracle import loss_fuseab as oab

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_json("fuseab_cases.json"), ids=lambda c: c[0])
def test_fuseab_loss_matches_reference_golden(case):
    from yolov6_b200.loss_fuseab import ComputeLoss
    name, img, B, seed, iou_type, drop = case
    g = golden_npz("fuseab.npz")
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = oab.synthetic_ab_outputs(B, sizes, 80, seed)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
    dev = torch.device("cuda:0")
    psd, pdd = ps.to(dev).requires_grad_(True), pd.to(dev).requires_grad_(True)
    feats = [torch.zeros(B, 8, h, w, device=dev) for h, w in sizes]
    cl = ComputeLoss(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type=iou_type)
    loss, items = cl((feats, psd, pdd), targets.to(dev), 0, 1, img, img)
    loss.backward()
    ref_loss = float(g[f"{name}_loss"])
    print(name, "loss", float(loss), "reference", ref_loss)
    assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss)
    np.testing.assert_allclose(items.cpu().numpy(), g[f"{name}_items"], rtol=1e-5, atol=1e-7)
    fg = cl.last_assignment.fg.bool().cpu().numpy()
    assert np.array_equal(np.packbits(fg), g[f"{name}_fg_from_grad"]), "positives differ from the reference"
    fgt = torch.from_numpy(fg)
    np.testing.assert_allclose(pdd.grad.cpu()[fgt].double().numpy(), g[f"{name}_grad_distri_rows"], rtol=2e-4, atol=1e-7)
    assert float(pdd.grad.cpu()[~fgt].abs().max()) == 0.0
    np.testing.assert_allclose(psd.grad.cpu()[fgt].double().numpy(), g[f"{name}_grad_scores_rows"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(psd.grad.cpu().flatten()[:4096].double().numpy(), g[f"{name}_grad_scores_head"], rtol=2e-4, atol=1e-8)
    got_abs, ref_abs = float(psd.grad.double().abs().sum()), float(g[f"{name}_grad_scores_abs"])
    assert abs(got_abs - ref_abs) <= 1e-4 * ref_abs


def test_fuseab_training_step_through_the_dropin_api():
    """What Trainer.train_in_steps does with --fuse_ab (core/engine.py:161-166): both losses on the five outputs, one backward;
    every parameter -- including the two extra pred convs per level -- receives a finite gradient, eval mode ignores the branch."""
    from yolov6_b200.loss import ComputeLoss
    from yolov6_b200.loss_fuseab import ComputeLoss as ComputeLossAB
    from yolov6_b200.model import build_model
    dev = torch.device("cuda:0")
    sd = fab.fabricate_state_dict(golden_keys("yolov6n_fuseab"), seed=0)
    for k in sd:
        if (".cls_preds" in k or ".reg_preds" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    m = build_model("yolov6n", 80, dev, fuse_ab=True)
    m.load_state_dict(sd)
    m.train()
    S, B = 128, 2
    x = fab.synthetic_images(B, S, S, seed=3).to(dev)
    targets = oloss.synthetic_targets(B, seed=4).to(dev)
    crit = ComputeLoss(num_classes=80, ori_img_size=S, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type="siou")
    crit_ab = ComputeLossAB(num_classes=80, ori_img_size=S, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type="siou")
    preds, _ = m(x)
    assert len(preds) == 5 and preds[1].shape == (B, 3 * preds[3].shape[1], 80) and preds[2].shape == (B, 3 * preds[3].shape[1], 4)
    loss, items = crit((preds[0], preds[3], preds[4]), targets, 0, 0, S, S)
    loss_ab, items_ab = crit_ab(preds[:3], targets, 0, 0, S, S)
    # the ab loss on the model's own outputs equals the oracle's on the same tensors
    ref, ref_items = oab.compute_loss_ab([(S // s, S // s) for s in (8, 16, 32)], preds[1].detach().cpu(), preds[2].detach().cpu(),
                                         targets.cpu(), strides=[8, 16, 32], ori_img_size=S, iou_type="siou")
    assert abs(float(loss_ab) - float(ref)) <= 1e-5 * abs(float(ref)), (float(loss_ab), float(ref))
    (loss + loss_ab).backward()
    torch.cuda.synchronize()
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
    for i in range(3):
        assert float(m.detect.cls_preds_ab[i].weight.grad.abs().sum()) > 0 and float(m.detect.reg_preds_ab[i].bias.grad.abs().sum()) > 0
    # eval mode: the anchor-free branch only (effidehead_fuseab.py:141-199) == the same weights in a plain model
    m.eval()
    plain = build_model("yolov6n", 80, dev)
    plain.load_state_dict({k: v for k, v in m.state_dict().items() if "_ab." not in k})
    plain.eval()
    with torch.no_grad():
        assert torch.equal(m(x)[0], plain(x)[0])


def test_fuseab_train_step_graph_equals_autograd_path():
    """TrainStep (CUDA-graph segments) with both losses writes the same flat gradient as the compatible autograd path."""
    from yolov6_b200.loss import ComputeLoss
    from yolov6_b200.loss_fuseab import ComputeLoss as ComputeLossAB
    from yolov6_b200.model import build_model
    from yolov6_b200.step import TrainStep
    dev = torch.device("cuda:0")
    sd = fab.fabricate_state_dict(golden_keys("yolov6n_fuseab"), seed=0)
    for k in sd:
        if (".cls_preds" in k or ".reg_preds" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    S, B = 128, 2
    x = fab.synthetic_images(B, S, S, seed=3).to(dev)
    targets = oloss.synthetic_targets(B, seed=4).to(dev)
    kw = dict(num_classes=80, ori_img_size=S, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type="siou")
    m1 = build_model("yolov6n", 80, dev, fuse_ab=True)
    m1.load_state_dict(sd)
    m1.train()
    preds, _ = m1(x)
    l1, _ = ComputeLoss(**kw)((preds[0], preds[3], preds[4]), targets, 0, 0, S, S)
    l2, _ = ComputeLossAB(**kw)(preds[:3], targets, 0, 0, S, S)
    (l1 + l2).backward()
    ref = {n: p.grad.detach().clone() for n, p in m1.named_parameters() if p.grad is not None}
    m2 = build_model("yolov6n", 80, dev, fuse_ab=True)
    m2.load_state_dict(sd)
    step = TrainStep(m2, ComputeLoss(**kw), B, S, S, max_gt=64, compute_loss_ab=ComputeLossAB(**kw), graph=True)
    step.load(x, targets)
    out = step.run(epoch_num=0)
    torch.cuda.synchronize()
    assert abs(float(out[0]) - float(l1 + l2)) <= 1e-4 * abs(float(l1 + l2)), (float(out[0]), float(l1 + l2))
    fl = step.eng.flat
    worst = 0.0
    for n, gref in ref.items():
        if float(gref.norm()) < 1e-12:
            continue
        got = fl.grad_view(n)
        e = float((got.double() - gref.double()).norm() / gref.double().norm())
        worst = max(worst, e)
        assert e < 5e-3, f"{n}: {e:.3e}"         # same kernels; fp32 atomics make the accumulation order differ
    print("fuse_ab TrainStep vs autograd path: worst relative gradient difference", worst)
    assert any("_ab." in n for n in ref)


@pytest.mark.parametrize("case", golden_json("distill_cases.json"), ids=lambda c: c[0])
def test_distill_loss_matches_reference_golden(case):
    """`yolov6_b200.loss_distill.ComputeLoss` (yv6_det_loss with the > 0 rule + two yv6_kl_rows terms) against the reference's
    loss_distill.py goldens: loss / items 1e-5, positives exact, gradients rtol 2e-4."""
    from yolov6_b200.loss_distill import ComputeLoss
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = case
    g = golden_npz("distill.npz")
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed)
    tps, tpd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed + 100)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
    dev = torch.device("cuda:0")
    psd, pdd = ps.to(dev).requires_grad_(True), pd.to(dev).requires_grad_(True)
    feats = [torch.zeros(B, 8, h, w, device=dev) for h, w in sizes]
    cl = ComputeLoss(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=warm, use_dfl=True, reg_max=16, iou_type=iou_type,
                     distill_weight={"class": 1.0, "dfl": 1.0}, distill_feat=False)
    loss, items = cl((feats, psd, pdd), (feats, tps.to(dev), tpd.to(dev)), None, None, targets.to(dev), epoch, max_epoch, T, 1, img, img)
    loss.backward()
    ref = float(g[f"{name}_loss"])
    print(name, "loss", float(loss.detach()), "reference", ref, "items", items.tolist())
    assert abs(float(loss.detach()) - ref) <= 1e-5 * abs(ref)
    np.testing.assert_allclose(items.cpu().numpy(), g[f"{name}_items"], rtol=1e-5, atol=1e-7)
    fg = cl.last_assignment.fg.bool().cpu()
    assert np.array_equal(np.packbits(fg.numpy()), g[f"{name}_pos"]), "positives differ from the reference"
    np.testing.assert_allclose(pdd.grad.cpu()[fg].double().numpy(), g[f"{name}_grad_distri_rows"], rtol=2e-4, atol=1e-7)
    assert float(pdd.grad.cpu()[~fg].abs().max()) == 0.0
    np.testing.assert_allclose(psd.grad.cpu().flatten()[:4096].double().numpy(), g[f"{name}_grad_scores_head"], rtol=2e-4, atol=1e-7)
    got_abs, ref_abs = float(psd.grad.double().abs().sum()), float(g[f"{name}_grad_scores_abs"])
    assert abs(got_abs - ref_abs) <= 1e-4 * ref_abs


@pytest.mark.parametrize("case", golden_json("distill_cases.json"), ids=lambda c: "ns_" + c[0])
def test_distill_ns_loss_matches_reference_golden(case):
    """`yolov6_b200.loss_distill.ComputeLossNS` against the reference's loss_distill_ns.py goldens (three student tensors)."""
    from yolov6_b200.loss_distill import ComputeLossNS
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = case
    name = "ns_" + name
    g = golden_npz("distill.npz")
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed)
    _, pl = fab.synthetic_head_outputs(B, sizes, 80, 4, seed + 200)
    tps, tpd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed + 100)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
    dev = torch.device("cuda:0")
    psd, pdd, pld = (t.to(dev).requires_grad_(True) for t in (ps, pd, pl))
    feats = [torch.zeros(B, 8, h, w, device=dev) for h, w in sizes]
    cl = ComputeLossNS(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=warm, use_dfl=True, reg_max=16, iou_type=iou_type,
                       distill_weight={"class": 1.0, "dfl": 1.0}, distill_feat=False)
    loss, items = cl((feats, psd, pdd, pld), (feats, tps.to(dev), tpd.to(dev)), None, None, targets.to(dev), epoch, max_epoch, T, 1, img, img)
    loss.backward()
    ref = float(g[f"{name}_loss"])
    print(name, "loss", float(loss.detach()), "reference", ref, "items", items.tolist())
    assert abs(float(loss.detach()) - ref) <= 1e-5 * abs(ref)
    np.testing.assert_allclose(items.cpu().numpy(), g[f"{name}_items"], rtol=1e-5, atol=1e-7)
    fg = cl.last_assignment.fg.bool().cpu()
    assert np.array_equal(np.packbits(fg.numpy()), g[f"{name}_pos"]), "positives differ from the reference"
    np.testing.assert_allclose(pld.grad.cpu()[fg].double().numpy(), g[f"{name}_grad_lrtb_rows"], rtol=2e-4, atol=1e-7)
    assert float(pld.grad.cpu()[~fg].abs().max()) == 0.0
    np.testing.assert_allclose(pdd.grad.cpu()[fg].double().numpy(), g[f"{name}_grad_distri_rows"], rtol=2e-4, atol=1e-7)
    got_abs, ref_abs = float(psd.grad.double().abs().sum()), float(g[f"{name}_grad_scores_abs"])
    assert abs(got_abs - ref_abs) <= 1e-4 * ref_abs


def test_distill_ns_student_with_fuseab_teacher_through_the_dropin_api():
    """`--distill` for an N / S model (core/engine.py:153-159, 311-322, 429-441): student = build_model(distill_ns=True) on a config with
    use_dfl / reg_max 16, teacher = build_model(fuse_ab=True) of the same config, loss = loss_distill_ns.  One training step gives every
    student parameter a finite gradient; the eval forward of the student (lrtb branch, no DFL) matches the oracle within the bf16 bar."""
    from oracle import model as om
    from yolov6_b200 import configs
    from yolov6_b200.loss_distill import ComputeLossNS
    from yolov6_b200.model import build_model
    dev = torch.device("cuda:0")
    cfg = configs.get_config("yolov6n")
    cfg["head"]["use_dfl"], cfg["head"]["reg_max"] = True, 16
    keys = golden_keys("yolov6n_distill_ns")
    sd = fab.fabricate_state_dict(keys, seed=0)
    for k in sd:
        if (".cls_preds" in k or ".reg_preds" in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    student = build_model(cfg, 80, dev, distill_ns=True)
    student.load_state_dict(sd)
    teacher = build_model(cfg, 80, dev, fuse_ab=True)          # random init is enough here: its outputs are just inputs of the loss
    S, B = 128, 2
    x = fab.synthetic_images(B, S, S, seed=3).to(dev)
    targets = oloss.synthetic_targets(B, seed=4).to(dev)
    student.train()
    teacher.train()
    preds, s_feats = student(x)
    with torch.no_grad():
        t_preds, t_feats = teacher(x)
    assert len(preds) == 4 and preds[2].shape[2] == 68 and preds[3].shape[2] == 4 and len(t_preds) == 5 and t_preds[-1].shape[2] == 68
    crit = ComputeLossNS(num_classes=80, ori_img_size=S, warmup_epoch=0, use_dfl=True, reg_max=16, iou_type="siou",
                         distill_weight={"class": 1.0, "dfl": 1.0}, distill_feat=False)
    loss, items = crit(preds, t_preds, s_feats, t_feats, targets, 10, 300, 20.0, 0, S, S)
    assert items.shape == (4,) and bool(torch.isfinite(loss))
    loss.backward()
    torch.cuda.synchronize()
    for n, p in student.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
    assert float(student.detect.reg_preds_dist[1].weight.grad.abs().sum()) > 0 and float(student.detect.reg_preds[1].weight.grad.abs().sum()) > 0
    student.eval().set_precision("fp32")
    with torch.no_grad():
        got = student(x)[0].cpu().double()
        cur = {k: v.detach().cpu() for k, v in student.state_dict().items()}    # the training forward moved the BatchNorm running statistics
        ref = om.forward({k: v.double() if v.is_floating_point() else v for k, v in cur.items()}, dict(om.CONFIGS["yolov6n"], use_dfl=True, reg_max=16),
                         x.cpu().double(), distill_ns=True)
    err = float(((got - ref).abs() / (1 + ref.abs())).max())
    print("distill_ns eval (fp32-equivalent mode) vs oracle:", err)
    assert got.shape == ref.shape and err < 1e-4


def test_distill_feature_term_matches_reference_golden():
    """distill_feat=True on the GPU path: loss / items 1e-5 and the gradient w.r.t. the student's feature maps (yv6_kl_rows over
    (image, channel) rows of H*W positions) against the reference golden."""
    from yolov6_b200.loss_distill import ComputeLoss
    case = golden_json("distill_cases.json")[0]
    name, img, B, seed, iou_type, warm, epoch, max_epoch, T, drop = case
    g = golden_npz("distill.npz")
    strides = [8, 16, 32]
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed)
    tps, tpd = fab.synthetic_head_outputs(B, sizes, 80, 68, seed + 100)
    targets = oloss.drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=80), drop)
    gf = torch.Generator().manual_seed(seed + 300)
    s_feats = [torch.randn(B, c, h, w, generator=gf) for c, (h, w) in zip((32, 64, 128), sizes)]
    t_feats = [torch.randn(B, c, h, w, generator=gf) * 1.3 for c, (h, w) in zip((32, 64, 128), sizes)]
    dev = torch.device("cuda:0")
    psd, pdd = ps.to(dev).requires_grad_(True), pd.to(dev).requires_grad_(True)
    sfd = [f.to(dev).requires_grad_(True) for f in s_feats]
    feats = [torch.zeros(B, 8, h, w, device=dev) for h, w in sizes]
    cl = ComputeLoss(fpn_strides=strides, num_classes=80, ori_img_size=img, warmup_epoch=warm, use_dfl=True, reg_max=16, iou_type=iou_type,
                     distill_weight={"class": 1.0, "dfl": 1.0}, distill_feat=True)
    loss, items = cl((feats, psd, pdd), (feats, tps.to(dev), tpd.to(dev)), sfd, [f.to(dev) for f in t_feats], targets.to(dev), epoch, max_epoch,
                     T, 1, img, img)
    loss.backward()
    ref = float(g[f"feat_{name}_loss"])
    print("feat loss", float(loss.detach()), "reference", ref, items.tolist())
    assert abs(float(loss.detach()) - ref) <= 1e-5 * abs(ref)
    np.testing.assert_allclose(items.cpu().numpy(), g[f"feat_{name}_items"], rtol=1e-5, atol=1e-7)
    for l, f in enumerate(sfd):
        np.testing.assert_allclose(f.grad.cpu().flatten()[:2048].double().numpy(), g[f"feat_{name}_grad_feat{l}_head"], rtol=2e-4, atol=1e-9)
        assert abs(float(f.grad.double().abs().sum()) - float(g[f"feat_{name}_grad_feat{l}_abs"])) <= 1e-4 * float(g[f"feat_{name}_grad_feat{l}_abs"])


def test_feature_maps_are_differentiable_outputs_of_the_training_forward():
    """model.return_featmaps = True: the training forward returns the real neck outputs (yolo.py:37-39) and a loss on them reaches
    the backbone / neck parameters; the head-only gradients are unchanged by the extra outputs."""
    from yolov6_b200.model import build_model
    dev = torch.device("cuda:0")
    sd = fab.fabricate_state_dict(golden_keys("yolov6n"), seed=0)
    for k in sd:
        if (".cls_preds." in k or ".reg_preds." in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    x = fab.synthetic_images(2, 128, 128, seed=3).to(dev)
    gen = torch.Generator().manual_seed(9)

    def run(with_feats, feat_loss):
        m = build_model("yolov6n", 80, dev)
        m.load_state_dict(sd)
        m.train()
        m.return_featmaps = with_feats
        (feats, cls, reg), fmaps = m(x)
        L = (cls * wc).sum() + (reg * wr).sum()
        if feat_loss:
            L = L + sum((f * w).sum() for f, w in zip(fmaps, wf))
        L.backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}, fmaps

    m0 = build_model("yolov6n", 80, dev)
    m0.load_state_dict(sd)
    m0.train()
    (f0, c0, r0), _ = m0(x)
    wc, wr = torch.randn(c0.shape, generator=gen).to(dev), torch.randn(r0.shape, generator=gen).to(dev)
    sizes = [tuple(f.shape[2:]) for f in f0]
    chans = [t.c for t in m0.graph.feat]
    wf = [torch.randn(2, c, h, w, generator=gen).to(dev) for c, (h, w) in zip(chans, sizes)]
    base, _ = run(False, False)
    same, fm = run(True, False)
    assert [tuple(f.shape) for f in fm] == [(2, c, h, w) for c, (h, w) in zip(chans, sizes)] and all(f.requires_grad for f in fm)
    worst = max(float((same[n] - base[n]).norm() / (base[n].norm() + 1e-30)) for n in base)
    print("head-only gradients with / without feature-map outputs: worst rel diff", worst)
    assert worst < 5e-3
    both, _ = run(True, True)
    moved = float((both["neck.Rep_n4.conv1.rbr_dense.conv.weight"] - base["neck.Rep_n4.conv1.rbr_dense.conv.weight"]).norm())
    head_same = float((both["detect.cls_preds.0.weight"] - base["detect.cls_preds.0.weight"]).norm() / base["detect.cls_preds.0.weight"].norm())
    assert moved > 0 and head_same < 5e-3          # the feature loss reaches the neck, the preds see only the head loss
