"""GPU parity AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs 2, 3, 5) against goldens minted by the unmodified
reference (tests/golden/make_golden_configs.py): YOLOv6-S 640x640 (A = 8400, the 480-tile persistent convs and halo tile
modes at their real sizes) and YOLOv6-L6 1280x1280 (A = 34000, P6 head) forwards in both precision modes; ComputeLoss at
640x640 batch 32 with COCO-shaped targets; batched NMS at B = 32, A = 8400 with the Evaler's settings, incl. the regime
with ~390 k candidates per image that exercises max_nms = 30000 (nms.py:90-91)."""
import numpy as np
import pytest
import torch

from conftest import golden_keys, golden_npz, same_up_to_score_ties
from oracle import fabricate as fab
from oracle import loss as oloss

pytestmark = pytest.mark.gpu
MODEL_CASES = {"yolov6s": (4, 640, 16), "yolov6l6": (1, 1280, 32)}
EVAL_KW = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)


def rel_err(a, b):
    return float((np.abs(a - b) / (1.0 + np.abs(b))).max())


@pytest.mark.parametrize("name", list(MODEL_CASES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_at_benchmark_size_matches_reference(name, precision):
    from yolov6_b200.model import build_model
    B, size, step = MODEL_CASES[name]
    g = golden_npz("configs.npz")
    x = fab.synthetic_images(B, size, size, seed=40)
    assert abs(fab.checksum(x) - float(g[f"{name}_x_checksum"])) <= 1e-9 * abs(float(g[f"{name}_x_checksum"])), "RNG drift"
    m = build_model(name, 80, torch.device("cuda:0"))
    m.load_state_dict(fab.fabricate_state_dict(golden_keys(name), seed=0), strict=True)
    m.eval().set_precision(precision)
    with torch.no_grad():
        out = m(x.cuda())[0].cpu().double().numpy()
    A = out.shape[1]
    e_rows = rel_err(out[:, ::step], g[f"{name}_rows"].astype(np.float64))
    # the rows that are not stored are covered by float64 column sums: |sum error| <= tol * (A + sum |v|)
    e_sum = float((np.abs(out.sum(1) - g[f"{name}_colsum"]) / (A + g[f"{name}_abs_colsum"])).max())
    print(f"{name}@{size} {precision}: sampled rows {e_rows:.2e}, column sums {e_sum:.2e}")
    tol = 1e-4 if precision == "fp32" else 6e-2        # BASELINE.json's bar / the documented speed mode
    assert e_rows < tol and e_sum < tol


def test_compute_loss_at_640_batch_32_matches_reference():
    from yolov6_b200.assigners import expand
    from yolov6_b200.loss import ComputeLoss
    g = golden_npz("configs.npz")
    B, img, strides, nc = 32, 640, [8, 16, 32], 80
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, nc, 4, seed=60)
    targets = oloss.synthetic_targets(B, seed=61, num_classes=nc)
    chk = fab.checksum(ps) + fab.checksum(pd) + fab.checksum(targets)
    assert abs(chk - float(g["loss640_in_checksum"])) <= 1e-9 * abs(chk), "RNG drift"
    dev = torch.device("cuda:0")
    psd, pdd = ps.to(dev).requires_grad_(True), pd.to(dev).requires_grad_(True)
    feats = [torch.zeros(B, 8, h, w, device=dev) for h, w in sizes]
    cl = ComputeLoss(fpn_strides=strides, num_classes=nc, ori_img_size=img, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type="giou")
    loss, items = cl((feats, psd, pdd), targets.to(dev), 0, 1, img, img)
    loss.backward()
    c = cl.last_assignment
    assert c.G < 64, f"targets are padded to the largest per-image count, not to the batch total (G = {c.G})"
    fg = c.fg.bool().cpu().numpy()
    assert np.array_equal(np.packbits(fg), g["loss640_fg"]), "fg mask differs from the reference"
    labels, bboxes, scores, _ = expand(c, -1)
    assert np.array_equal(labels.cpu().numpy()[fg], g["loss640_labels_fg"].astype(np.int64))
    stride_col = torch.cat([torch.full((h * w,), float(s), dtype=torch.float64) for (h, w), s in zip(sizes, strides)])
    got_boxes = (bboxes.cpu() / stride_col.view(1, -1, 1)).numpy()[fg]
    np.testing.assert_allclose(got_boxes, g["loss640_bboxes_fg"].astype(np.float64), rtol=1e-6)
    sc = scores.cpu()[torch.from_numpy(fg)]
    np.testing.assert_allclose(sc.max(1).values.numpy(), g["loss640_score_fg"], rtol=1e-9)
    assert int((sc != 0).sum(1).max()) <= 1
    assert abs(loss.item() - float(g["loss640_loss"])) <= 1e-6 * abs(float(g["loss640_loss"]))
    np.testing.assert_allclose(items.cpu().numpy(), g["loss640_items"], rtol=1e-6, atol=1e-9)
    gs, gd = psd.grad.cpu(), pdd.grad.cpu()
    fgt = torch.from_numpy(fg)
    got_cls = gs[fgt].gather(1, labels.cpu()[fgt].long().unsqueeze(1)).squeeze(1).double().numpy()
    np.testing.assert_allclose(got_cls, g["loss640_grad_scores_fg_cls"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(gs.double().abs().sum(-1)[:, ::64].numpy(), g["loss640_grad_scores_rowabs"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(gd[fgt].double().numpy(), g["loss640_grad_distri_fg"], rtol=1e-4, atol=1e-7)
    for key, t in (("loss640_grad_scores_abs", gs), ("loss640_grad_distri_abs", gd)):
        assert abs(t.double().abs().sum().item() - float(g[key])) <= 1e-5 * float(g[key])
    assert float(gd[~fgt].abs().max()) == 0.0


@pytest.mark.parametrize("tag,B,gen", [("sparse", 32, "synthetic_predictions_sparse"), ("dense", 4, "synthetic_predictions")])
def test_nms_at_eval_settings_batch_matches_reference(tag, B, gen):
    """sparse: ~8000 candidates per image (the benchmark's regime); dense: ~390 000 per image, of which only the 30000
    best by confidence enter the suppression (nms.py:90-91) -- the CUDA path selects them with a score histogram."""
    from yolov6_b200.nms import non_max_suppression
    g = golden_npz("configs.npz")
    p = getattr(fab, gen)(B, 8400, 80, 70)
    assert abs(fab.checksum(p) - float(g[f"nms_{tag}_checksum"])) <= 1e-9 * abs(float(g[f"nms_{tag}_checksum"])), "RNG drift"
    out = [o.cpu().numpy() for o in non_max_suppression(p.cuda(), **EVAL_KW)]
    counts = np.array([o.shape[0] for o in out])
    assert np.array_equal(counts, g[f"nms_{tag}_counts"]), (counts.tolist(), g[f"nms_{tag}_counts"].tolist())
    off = 0
    for b, o in enumerate(out):      # dense: two kept rows of image 0 tie in confidence; the reference's argsort is unstable
        ref = g[f"nms_{tag}_rows"][off:off + counts[b]]
        assert np.array_equal(o, ref) if tag == "sparse" else same_up_to_score_ties(o, ref), f"image {b}: kept rows differ from the reference"
        off += counts[b]


def test_nms_extra_regimes_match_reference_golden():
    """> 30000 candidates (single image) and multi_label + classes + agnostic: goldens of make_golden_nms_extra.py."""
    from conftest import golden_json
    from yolov6_b200.nms import non_max_suppression
    g = golden_npz("nms_extra.npz")
    for i, (B, A, nc, seed, kw) in enumerate(golden_json("nms_extra_cases.json")):
        p = fab.synthetic_predictions(B, A, nc, seed)
        out = [o.cpu().numpy() for o in non_max_suppression(p.cuda(), **kw)]
        counts = np.array([o.shape[0] for o in out])
        assert np.array_equal(counts, g[f"c{i}_counts"]), (i, counts.tolist())
        rows = np.concatenate(out) if counts.sum() else np.zeros((0, 6), np.float32)
        if B == 1 and same_up_to_score_ties(rows, g[f"c{i}_rows"]):
            continue      # case 0 holds two kept rows with bit-identical confidence; above max_nms the reference's argsort is unstable
        bad = np.where((rows != g[f"c{i}_rows"]).any(1))[0]
        assert bad.size == 0, f"case {i}: {bad.size} kept rows differ from the reference, first at {bad[:5]}: {rows[bad[:2]]} vs {g[f'c{i}_rows'][bad[:2]]}"


def test_nms_between_sort_capacities():
    """Candidate counts between the shared-memory sort (16384 keys) and the key capacity (65536), and just above it."""
    from oracle import nms as onms
    from yolov6_b200.nms import non_max_suppression
    for A, conf in ((2000, 0.35), (2600, 0.2), (4000, 0.25)):      # ~ 37 k, ~ 69 k, ~ 93 k candidates per image
        p = fab.synthetic_predictions(2, A, 80, seed=80 + A)
        kw = dict(conf_thres=conf, iou_thres=0.6, multi_label=True, max_det=200)
        ncand = int(((p[..., 5:] * p[..., 4:5]) > conf).sum((1, 2)).max())
        out = [o.cpu().numpy() for o in non_max_suppression(p.cuda(), **kw)]
        ref = onms.non_max_suppression(p.numpy(), **kw)
        for a, b in zip(out, ref):
            assert a.shape == b.shape, f"A={A} ({ncand} candidates): {a.shape} vs {b.shape}"
            bad = np.where((a != b).any(1))[0]
            assert bad.size == 0, f"A={A} ({ncand} candidates): {bad.size} rows differ, first at {bad[:5]}: {a[bad[:2]]} vs {b[bad[:2]]}"
