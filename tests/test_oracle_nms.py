"""Pins oracle/nms.py (incl. the restated torchvision greedy NMS) bit-exactly against the reference's
non_max_suppression outputs stored by tests/golden/make_golden.py."""
import numpy as np
import pytest

from conftest import golden_json, golden_npz, same_up_to_score_ties
from oracle import fabricate as fab
from oracle import nms as onms


def test_nms_oracle_bit_exact_vs_reference():
    g = golden_npz("nms.npz")
    total = 0
    for i, (B, A, nc, seed, kw) in enumerate(golden_json("nms_cases.json")):
        p = fab.synthetic_predictions(B, A, nc, seed)
        assert abs(fab.checksum(p) - float(g[f"c{i}_checksum"])) <= 1e-9 * abs(float(g[f"c{i}_checksum"])), "RNG drift"
        out = onms.non_max_suppression(p.numpy(), **kw)
        counts = np.array([o.shape[0] for o in out])
        assert np.array_equal(counts, g[f"c{i}_counts"]), (i, counts, g[f"c{i}_counts"])
        rows = np.concatenate(out) if counts.sum() else np.zeros((0, 6), np.float32)
        assert np.array_equal(rows, g[f"c{i}_rows"]), f"case {i}: kept rows differ"
        total += int(counts.sum())
    assert total > 3000


def test_greedy_nms_semantics():
    # ties keep the lower index; IoU == thr is kept (strict >); float IoU vs double threshold
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [20, 20, 30, 30]], np.float32)
    scores = np.array([0.5, 0.5, 0.5, 0.5], np.float32)
    assert onms.greedy_nms(boxes, scores, 0.5).tolist() == [0, 2]
    b = np.array([[0, 0, 2, 1], [1, 0, 3, 1]], np.float32)  # IoU = 1/3 (float32) > 1/3 (double) -> suppressed
    assert onms.greedy_nms(b, np.array([0.9, 0.8], np.float32), 1.0 / 3.0).tolist() == [0]
    assert onms.greedy_nms(b, np.array([0.9, 0.8], np.float32), 0.34).tolist() == [0, 1]
    assert onms.greedy_nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).shape == (0,)


def test_nms_empty_and_class_offset():
    p = fab.synthetic_predictions(2, 50, 4, seed=9).numpy()
    assert all(o.shape == (0, 6) for o in onms.non_max_suppression(p, conf_thres=1.0))
    # identical boxes of different classes survive unless agnostic
    q = np.zeros((1, 2, 9), np.float32)
    q[0, :, :4] = [100, 100, 50, 50]
    q[0, :, 4] = 1
    q[0, 0, 5] = 0.9
    q[0, 1, 7] = 0.8
    assert onms.non_max_suppression(q, 0.25, 0.45)[0].shape[0] == 2
    assert onms.non_max_suppression(q, 0.25, 0.45, agnostic=True)[0].shape[0] == 1


def test_nms_oracle_extra_regimes_vs_reference():
    """More than max_nms = 30000 candidates in one image (only the 30000 best enter torchvision.ops.nms, nms.py:90-91) and
    multi_label + class filter + agnostic together; goldens from tests/golden/make_golden_nms_extra.py."""
    g = golden_npz("nms_extra.npz")
    for i, (B, A, nc, seed, kw) in enumerate(golden_json("nms_extra_cases.json")):
        p = fab.synthetic_predictions(B, A, nc, seed)
        assert abs(fab.checksum(p) - float(g[f"c{i}_checksum"])) <= 1e-9 * abs(float(g[f"c{i}_checksum"])), "RNG drift"
        out = onms.non_max_suppression(p.numpy(), **kw)
        counts = np.array([o.shape[0] for o in out])
        assert np.array_equal(counts, g[f"c{i}_counts"]), (i, counts, g[f"c{i}_counts"])
        # The reference pre-sorts the > 30000 candidates with an UNSTABLE argsort (nms.py:90-91), so rows with bit-identical
        # confidence may come out in either order; compare after ordering each tie group canonically (SURVEY A.3 caveat).
        def canon(r):
            return r[np.lexsort((r[:, 1], r[:, 0], r[:, 5], -r[:, 4]))]
        got, ref = np.concatenate(out), g[f"c{i}_rows"]
        assert np.array_equal(got[:, 4], ref[:, 4]), f"extra case {i}: confidences differ"
        assert np.array_equal(canon(got), canon(ref)), f"extra case {i}: kept rows differ"


def test_oracle_matches_reference_at_eval_batch_settings():
    """B = 32 / A = 8400 Evaler settings incl. the ~390 k-candidate regime (tests/golden/make_golden_configs.py)."""
    g = golden_npz("configs.npz")
    kw = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
    for tag, B, gen in (("sparse", 4, fab.synthetic_predictions_sparse), ("dense", 1, fab.synthetic_predictions)):
        full_B = 32 if tag == "sparse" else 4
        p = gen(full_B, 8400, 80, 70)[:B]           # the oracle is slow: first images only
        out = onms.non_max_suppression(p.numpy(), **kw)
        counts = g[f"nms_{tag}_counts"]
        rows = g[f"nms_{tag}_rows"]
        off = 0
        for b in range(B):
            assert out[b].shape[0] == counts[b]
            assert same_up_to_score_ties(out[b], rows[off:off + counts[b]])
            off += counts[b]


def _iou(a, b):
    x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
    x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa, ab = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]), (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter)


@pytest.mark.parametrize("seed,kw", [(3, dict(conf_thres=0.25, iou_thres=0.45)),
                                     (4, dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)),
                                     (5, dict(conf_thres=0.1, iou_thres=0.5, agnostic=True, max_det=20)),
                                     (6, dict(conf_thres=0.2, iou_thres=0.6, classes=[1, 3]))])
def test_nms_size_independent_properties(seed, kw):
    """Properties every output of nms.py:31-105 has whatever the input size (the GPU tests check the same at A = 8400, B = 32
    through bit-exactness with this oracle): rows sorted by score, at most max_det of them, all above conf_thres, kept boxes of a
    class (of any class when agnostic) never overlap by more than iou_thres, class filter honoured, and the function is
    idempotent on its own output."""
    p = fab.synthetic_predictions(3, 400, 6, seed=seed).numpy()
    outs = onms.non_max_suppression(p, **kw)
    assert len(outs) == 3 and any(o.shape[0] for o in outs)
    for o in outs:
        assert o.shape[1] == 6 and o.shape[0] <= kw.get("max_det", 300)
        if not o.shape[0]:
            continue
        assert np.all(np.diff(o[:, 4]) <= 0), "rows are sorted by score"
        assert np.all(o[:, 4] > kw["conf_thres"])
        if "classes" in kw:
            assert set(o[:, 5].astype(int)) <= set(kw["classes"])
        iou = _iou(o[:, :4], o[:, :4])
        same = np.ones_like(iou, bool) if kw.get("agnostic") else (o[:, 5][:, None] == o[:, 5][None, :])
        np.fill_diagonal(same, False)
        assert not np.any((iou > kw["iou_thres"] + 1e-6) & same), "a kept box overlaps a better kept box of its class"
        # idempotence: the kept detections, fed back as predictions (obj = 1, one-hot class score), all survive in the same order
        q = np.zeros((1, o.shape[0], 5 + 6), np.float32)
        q[0, :, 0], q[0, :, 1] = (o[:, 0] + o[:, 2]) / 2, (o[:, 1] + o[:, 3]) / 2
        q[0, :, 2], q[0, :, 3] = o[:, 2] - o[:, 0], o[:, 3] - o[:, 1]
        q[0, :, 4] = 1.0
        q[0, np.arange(o.shape[0]), 5 + o[:, 5].astype(int)] = o[:, 4]
        again = onms.non_max_suppression(q, **kw)[0]
        assert again.shape == o.shape
        np.testing.assert_allclose(again[:, 4:], o[:, 4:], rtol=0, atol=0)
        np.testing.assert_allclose(again[:, :4], o[:, :4], rtol=0, atol=1e-3)
