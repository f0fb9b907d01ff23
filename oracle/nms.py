"""oracle/nms.py -- CPU restatement of the reference post-processing (TEST INFRASTRUCTURE ONLY).

Follows `non_max_suppression` in reference yolov6/utils/nms.py:31-105 step by step in numpy
float32, and re-states the greedy suppression of `torchvision.ops.nms` (third-party, pinned only as
`torchvision>=0.9.0` in the reference's requirements.txt:5; 0.26.0 is installed here, compiled,
no source on disk).  The published CPU algorithm (torchvision/csrc/ops/cpu/nms_kernel.cpp) is:
stable sort by score descending; walk the order; a kept box i suppresses every later box j with
    inter / (area_i + area_j - inter) > iou_threshold
where inter = max(0, xx2-xx1) * max(0, yy2-yy1) in the boxes' dtype (float32), areas are
(x2-x1)*(y2-y1), and the comparison promotes the float32 ratio against the double threshold.
`tests/golden/make_golden.py` pins this restatement against the live reference + torchvision.
"""
import numpy as np

MAX_WH = np.float32(4096.0)   # nms.py:54
MAX_NMS = 30000               # nms.py:55


def greedy_nms(boxes, scores, iou_thres):
    """boxes [n,4] float32 xyxy, scores [n] float32 -> kept indices (int64) in descending-score order."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores.astype(np.float64), kind="stable")  # descending, ties -> lower index first
    sx1, sy1, sx2, sy2, sa = x1[order], y1[order], x2[order], y2[order], areas[order]
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float64(iou_thres)
    zero = np.float32(0)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 == n:
            break
        xx1 = np.maximum(sx1[i], sx1[i + 1:])
        yy1 = np.maximum(sy1[i], sy1[i + 1:])
        xx2 = np.minimum(sx2[i], sx2[i + 1:])
        yy2 = np.minimum(sy2[i], sy2[i + 1:])
        w = np.maximum(zero, xx2 - xx1)
        h = np.maximum(zero, yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (sa[i] + sa[i + 1:] - inter)
        suppressed[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def xywh2xyxy(x):
    """nms.py:21-28."""
    y = np.empty_like(x)
    half_w = x[:, 2] / np.float32(2)
    half_h = x[:, 3] / np.float32(2)
    y[:, 0] = x[:, 0] - half_w
    y[:, 1] = x[:, 1] - half_h
    y[:, 2] = x[:, 0] + half_w
    y[:, 3] = x[:, 1] + half_h
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300, return_index=False):
    """prediction: [B, A, 5+nc] float32 array.  Returns a list of B arrays [k,6] = (xyxy, conf, cls).
    With return_index=True also returns, per image, the (anchor index, class) of every kept row."""
    pred = np.asarray(prediction, dtype=np.float32)
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    nc = pred.shape[2] - 5
    conf_t = np.float32(conf_thres)
    cand = (pred[..., 4] > conf_t) & (pred[..., 5:].max(-1) > conf_t)           # nms.py:48
    multi_label = bool(multi_label) and nc > 1                                  # nms.py:57
    out, out_idx = [], []
    for b in range(pred.shape[0]):
        anchors = np.nonzero(cand[b])[0]
        x = pred[b][anchors].copy()
        empty = (np.zeros((0, 6), np.float32), np.zeros((0, 2), np.int64))
        if x.shape[0] == 0:
            out.append(empty[0]); out_idx.append(empty[1]); continue
        x[:, 5:] *= x[:, 4:5]                                                   # nms.py:69
        box = xywh2xyxy(x[:, :4])                                               # nms.py:72
        if multi_label:                                                         # nms.py:75-77
            bi, ci = np.nonzero(x[:, 5:] > conf_t)
            det = np.concatenate([box[bi], x[bi, ci + 5][:, None], ci[:, None].astype(np.float32)], 1)
            src = np.stack([anchors[bi], ci], 1)
        else:                                                                   # nms.py:79-80
            ci = x[:, 5:].argmax(1)
            conf = x[np.arange(x.shape[0]), ci + 5]
            sel = conf > conf_t
            det = np.concatenate([box, conf[:, None], ci[:, None].astype(np.float32)], 1)[sel]
            src = np.stack([anchors, ci], 1)[sel]
        if classes is not None:                                                 # nms.py:83-84
            sel = np.isin(det[:, 5], np.asarray(classes, dtype=np.float32))
            det, src = det[sel], src[sel]
        n = det.shape[0]
        if n == 0:
            out.append(empty[0]); out_idx.append(empty[1]); continue
        if n > MAX_NMS:                                                         # nms.py:90-91
            o = np.argsort(-det[:, 4].astype(np.float64), kind="stable")[:MAX_NMS]
            det, src = det[o], src[o]
        offs = det[:, 5:6] * (np.float32(0) if agnostic else MAX_WH)            # nms.py:94
        keep = greedy_nms(det[:, :4] + offs, det[:, 4], iou_thres)[:max_det]    # nms.py:95-98
        out.append(det[keep]); out_idx.append(src[keep])
    return (out, out_idx) if return_index else out
