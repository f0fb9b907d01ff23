"""oracle/loss_fuseab.py -- CPU restatement of the anchor-aided (fuse_ab) loss (TEST INFRASTRUCTURE ONLY).

Follows yolov6/models/losses/loss_fuseab.py:48-148 (ComputeLoss.__call__), :161-170 (VarifocalLoss), :173-229
(BboxLoss): anchors of mode 'ab' (anchor_generator.py:53-55), boxes from (x_off, y_off, w, h) around the cell centre,
TaskAlignedAssigner(topk=26) at every epoch, no DFL term for the 4-channel ab head, sums divided by target_scores_sum
whenever it is > 0.  Pinned by tests/golden/make_golden_fuseab.py (the unmodified reference, run on CPU) through
tests/test_oracle_fuseab.py.  Only tests/ may import this file.
"""
import torch
import torch.nn.functional as F

from . import assign
from .loss import iou_loss, preprocess_targets


def ab_anchors(sizes, strides, cell_offset=0.5, dtype=torch.float32, num_anchors=3):
    """generate_anchors(is_eval=False, mode='ab'), anchor_generator.py:36-63: centres in pixels and the stride column,
    each level's block repeated three times."""
    pts, strs = [], []
    for (h, w), s in zip(sizes, strides):
        sx = (torch.arange(w, dtype=dtype) + cell_offset) * s
        sy = (torch.arange(h, dtype=dtype) + cell_offset) * s
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack([xx, yy], -1).reshape(-1, 2).repeat(num_anchors, 1))
        strs.append(torch.full((h * w * num_anchors, 1), float(s), dtype=dtype))
    return torch.cat(pts), torch.cat(strs)


def compute_loss_ab(sizes, pred_scores, pred_distri, targets, *, strides, num_classes=80, ori_img_size=640, iou_type="giou",
                    loss_weight=None, return_assign=False):
    """pred_scores [B,3A,nc] post-sigmoid, pred_distri [B,3A,4] = (x_off, y_off, w, h) in stride units.
    Returns (loss, items[iou, dfl, cls]) like the reference."""
    lw = loss_weight or {"class": 1.0, "iou": 2.5, "dfl": 0.5}
    B = pred_scores.shape[0]
    anchor_points, stride_t = ab_anchors(sizes, strides, dtype=pred_scores.dtype)
    scale = torch.tensor([ori_img_size] * 4, dtype=pred_scores.dtype)                      # loss_fuseab.py:62
    t = preprocess_targets(targets, B, scale)
    gt_labels, gt_bboxes = t[:, :, :1], t[:, :, 1:]
    mask_gt = (gt_bboxes.sum(-1, keepdim=True) > 0).float()                                # :69
    anchor_points_s = anchor_points / stride_t                                             # :71
    xy = pred_distri[..., :2] + anchor_points_s                                            # :72 (in place in the reference)
    wh = pred_distri[..., 2:4]
    pred_bboxes = torch.cat([xy - wh / 2, (xy - wh / 2) + wh], -1)                          # xywh2xyxy, general.py:54-61
    with torch.no_grad():
        tl, tb, ts, fg, gi = assign.tal_assign(pred_scores.detach(), pred_bboxes.detach() * stride_t, anchor_points, gt_labels,
                                               gt_bboxes, mask_gt, topk=26, num_classes=num_classes)   # :40, :78-86
    tb = tb / stride_t                                                                     # :124
    tl = torch.where(fg > 0, tl, torch.full_like(tl, num_classes))
    one_hot = F.one_hot(tl.long(), num_classes + 1)[..., :-1]
    weight = 0.75 * pred_scores.pow(2.0) * (1 - one_hot) + ts * one_hot                    # VarifocalLoss :164-168
    loss_cls = (F.binary_cross_entropy(pred_scores.float(), ts.float(), reduction="none") * weight).sum()
    tss = ts.sum()
    if tss > 0:                                                                            # :139
        loss_cls = loss_cls / tss
    if fg.sum() > 0:                                                                       # BboxLoss :185-206
        bw = ts.sum(-1)[fg].unsqueeze(-1)
        loss_iou = (iou_loss(pred_bboxes[fg], tb[fg], iou_type) * bw).sum()
        if tss != 0:
            loss_iou = loss_iou / tss
    else:
        loss_iou = pred_distri.sum() * 0.0
    loss_dfl = pred_distri.sum() * 0.0                                                     # use_dfl False for the ab head (engine.py:299-309)
    loss = lw["class"] * loss_cls + lw["iou"] * loss_iou + lw["dfl"] * loss_dfl
    items = torch.stack([lw["iou"] * loss_iou, lw["dfl"] * loss_dfl, lw["class"] * loss_cls]).detach()
    if return_assign:
        return loss, items, dict(labels=tl, bboxes=tb, scores=ts, fg=fg, gt_idx=gi, pred_bboxes=pred_bboxes)
    return loss, items


def synthetic_ab_outputs(batch, sizes, num_classes, seed=0):
    """Seeded stand-ins for the ab head's training outputs: scores in (0, 1) with a few confident entries, offsets around
    the cell centre and positive box sizes of a few cells (stride units)."""
    g = torch.Generator().manual_seed(seed)
    A3 = 3 * sum(h * w for h, w in sizes)
    logits = torch.randn(batch, A3, num_classes, generator=g) * 1.5 - 3.0
    scores = torch.sigmoid(logits)
    xy = torch.randn(batch, A3, 2, generator=g) * 0.7
    wh = torch.rand(batch, A3, 2, generator=g) * 7.0 + 0.5
    return scores, torch.cat([xy, wh], -1)
