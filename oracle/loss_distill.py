"""oracle/loss_distill.py -- CPU restatement of the self-distillation loss of the M / L models (TEST INFRASTRUCTURE ONLY).

Follows yolov6/models/losses/loss_distill.py:59-211 (ComputeLoss.__call__), :213-222 (distill_loss_cls), :287-361 (BboxLoss with
distill_loss_dfl): the anchor-free detection loss with the "> 0" normalisation rule plus two temperature-scaled KL terms against
a teacher's head outputs, both multiplied by a cosine decay.  The channel-wise feature term (`distill_feat`) is not restated
(the product does not build it).  Pinned by tests/golden/make_golden_distill.py through tests/test_oracle_distill.py.
"""
import math

import torch
import torch.nn.functional as F

from . import loss as oloss


def kl_rows(student, teacher, temperature):
    """distill_loss_cls / distill_loss_dfl core: sum over rows of KL(softmax(t/T) || softmax(s/T)), per row."""
    ps = F.softmax(student / temperature, dim=1)
    pt = F.softmax(teacher / temperature, dim=1)
    return F.kl_div(torch.log(ps), pt, reduction="none").sum(1)


def cw_loss(s_feats, t_feats, temperature=1.0):
    """distill_loss_cw, loss_distill.py:223-245: per level, KL over the H*W positions of every (image, channel) row, / (N*C)."""
    total = 0.0
    for sf, tf in zip(s_feats[:3], t_feats[:3]):
        N, C, H, W = sf.shape
        total = total + F.kl_div(F.log_softmax(sf.reshape(N, C, H * W) / temperature, dim=2),
                                 F.log_softmax(tf.reshape(N, C, H * W).detach() / temperature, dim=2), reduction="sum",
                                 log_target=True) * (temperature * temperature) / (N * C)
    return total


def compute_loss_distill(sizes, pred_scores, pred_distri, t_pred_scores, t_pred_distri, targets, *, strides, epoch_num, max_epoch,
                         temperature, num_classes=80, ori_img_size=640, warmup_epoch=0, use_dfl=True, reg_max=16, iou_type="giou",
                         loss_weight=None, distill_weight=None, s_feats=None, t_feats=None):
    """Returns (loss, items[iou, dfl_all, cls_all, cwd]) like the reference; the feature term is on when s_feats / t_feats are given."""
    lw = loss_weight or {"class": 1.0, "iou": 2.5, "dfl": 0.5, "cwd": 10.0}
    dw = distill_weight or {"class": 1.0, "dfl": 1.0}
    base, items, a = oloss.compute_loss(sizes, pred_scores, pred_distri, targets, strides=strides, num_classes=num_classes,
                                        ori_img_size=ori_img_size, warmup_epoch=warmup_epoch, epoch_num=epoch_num, use_dfl=use_dfl,
                                        reg_max=reg_max, iou_type=iou_type, loss_weight=lw, return_assign=True, norm_gt_zero=True)
    T = float(temperature)
    d_cls = kl_rows(pred_scores.reshape(-1, num_classes), t_pred_scores.reshape(-1, num_classes), T).sum() * T * T       # :213-222
    fg, ts = a["fg"], a["scores"]
    tss = ts.sum()
    if use_dfl and fg.sum() > 0:                                                                                       # :306-323
        R = reg_max + 1
        s_pos = pred_distri[fg].reshape(-1, R)
        t_pos = t_pred_distri[fg].reshape(-1, R)
        d = kl_rows(s_pos, t_pos, T).mean() * T * T                                                                     # :351-361
        bw = ts.sum(-1)[fg].unsqueeze(-1)
        d_dfl = (d * bw).sum()
        if tss != 0:
            d_dfl = d_dfl / tss
    else:
        d_dfl = pred_distri.sum() * 0.0
    decay = ((1 - math.cos(epoch_num * math.pi / max_epoch)) / 2) * (0.01 - 1) + 1                                       # :196
    d_cls, d_dfl = d_cls * decay, d_dfl * decay
    d_cw = cw_loss(s_feats, t_feats) * decay if s_feats is not None else torch.zeros(())
    loss = base + lw["class"] * dw["class"] * d_cls + lw["dfl"] * dw["dfl"] * d_dfl + lw["cwd"] * d_cw
    items4 = torch.stack([items[0], items[1] + (lw["dfl"] * dw["dfl"] * d_dfl).detach(), items[2] + (lw["class"] * dw["class"] * d_cls).detach(),
                          (lw["cwd"] * d_cw).detach().to(items.dtype) if torch.is_tensor(d_cw) else torch.zeros(())]).detach()
    return loss, items4


def compute_loss_distill_ns(sizes, pred_scores, pred_distri, pred_lrtb, t_pred_scores, t_pred_distri, targets, *, strides, epoch_num,
                            max_epoch, temperature, num_classes=80, ori_img_size=640, reg_max=16, iou_type="giou", loss_weight=None,
                            distill_weight=None):
    """yolov6/models/losses/loss_distill_ns.py:58-211: the M / L distillation loss on (scores, DFL distributions) with the
    TaskAlignedAssigner at every epoch, plus the IoU loss of the lrtb branch's boxes against the same assignment (:93, :283-292)."""
    from . import assign
    lw = loss_weight or {"class": 1.0, "iou": 2.5, "dfl": 0.5, "cwd": 10.0}
    loss, items = compute_loss_distill(sizes, pred_scores, pred_distri, t_pred_scores, t_pred_distri, targets, strides=strides,
                                       epoch_num=epoch_num, max_epoch=max_epoch, temperature=temperature, num_classes=num_classes,
                                       ori_img_size=ori_img_size, warmup_epoch=0, use_dfl=True, reg_max=reg_max, iou_type=iou_type,
                                       loss_weight=lw, distill_weight=distill_weight)
    _, _, a = oloss.compute_loss(sizes, pred_scores.detach(), pred_distri.detach(), targets, strides=strides, num_classes=num_classes,
                                 ori_img_size=ori_img_size, warmup_epoch=0, epoch_num=epoch_num, use_dfl=True, reg_max=reg_max,
                                 iou_type=iou_type, loss_weight=lw, return_assign=True, norm_gt_zero=True)
    fg, ts, tb = a["fg"], a["scores"], a["bboxes"]
    _, anchor_points, _, stride_t = assign.train_anchors(sizes, strides, dtype=pred_scores.dtype)
    aps = anchor_points / stride_t
    lt, rb = torch.split(pred_lrtb, 2, -1)                                               # dist2bbox(xyxy), general.py:32-38
    boxes = torch.cat([aps - lt, aps + rb], -1)
    tss = ts.sum()
    if fg.sum() > 0:
        bw = ts.sum(-1)[fg].unsqueeze(-1)
        iou_lrtb = (oloss.iou_loss(boxes[fg], tb[fg], iou_type) * bw).sum()
        if tss != 0:
            iou_lrtb = iou_lrtb / tss
    else:
        iou_lrtb = pred_lrtb.sum() * 0.0
    loss = loss + lw["iou"] * iou_lrtb
    items = items.clone()
    items[0] = items[0] + (lw["iou"] * iou_lrtb).detach()
    return loss, items
