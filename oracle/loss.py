"""oracle/loss.py -- CPU restatement of the detection loss (TEST INFRASTRUCTURE ONLY).

Re-states ComputeLoss.__call__ (reference yolov6/models/losses/loss.py:52-182) with its pieces:
preprocess (:184-192), bbox_decode (:194-198), VarifocalLoss (:201-211), BboxLoss (:214-278),
IOUloss giou/siou/ciou/diou (yolov6/utils/figure_iou.py:23-100), dist2bbox / bbox2dist
(yolov6/utils/general.py:32-52).  Differentiable (plain torch autograd on CPU) so the CUDA
forward+backward kernel can be checked on loss value AND gradients.  float64 wherever the
reference promotes to float64.  Pinned against the live reference by
tests/golden/make_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import assign


def preprocess_targets(targets, batch_size, scale_wh):
    """loss.py:184-192: ragged [n,6] (img, cls, cx, cy, w, h in [0,1]) -> padded [B,G,5] float64
    (cls, x1, y1, x2, y2 in pixels); pad rows are [-1, 0, 0, 0, 0]."""
    rows = [[] for _ in range(batch_size)]
    for item in targets.detach().cpu().numpy().tolist():
        rows[int(item[0])].append(item[1:])
    G = max((len(r) for r in rows), default=0)
    out = np.zeros((batch_size, G, 5), dtype=np.float64)
    out[:, :, 0] = -1.0
    for b, r in enumerate(rows):
        if r:
            out[b, :len(r)] = np.asarray(r, dtype=np.float64)
    t = torch.from_numpy(out)
    if G == 0:
        return t
    xywh = t[:, :, 1:5] * scale_wh.to(torch.float64)
    x1 = xywh[..., 0] - xywh[..., 2] * 0.5                 # general.py:55-61 (in-place order preserved)
    y1 = xywh[..., 1] - xywh[..., 3] * 0.5
    x2 = x1 + xywh[..., 2]
    y2 = y1 + xywh[..., 3]
    t[:, :, 1:5] = torch.stack([x1, y1, x2, y2], -1)
    return t


def iou_loss(b1, b2, iou_type="giou", eps=1e-10):
    """IOUloss(box_format='xyxy', eps=1e-10), figure_iou.py:23-100; b1, b2: [P,4] -> [P,1]."""
    b1_x1, b1_y1, b1_x2, b1_y2 = torch.split(b1, 1, dim=-1)
    b2_x1, b2_y1, b2_x2, b2_y2 = torch.split(b2, 1, dim=-1)
    inter = (torch.min(b1_x2, b2_x2) - torch.max(b1_x1, b2_x1)).clamp(0) * \
            (torch.min(b1_y2, b2_y2) - torch.max(b1_y1, b2_y1)).clamp(0)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.max(b1_x2, b2_x2) - torch.min(b1_x1, b2_x1)
    ch = torch.max(b1_y2, b2_y2) - torch.min(b1_y1, b2_y1)
    if iou_type == "giou":
        c_area = cw * ch + eps
        iou = iou - (c_area - union) / c_area
    elif iou_type in ("diou", "ciou"):
        c2 = cw ** 2 + ch ** 2 + eps
        rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2) ** 2 + (b2_y1 + b2_y2 - b1_y1 - b1_y2) ** 2) / 4
        if iou_type == "diou":
            iou = iou - rho2 / c2
        else:
            v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
            with torch.no_grad():
                alpha = v / (v - iou + (1 + eps))
            iou = iou - (rho2 / c2 + v * alpha)
    elif iou_type == "siou":
        s_cw = (b2_x1 + b2_x2 - b1_x1 - b1_x2) * 0.5 + eps
        s_ch = (b2_y1 + b2_y2 - b1_y1 - b1_y2) * 0.5 + eps
        sigma = torch.pow(s_cw ** 2 + s_ch ** 2, 0.5)
        sin_a1 = torch.abs(s_cw) / sigma
        sin_a2 = torch.abs(s_ch) / sigma
        sin_a = torch.where(sin_a1 > pow(2, 0.5) / 2, sin_a2, sin_a1)
        angle_cost = torch.cos(torch.arcsin(sin_a) * 2 - math.pi / 2)
        rho_x = (s_cw / cw) ** 2
        rho_y = (s_ch / ch) ** 2
        gamma = angle_cost - 2
        distance_cost = 2 - torch.exp(gamma * rho_x) - torch.exp(gamma * rho_y)
        omiga_w = torch.abs(w1 - w2) / torch.max(w1, w2)
        omiga_h = torch.abs(h1 - h2) / torch.max(h1, h2)
        shape_cost = torch.pow(1 - torch.exp(-1 * omiga_w), 4) + torch.pow(1 - torch.exp(-1 * omiga_h), 4)
        iou = iou - 0.5 * (distance_cost + shape_cost)
    else:
        raise ValueError(iou_type)
    return 1.0 - iou


def decode_pred(pred_distri, anchor_points_s, use_dfl, reg_max):
    """bbox_decode, loss.py:194-198 + dist2bbox(xyxy), general.py:32-38."""
    if use_dfl:
        B, A, _ = pred_distri.shape
        proj = torch.linspace(0, reg_max, reg_max + 1)
        pred_distri = F.softmax(pred_distri.view(B, A, 4, reg_max + 1), dim=-1).matmul(proj)
    lt, rb = torch.split(pred_distri, 2, -1)
    return torch.cat([anchor_points_s - lt, anchor_points_s + rb], -1)


def compute_loss(sizes, pred_scores, pred_distri, targets, *, strides, num_classes=80, ori_img_size=640,
                 warmup_epoch=0, epoch_num=0, use_dfl=False, reg_max=0, iou_type="giou",
                 loss_weight=None, return_assign=False, norm_gt_zero=False):
    """sizes: [(h,w)] per level; pred_scores [B,A,nc] (post-sigmoid), pred_distri [B,A,4*(reg_max+1)];
    targets [n,6].  Returns (loss, loss_items[iou, dfl, cls]) like the reference."""
    lw = loss_weight or {"class": 1.0, "iou": 2.5, "dfl": 0.5}
    norm_thr = 0 if norm_gt_zero else 1      # loss.py:168-169 divides when the sum exceeds 1, loss_distill.py / loss_fuseab.py when it is > 0
    B = pred_scores.shape[0]
    anchors, anchor_points, n_list, stride_t = assign.train_anchors(sizes, strides, dtype=pred_scores.dtype)
    scale = torch.tensor([ori_img_size] * 4, dtype=pred_scores.dtype)           # loss.py:72
    t = preprocess_targets(targets, B, scale)
    gt_labels, gt_bboxes = t[:, :, :1], t[:, :, 1:]
    mask_gt = (gt_bboxes.sum(-1, keepdim=True) > 0).float()                      # loss.py:79
    anchor_points_s = anchor_points / stride_t
    pred_bboxes = decode_pred(pred_distri, anchor_points_s, use_dfl, reg_max)    # loss.py:83
    with torch.no_grad():
        if epoch_num < warmup_epoch:                                             # loss.py:86-103
            tl, tb, ts, fg, gi = assign.atss_assign(anchors, n_list, gt_labels, gt_bboxes, mask_gt,
                                                    pred_bboxes.detach() * stride_t, num_classes=num_classes)
        else:
            tl, tb, ts, fg, gi = assign.tal_assign(pred_scores.detach(), pred_bboxes.detach() * stride_t,
                                                   anchor_points, gt_labels, gt_bboxes, mask_gt,
                                                   num_classes=num_classes)
    tb = tb / stride_t                                                           # loss.py:158
    tl = torch.where(fg > 0, tl, torch.full_like(tl, num_classes))               # loss.py:161
    one_hot = F.one_hot(tl.long(), num_classes + 1)[..., :-1]
    # VarifocalLoss, loss.py:205-211 (alpha .75, gamma 2; gradient flows through the weight too)
    weight = 0.75 * pred_scores.pow(2.0) * (1 - one_hot) + ts * one_hot
    loss_cls = (F.binary_cross_entropy(pred_scores.float(), ts.float(), reduction="none") * weight).sum()
    tss = ts.sum()
    if tss > norm_thr:                                                                  # loss.py:168-169
        loss_cls = loss_cls / tss
    # BboxLoss, loss.py:222-263
    num_pos = fg.sum()
    if num_pos > 0:
        pb = pred_bboxes[fg]
        tbp = tb[fg]
        bw = ts.sum(-1)[fg].unsqueeze(-1)
        loss_iou = (iou_loss(pb, tbp, iou_type) * bw).sum()
        if tss > norm_thr:
            loss_iou = loss_iou / tss
        if use_dfl:
            pd_pos = pred_distri[fg].view(-1, 4, reg_max + 1)
            x1y1, x2y2 = torch.split(tb, 2, -1)                                  # bbox2dist, general.py:46-52
            ltrb = torch.cat([anchor_points_s - x1y1, x2y2 - anchor_points_s], -1).clip(0, reg_max - 0.01)[fg]
            tl_ = ltrb.to(torch.long)
            tr_ = tl_ + 1
            wl = tr_.to(torch.float) - ltrb
            wr = 1 - wl
            ce_l = F.cross_entropy(pd_pos.view(-1, reg_max + 1), tl_.view(-1), reduction="none").view(tl_.shape) * wl
            ce_r = F.cross_entropy(pd_pos.view(-1, reg_max + 1), tr_.view(-1), reduction="none").view(tl_.shape) * wr
            loss_dfl = ((ce_l + ce_r).mean(-1, keepdim=True) * bw).sum()
            if tss > norm_thr:
                loss_dfl = loss_dfl / tss
        else:
            loss_dfl = pred_distri.sum() * 0.0
    else:
        loss_iou = pred_distri.sum() * 0.0
        loss_dfl = pred_distri.sum() * 0.0
    loss = lw["class"] * loss_cls + lw["iou"] * loss_iou + lw["dfl"] * loss_dfl
    items = torch.stack([lw["iou"] * loss_iou, lw["dfl"] * loss_dfl, lw["class"] * loss_cls]).detach()
    if return_assign:
        return loss, items, dict(labels=tl, bboxes=tb, scores=ts, fg=fg, gt_idx=gi, pred_bboxes=pred_bboxes,
                                 gt_labels=gt_labels, gt_bboxes=gt_bboxes, mask_gt=mask_gt)
    return loss, items


def drop_targets(targets, drop):
    """Golden-case helper (12th field of tests/golden/loss_cases.json): remove the boxes of image 1 ("img1") or of
    every image ("all") -- the ragged / empty inputs of ComputeLoss.preprocess (loss.py:184-192)."""
    if drop == "img1":
        return targets[targets[:, 0] != 1]
    if drop == "all":
        return targets[:0]
    return targets


def synthetic_targets(batch, seed=1, mean_per_image=7.3, num_classes=80):
    """COCO-shaped synthetic labels: per image n ~ clip(Poisson(7.3), 1, 60);
    cls ~ U{0..nc-1}; w,h ~ U(.02,.6); centres uniform such that the box stays inside.  [n,6] fp32."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(batch):
        n = int(torch.poisson(torch.tensor([mean_per_image]), generator=g).clamp(1, 60).item())
        cls = torch.randint(0, num_classes, (n, 1), generator=g).float()
        wh = torch.rand(n, 2, generator=g) * 0.58 + 0.02
        cxy = wh / 2 + torch.rand(n, 2, generator=g) * (1 - wh)
        rows.append(torch.cat([torch.full((n, 1), float(b)), cls, cxy, wh], 1))
    return torch.cat(rows).float()
