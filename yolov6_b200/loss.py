"""Drop-in `ComputeLoss` backed by the sm_100a assignment + fused loss kernels.

Same constructor and call signature as the reference (yolov6/models/losses/loss.py:14-182):
    ComputeLoss(fpn_strides, grid_cell_size, grid_cell_offset, num_classes, ori_img_size, warmup_epoch,
                use_dfl, reg_max, iou_type, loss_weight)
    loss, loss_items = compute_loss(outputs, targets, epoch_num, step_num, batch_height, batch_width)
`loss` is a float64 scalar that is differentiable w.r.t. pred_scores / pred_distri (the gradients are
produced by the same kernel launch as the forward values); `loss_items` = [iou, dfl, cls] weighted and
detached (loss.py:179-182).  The whole path (target padding, box decode, TAL or ATSS, VFL + IoU + DFL,
normalisation) runs on the device with no host synchronisation; failures surface as RuntimeError.
"""
import ctypes as C

import torch

from . import _lib
from .assigners import _p, atss_compact, generate_anchors_train, tal_compact

IOU_TYPES = {"giou": 0, "siou": 1, "ciou": 2, "diou": 3}


class _DetLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_scores, pred_distri, state):
        ctx.save_for_backward(state["grad_scores"], state["grad_distri"])
        ctx.shapes = (pred_scores.dtype, pred_distri.dtype)
        return state["out"][0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        gs, gd = ctx.saved_tensors
        return (gs * grad_out).to(ctx.shapes[0]), (gd * grad_out).to(ctx.shapes[1]), None


class ComputeLoss:
    def __init__(self, fpn_strides=[8, 16, 32], grid_cell_size=5.0, grid_cell_offset=0.5, num_classes=80,
                 ori_img_size=640, warmup_epoch=4, use_dfl=True, reg_max=16, iou_type='giou',
                 loss_weight={'class': 1.0, 'iou': 2.5, 'dfl': 0.5}):
        self.fpn_strides = list(fpn_strides)
        self.grid_cell_size, self.grid_cell_offset = grid_cell_size, grid_cell_offset
        self.num_classes, self.ori_img_size = num_classes, ori_img_size
        self.warmup_epoch = warmup_epoch
        self.use_dfl, self.reg_max = use_dfl, reg_max
        self.iou_type = iou_type.lower()
        if self.iou_type not in IOU_TYPES:
            raise ValueError(f"unknown iou_type {iou_type}")
        self.loss_weight = loss_weight
        self._anchor_key, self._anchors = None, None
        self.last_assignment = None
        self._norm_gt_zero = 0      # subclasses: 1 = divide by target_scores_sum whenever it is > 0 (loss_distill.py:190-191, 318-323)

    def _get_anchors(self, sizes, device):
        key = (tuple(sizes), str(device))
        if key != self._anchor_key:       # cached like loss.py:63-69
            self._anchors = generate_anchors_train(sizes, self.fpn_strides, device, self.grid_cell_size, self.grid_cell_offset)
            self._anchor_key = key
        return self._anchors

    def __call__(self, outputs, targets, epoch_num, step_num, batch_height, batch_width):
        feats, pred_scores, pred_distri = outputs
        sizes = [tuple(f.shape[2:]) for f in feats]
        state = self.forward_backward(pred_scores, pred_distri, sizes, targets, epoch_num, batch_height, batch_width)
        loss = _DetLossFn.apply(pred_scores, pred_distri, state)
        return loss, state["out"][1:4].detach().clone()

    def forward_backward(self, pred_scores, pred_distri, sizes, targets, epoch_num, batch_height, batch_width, max_gt=None,
                         grad_scores=None, grad_distri=None, grad_scale=1.0, reuse=None, weights=None):
        """Loss value and its gradients w.r.t. the head outputs in one pass (no autograd): returns
        {"out": float64[8] (loss, iou, dfl, cls, target_scores_sum, num_pos), "grad_scores", "grad_distri", "gt_count"}.
        `max_gt` fixes the padded target count G (static shapes for CUDA-graph capture; images with more boxes are
        reported through gt_count > G); otherwise G is the largest per-image count, as in loss.py:184-192 (one small
        host read when `targets` lives on the device).  `grad_*` let the caller provide the output buffers.
        `reuse` = the state of an earlier call: its padded targets and label assignment are used as they are (a second box
        branch scored against the same assignment, loss_distill_ns.py:93,150-160); `weights` overrides the (class, iou, dfl) weights."""
        dev = pred_scores.device
        if dev.type != "cuda":
            raise RuntimeError("yolov6_b200.ComputeLoss runs on CUDA tensors only (no CPU fallback)")
        anchors, anchor_points, n_list, stride_t = self._get_anchors(sizes, dev)
        B, A, nc = pred_scores.shape
        R = pred_distri.shape[2]
        lib, h, sp = _lib.lib(), _lib.handle(dev.index or 0), _lib.stream_ptr()
        ps = pred_scores.detach().float().contiguous()
        pd = pred_distri.detach().float().contiguous()
        strides = stride_t.reshape(-1).contiguous()
        if reuse is not None:       # second box branch against the first call's padded targets and assignment
            gt, gt_count, G, c, mask = reuse["gt"], reuse["gt_count"], reuse["G"], reuse["assignment"], reuse["mask"]
            tg = pboxes = None
        else:
            n = targets.shape[0]
            if max_gt is not None:
                G = max(int(max_gt), 1)
            elif n == 0:
                G = 1
            else:   # loss.py:184-192 pads to the largest per-image count
                img = targets[:, 0].detach()
                valid = img[(img >= 0) & (img < B)].long()
                G = max(int(torch.bincount(valid, minlength=1).max()), 1) if valid.numel() else 1
            tg = targets.detach().float().contiguous().to(dev)
            gt = torch.empty(B, G, 5, dtype=torch.float64, device=dev)
            gt_count = torch.empty(B, dtype=torch.int32, device=dev)
            _lib.check(lib.yv6_targets_pad(h, _p(tg), n, B, G, float(batch_width), float(batch_height), _p(gt), _p(gt_count), sp))
            mask = (gt[:, :, 1:].sum(-1) > 0).to(torch.uint8).contiguous()          # loss.py:79
            # predicted boxes in pixels for the assigner (loss.py:82-83, 94/100)
            pboxes = torch.empty(B, A, 4, dtype=torch.float32, device=dev)
            _lib.check(lib.yv6_box_decode(h, _p(pd), _p(anchor_points), _p(strides), B, A, R, 1, _p(pboxes), sp))
            if epoch_num < self.warmup_epoch:
                c = atss_compact(anchors, n_list, gt, mask, pboxes, nc, 9)
            else:
                c = tal_compact(ps, pboxes, anchor_points, gt, mask, 13, 1.0, 6.0)
            self.last_assignment = c
        d = _lib.LossDesc()
        if grad_scores is None:
            grad_scores = torch.empty_like(ps)
        if grad_distri is None:
            grad_distri = torch.empty_like(pd)
        out = torch.zeros(8, dtype=torch.float64, device=dev)
        ws = torch.empty(int(lib.yv6_det_loss_workspace_bytes(B, A)), dtype=torch.uint8, device=dev)
        d.pred_scores, d.pred_distri, d.anc_points, d.strides = ps.data_ptr(), pd.data_ptr(), anchor_points.data_ptr(), strides.data_ptr()
        d.gt, d.gt_idx, d.fg, d.norm = gt.data_ptr(), c.gt_idx.data_ptr(), c.fg.data_ptr(), c.norm.data_ptr()
        d.B, d.A, d.G, d.nc, d.reg_ch = B, A, G, nc, R
        d.iou_type = IOU_TYPES[self.iou_type]
        d.w_cls, d.w_iou, d.w_dfl = float(self.loss_weight['class']), float(self.loss_weight['iou']), float(self.loss_weight['dfl'])
        if weights is not None:
            d.w_cls, d.w_iou, d.w_dfl = (float(v) for v in weights)
        d.grad_scale = float(grad_scale)
        d.grad_scores, d.grad_distri, d.out = grad_scores.data_ptr(), grad_distri.data_ptr(), out.data_ptr()
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        d.norm_gt_zero = self._norm_gt_zero
        _lib.check(lib.yv6_det_loss(h, C.byref(d), sp))
        return {"grad_scores": grad_scores, "grad_distri": grad_distri, "out": out, "gt_count": gt_count, "G": G, "gt": gt, "mask": mask,
                "assignment": c, "keep": (ps, pd, gt, mask, pboxes, ws, tg)}
