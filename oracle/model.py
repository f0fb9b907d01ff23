"""oracle/model.py -- CPU restatement of the YOLOv6 forward pass (TEST INFRASTRUCTURE ONLY).

This file is the parity oracle for the network part of the hot path.  It is NOT product code: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import it; nothing under `yolov6_b200/` does.

It re-states, as plain functional PyTorch on CPU tensors (NCHW, fp32 or fp64), what the reference
modules compute in eval mode from a *train-form* reference `state_dict` (BatchNorm running stats,
three-branch RepVGG blocks) -- i.e. without any of the weight folding the product performs, so the
product's BN-fold / re-parameterisation (reference torch_utils.py:50-94, common.py:257-319) is
itself under test.  Each function cites the reference code it follows.

Pinning: `tests/golden/make_golden.py` imports the unmodified reference from /root/reference,
runs it on seeded inputs and stores small fixtures; `tests/test_oracle_model.py` checks this file
against them (logits, decoded boxes, per-stage feature statistics).  The train mode (`train_mode()`:
batch-statistics BatchNorm, train branch of Detect, BottleRep alpha) is pinned the same way by
`tests/golden/make_golden_train.py`: head outputs, a scalar loss and the gradient of every parameter of
YOLOv6-N and -M against the reference model in `.train()` mode in float64 (agreement 1e-9 / 1e-7).
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # reference yolov6/utils/torch_utils.py:41-43 (initialize_weights sets eps=1e-3)

# ---------------------------------------------------------------------------------------------
# model configurations (numbers restated from reference configs/yolov6{n,s,m,l6}.py)
# ---------------------------------------------------------------------------------------------
CONFIGS = {
    "yolov6n": dict(depth=0.33, width=0.25, backbone="EfficientRep", neck="RepBiFPANNeck",
                    bb_repeats=[1, 6, 12, 18, 6], bb_channels=[64, 128, 256, 512, 1024],
                    neck_repeats=[12, 12, 12, 12], neck_channels=[256, 128, 128, 256, 256, 512],
                    cspsppf=True, fuse_P2=True, csp_e=None, num_layers=3, strides=[8, 16, 32],
                    use_dfl=False, reg_max=0, iou_type="siou", mode="repvgg", atss_warmup_epoch=0),
    "yolov6s": dict(depth=0.33, width=0.50, backbone="EfficientRep", neck="RepBiFPANNeck",
                    bb_repeats=[1, 6, 12, 18, 6], bb_channels=[64, 128, 256, 512, 1024],
                    neck_repeats=[12, 12, 12, 12], neck_channels=[256, 128, 128, 256, 256, 512],
                    cspsppf=True, fuse_P2=True, csp_e=None, num_layers=3, strides=[8, 16, 32],
                    use_dfl=False, reg_max=0, iou_type="giou", mode="repvgg", atss_warmup_epoch=0),
    "yolov6m": dict(depth=0.60, width=0.75, backbone="CSPBepBackbone", neck="CSPRepBiFPANNeck",
                    bb_repeats=[1, 6, 12, 18, 6], bb_channels=[64, 128, 256, 512, 1024],
                    neck_repeats=[12, 12, 12, 12], neck_channels=[256, 128, 128, 256, 256, 512],
                    cspsppf=False, fuse_P2=True, csp_e=2.0 / 3, num_layers=3, strides=[8, 16, 32],
                    use_dfl=True, reg_max=16, iou_type="giou", mode="repvgg", atss_warmup_epoch=0),
    "yolov6l6": dict(depth=1.0, width=1.0, backbone="CSPBepBackbone_P6", neck="CSPRepBiFPANNeck_P6",
                     bb_repeats=[1, 6, 12, 18, 6, 6], bb_channels=[64, 128, 256, 512, 768, 1024],
                     neck_repeats=[12, 12, 12, 12, 12, 12], neck_channels=[512, 256, 128, 256, 512, 1024],
                     cspsppf=False, fuse_P2=True, csp_e=0.5, num_layers=4, strides=[8, 16, 32, 64],
                     use_dfl=True, reg_max=16, iou_type="giou", mode="conv_silu", atss_warmup_epoch=4),
}


def scaled_lists(cfg):
    """Depth / width scaling of reference yolov6/models/yolo.py:66-67."""
    reps = [(max(round(i * cfg["depth"]), 1) if i > 1 else i) for i in cfg["bb_repeats"] + cfg["neck_repeats"]]
    chans = [math.ceil(i * cfg["width"] / 8) * 8 for i in cfg["bb_channels"] + cfg["neck_channels"]]
    return reps, chans


# ---------------------------------------------------------------------------------------------
# layer restatements
# ---------------------------------------------------------------------------------------------
_BN_TRAIN = [False]


class train_mode:
    """Context manager: BatchNorm uses batch statistics (nn.BatchNorm2d.train(), biased variance in the
    normalisation) so that the train-form forward/backward of the reference can be re-stated for the
    gradient parity tests.  Running statistics are not updated (the oracle is stateless)."""

    def __enter__(self):
        _BN_TRAIN.append(True)

    def __exit__(self, *a):
        _BN_TRAIN.pop()


_QUANT = [False]


class bf16_storage:
    """Context manager: emulate the storage precision of the CUDA training path -- conv weights (except
    the fp32 stem) and every tensor the kernels write to HBM (raw conv outputs, block outputs) are rounded
    to bf16, arithmetic stays in the tensor's dtype.  Lets the gradient parity test compare like with
    like instead of measuring how bf16 rounding compounds through ~50 batch-normalised layers."""

    def __enter__(self):
        _QUANT.append(True)

    def __exit__(self, *a):
        _QUANT.pop()


def _q(t):
    return t.to(torch.bfloat16).to(t.dtype) if _QUANT[-1] else t


def _wq(w, name=""):
    return w if (not _QUANT[-1] or name.startswith("backbone.stem")) else w.to(torch.bfloat16).to(w.dtype)


def _bn(sd, p, x):
    """BatchNorm2d, eps=1e-3: running stats in eval mode, batch stats inside `train_mode()`."""
    w, b = sd[p + ".weight"].to(x.dtype), sd[p + ".bias"].to(x.dtype)
    if _BN_TRAIN[-1]:
        m = x.mean(dim=(0, 2, 3))
        v = x.var(dim=(0, 2, 3), unbiased=False)
        scale = w / torch.sqrt(v + BN_EPS)
        return x * scale.view(1, -1, 1, 1) + (b - m * scale).view(1, -1, 1, 1)
    m, v = sd[p + ".running_mean"].to(x.dtype), sd[p + ".running_var"].to(x.dtype)
    scale = w / torch.sqrt(v + BN_EPS)
    return x * scale.view(1, -1, 1, 1) + (b - m * scale).view(1, -1, 1, 1)


def _act(x, act):
    if act == "relu":
        return torch.relu(x)
    if act == "silu":
        return x * torch.sigmoid(x)
    assert act is None, act
    return x


def conv_module(sd, p, x, stride, act):
    """ConvModule.forward, reference yolov6/layers/common.py:26-49: conv(no bias, pad=k//2) -> BN -> act."""
    w = _wq(sd[p + ".conv.weight"].to(x.dtype), p)
    k = w.shape[-1]
    y = _act(_bn(sd, p + ".bn", _q(F.conv2d(x, w, None, stride=stride, padding=k // 2))), act)
    return _q(y) if act is not None or not p.endswith(("rbr_dense", "rbr_1x1")) else y


def conv_bn_act(sd, p, x, stride, act):
    """ConvBNReLU / ConvBNSiLU wrappers (common.py:57-76): a ConvModule stored under `.block`."""
    return conv_module(sd, p + ".block", x, stride, act)


def repvgg(sd, p, x, stride):
    """RepVGGBlock.forward in train form, common.py:245-255: relu(BN(3x3) + BN(1x1, pad 0) + BN(x))."""
    y = conv_module(sd, p + ".rbr_dense", x, stride, None)
    w1 = _wq(sd[p + ".rbr_1x1.conv.weight"].to(x.dtype), p)
    y = y + _bn(sd, p + ".rbr_1x1.bn", _q(F.conv2d(x, w1, None, stride=stride, padding=0)))
    if p + ".rbr_identity.weight" in sd:
        y = y + _bn(sd, p + ".rbr_identity", x)
    return _q(torch.relu(y))


def basic_block(sd, p, x, stride, mode):
    """get_block(training_mode), common.py:721-737."""
    if mode == "repvgg":
        return repvgg(sd, p, x, stride)
    return conv_bn_act(sd, p, x, stride, "silu" if mode == "conv_silu" else "relu")


def rep_block(sd, p, x, n, mode):
    """RepBlock with a plain basic block, common.py:569-588."""
    x = basic_block(sd, p + ".conv1", x, 1, mode)
    for i in range(n - 1):
        x = basic_block(sd, f"{p}.block.{i}", x, 1, mode)
    return x


def bottle_rep(sd, p, x, mode):
    """BottleRep.forward, common.py:591-608: conv2(conv1(x)) + alpha * x when Cin == Cout."""
    y = basic_block(sd, p + ".conv2", basic_block(sd, p + ".conv1", x, 1, mode), 1, mode)
    if y.shape[1] == x.shape[1]:
        y = y + sd[p + ".alpha"].to(x.dtype) * x
    return y


def rep_block_bottle(sd, p, x, n, mode):
    """RepBlock(block=BottleRep), common.py:579-582: n//2 BottleReps."""
    x = bottle_rep(sd, p + ".conv1", x, mode)
    for i in range(n // 2 - 1):
        x = bottle_rep(sd, f"{p}.block.{i}", x, mode)
    return x


def bepc3(sd, p, x, n, mode):
    """BepC3.forward, common.py:634-650: cv3(cat(m(cv1 x), cv2 x))."""
    act = "silu" if mode == "conv_silu" else "relu"
    a = rep_block_bottle(sd, p + ".m", conv_bn_act(sd, p + ".cv1", x, 1, act), n, mode)
    b = conv_bn_act(sd, p + ".cv2", x, 1, act)
    return conv_bn_act(sd, p + ".cv3", torch.cat((a, b), 1), 1, act)


def _pool5(x):
    return F.max_pool2d(x, 5, 1, 2)


def sppf(sd, p, x, act):
    """SPPFModule.forward, common.py:97-112 (stored under `.sppf`)."""
    p = p + ".sppf"
    x = conv_bn_act(sd, p + ".cv1", x, 1, act)
    y1 = _pool5(x)
    y2 = _pool5(y1)
    return conv_bn_act(sd, p + ".cv2", torch.cat([x, y1, y2, _pool5(y2)], 1), 1, act)


def cspsppf(sd, p, x, act):
    """CSPSPPFModule.forward, common.py:135-158 (stored under `.cspsppf`)."""
    p = p + ".cspsppf"
    x1 = conv_bn_act(sd, p + ".cv4", conv_bn_act(sd, p + ".cv3", conv_bn_act(sd, p + ".cv1", x, 1, act), 1, act), 1, act)
    y0 = conv_bn_act(sd, p + ".cv2", x, 1, act)
    y1 = _pool5(x1)
    y2 = _pool5(y1)
    y3 = conv_bn_act(sd, p + ".cv6", conv_bn_act(sd, p + ".cv5", torch.cat([x1, y1, y2, _pool5(y2)], 1), 1, act), 1, act)
    return conv_bn_act(sd, p + ".cv7", torch.cat((y0, y3), 1), 1, act)


def bifusion(sd, p, xs):
    """BiFusion.forward, common.py:695-718 (always ReLU): cv3(cat(up(x0), cv1(x1), down(cv2(x2))))."""
    x0 = _q(F.conv_transpose2d(xs[0], _wq(sd[p + ".upsample.upsample_transpose.weight"].to(xs[0].dtype)),
                               sd[p + ".upsample.upsample_transpose.bias"].to(xs[0].dtype), stride=2))
    x1 = conv_bn_act(sd, p + ".cv1", xs[1], 1, "relu")
    x2 = conv_bn_act(sd, p + ".downsample", conv_bn_act(sd, p + ".cv2", xs[2], 1, "relu"), 2, "relu")
    return conv_bn_act(sd, p + ".cv3", torch.cat((x0, x1, x2), 1), 1, "relu")


# ---------------------------------------------------------------------------------------------
# backbones / necks / head
# ---------------------------------------------------------------------------------------------
def backbone(sd, cfg, x):
    """EfficientRep (efficientrep.py:7-118), CSPBepBackbone (:250-374), CSPBepBackbone_P6 (:377-516)."""
    reps, _ = scaled_lists(cfg)
    mode = cfg["mode"]
    act = "silu" if mode == "conv_silu" else "relu"
    csp = cfg["backbone"].startswith("CSP")
    nstage = 6 if cfg["backbone"].endswith("P6") else 5
    outs = []
    x = basic_block(sd, "backbone.stem", x, 2, mode)
    for s in range(2, nstage + 1):
        p = f"backbone.ERBlock_{s}"
        x = basic_block(sd, p + ".0", x, 2, mode)
        x = bepc3(sd, p + ".1", x, reps[s - 1], mode) if csp else rep_block(sd, p + ".1", x, reps[s - 1], mode)
        if s == nstage:
            x = cspsppf(sd, p + ".2", x, act) if cfg["cspsppf"] else sppf(sd, p + ".2", x, act)
        if s > 2 or cfg["fuse_P2"]:
            outs.append(x)
    return outs


def neck(sd, cfg, feats):
    """RepBiFPANNeck (reppan.py:132-237), CSPRepBiFPANNeck (:666-785), CSPRepBiFPANNeck_P6 (:955-1116)."""
    reps, _ = scaled_lists(cfg)
    mode = cfg["mode"]
    csp = cfg["neck"].startswith("CSP")
    nb = len(cfg["bb_repeats"])

    def stage(p, x, n):
        return bepc3(sd, p, x, n, mode) if csp else rep_block(sd, p, x, n, mode)

    if not cfg["neck"].endswith("P6"):
        x3, x2, x1, x0 = feats
        fpn0 = conv_bn_act(sd, "neck.reduce_layer0", x0, 1, "relu")
        f0 = stage("neck.Rep_p4", bifusion(sd, "neck.Bifusion0", [fpn0, x1, x2]), reps[nb + 0])
        fpn1 = conv_bn_act(sd, "neck.reduce_layer1", f0, 1, "relu")
        pan2 = stage("neck.Rep_p3", bifusion(sd, "neck.Bifusion1", [fpn1, x2, x3]), reps[nb + 1])
        d1 = conv_bn_act(sd, "neck.downsample2", pan2, 2, "relu")
        pan1 = stage("neck.Rep_n3", torch.cat([d1, fpn1], 1), reps[nb + 2])
        d0 = conv_bn_act(sd, "neck.downsample1", pan1, 2, "relu")
        pan0 = stage("neck.Rep_n4", torch.cat([d0, fpn0], 1), reps[nb + 3])
        return [pan2, pan1, pan0]
    x4, x3, x2, x1, x0 = feats
    fpn0 = conv_bn_act(sd, "neck.reduce_layer0", x0, 1, "relu")
    f0 = stage("neck.Rep_p5", bifusion(sd, "neck.Bifusion0", [fpn0, x1, x2]), reps[nb + 0])
    fpn1 = conv_bn_act(sd, "neck.reduce_layer1", f0, 1, "relu")
    f1 = stage("neck.Rep_p4", bifusion(sd, "neck.Bifusion1", [fpn1, x2, x3]), reps[nb + 1])
    fpn2 = conv_bn_act(sd, "neck.reduce_layer2", f1, 1, "relu")
    pan3 = stage("neck.Rep_p3", bifusion(sd, "neck.Bifusion2", [fpn2, x3, x4]), reps[nb + 2])
    d2 = conv_bn_act(sd, "neck.downsample2", pan3, 2, "relu")
    pan2 = stage("neck.Rep_n4", torch.cat([d2, fpn2], 1), reps[nb + 3])
    d1 = conv_bn_act(sd, "neck.downsample1", pan2, 2, "relu")
    pan1 = stage("neck.Rep_n5", torch.cat([d1, fpn1], 1), reps[nb + 4])
    d0 = conv_bn_act(sd, "neck.downsample0", pan1, 2, "relu")
    pan0 = stage("neck.Rep_n6", torch.cat([d0, fpn0], 1), reps[nb + 5])
    return [pan3, pan2, pan1, pan0]


def head_raw(sd, cfg, feats):
    """Detect.forward up to the per-level outputs (effidehead.py:72-92 / 97-118).
    Returns (cls [B,A,nc] post-sigmoid, reg [B,A,4*(reg_max+1)] raw), levels concatenated over A."""
    cls_all, reg_all = [], []
    for i, x in enumerate(feats):
        x = conv_bn_act(sd, f"detect.stems.{i}", x, 1, "silu")
        cf = conv_bn_act(sd, f"detect.cls_convs.{i}", x, 1, "silu")
        rf = conv_bn_act(sd, f"detect.reg_convs.{i}", x, 1, "silu")
        c = F.conv2d(cf, _wq(sd[f"detect.cls_preds.{i}.weight"].to(x.dtype)), sd[f"detect.cls_preds.{i}.bias"].to(x.dtype))
        r = F.conv2d(rf, _wq(sd[f"detect.reg_preds.{i}.weight"].to(x.dtype)), sd[f"detect.reg_preds.{i}.bias"].to(x.dtype))
        cls_all.append(torch.sigmoid(c).flatten(2).permute(0, 2, 1))
        reg_all.append(r.flatten(2).permute(0, 2, 1))
    return torch.cat(cls_all, 1), torch.cat(reg_all, 1)


AB_ANCHORS_INIT = [[10, 13, 19, 19, 33, 23], [30, 61, 59, 59, 59, 119], [116, 90, 185, 185, 373, 326]]   # configs/yolov6{n,s,m}.py head.anchors_init


def head_ab(sd, cfg, feats, anchors_init=None):
    """Training branch of the fuse_ab head, effidehead_fuseab.py:94-140: besides the anchor-free outputs, per level
    cls_preds_ab / reg_preds_ab emit na = 3 predictions per pixel, reshaped to (b, na, h, w, .) and flattened to rows
    (anchor, pixel); reg[..., 2:4] = (2 sigmoid)^2 * anchors_init / stride.  Returns (cls_ab [B,3A,nc], reg_ab [B,3A,4])."""
    anchors_init = anchors_init or AB_ANCHORS_INIT
    na = 3
    cls_all, reg_all = [], []
    for i, x in enumerate(feats):
        b, _, h, w = x.shape
        x = conv_bn_act(sd, f"detect.stems.{i}", x, 1, "silu")
        cf = conv_bn_act(sd, f"detect.cls_convs.{i}", x, 1, "silu")
        rf = conv_bn_act(sd, f"detect.reg_convs.{i}", x, 1, "silu")
        c = F.conv2d(cf, _wq(sd[f"detect.cls_preds_ab.{i}.weight"].to(x.dtype)), sd[f"detect.cls_preds_ab.{i}.bias"].to(x.dtype))
        r = F.conv2d(rf, _wq(sd[f"detect.reg_preds_ab.{i}.weight"].to(x.dtype)), sd[f"detect.reg_preds_ab.{i}.bias"].to(x.dtype))
        c = torch.sigmoid(c).reshape(b, na, -1, h, w).permute(0, 1, 3, 4, 2)                 # :112-114
        r = r.reshape(b, na, -1, h, w).permute(0, 1, 3, 4, 2)                               # :116
        anc = (torch.tensor(anchors_init[i], dtype=x.dtype) / cfg["strides"][i]).reshape(1, na, 1, 1, 2)   # :35
        wh = ((r[..., 2:4].sigmoid() * 2) ** 2) * anc                                       # :117
        r = torch.cat([r[..., :2], wh], -1)
        cls_all.append(c.flatten(1, 3))
        reg_all.append(r.flatten(1, 3))
    return torch.cat(cls_all, 1), torch.cat(reg_all, 1)


def eval_anchor_points(sizes, strides, dtype=torch.float32):
    """generate_anchors(is_eval=True, mode='af'), anchor_generator.py:13-33: cell centres (+0.5) in
    grid units and the per-anchor stride column."""
    pts, strs = [], []
    for (h, w), s in zip(sizes, strides):
        sx = torch.arange(w, dtype=dtype) + 0.5
        sy = torch.arange(h, dtype=dtype) + 0.5
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack([xx, yy], -1).reshape(-1, 2))
        strs.append(torch.full((h * w, 1), float(s), dtype=dtype))
    return torch.cat(pts), torch.cat(strs)


def decode_eval(cfg, cls, reg, sizes):
    """Eval tail of Detect.forward, effidehead.py:106-139 + dist2bbox(xywh), general.py:32-43:
    DFL softmax.proj (when use_dfl), ltrb -> (cx,cy,w,h), x stride, cat [xywh, 1, cls]."""
    B, A, _ = reg.shape
    if cfg["use_dfl"]:
        R = cfg["reg_max"] + 1
        proj = torch.arange(R, dtype=reg.dtype)
        reg = (F.softmax(reg.reshape(B, A, 4, R), -1) * proj).sum(-1)
    pts, strs = eval_anchor_points(sizes, cfg["strides"], reg.dtype)
    lt, rb = reg[..., :2], reg[..., 2:]
    x1y1, x2y2 = pts - lt, pts + rb
    box = torch.cat([(x1y1 + x2y2) / 2, x2y2 - x1y1], -1) * strs
    return torch.cat([box, torch.ones(B, A, 1, dtype=reg.dtype), cls], -1)


def head_dist(sd, cfg, feats):
    """`reg_preds_dist` branch of the N / S distillation student (effidehead_distill_ns.py:93-101): DFL logits [B,A,4*(reg_max+1)]."""
    out = []
    for i, x in enumerate(feats):
        x = conv_bn_act(sd, f"detect.stems.{i}", x, 1, "silu")
        rf = conv_bn_act(sd, f"detect.reg_convs.{i}", x, 1, "silu")
        r = F.conv2d(rf, _wq(sd[f"detect.reg_preds_dist.{i}.weight"].to(x.dtype)), sd[f"detect.reg_preds_dist.{i}.bias"].to(x.dtype))
        out.append(r.flatten(2).permute(0, 2, 1))
    return torch.cat(out, 1)


def forward(sd, cfg, x, train_outputs=False, fuse_ab=False, distill_ns=False):
    """Model.forward, yolo.py:33-41.  x: [B,3,H,W] in [0,1].  Eval: [B,A,5+nc].
    train_outputs=True returns the train-mode head tensors (cls post-sigmoid, reg raw) computed
    with eval-mode BN -- used to pin the loss / assigner inputs; with fuse_ab also (cls_ab, reg_ab)."""
    feats = neck(sd, cfg, backbone(sd, cfg, x))
    sizes = [tuple(f.shape[2:]) for f in feats]
    cls, reg = head_raw(sd, cfg, feats)
    if train_outputs and fuse_ab:
        cls_ab, reg_ab = head_ab(sd, cfg, feats)
        return cls, reg, sizes, cls_ab, reg_ab
    if train_outputs and distill_ns:       # (cls, reg_lrtb, sizes, reg_dist): effidehead_distill_ns.py:104
        return cls, reg, sizes, head_dist(sd, cfg, feats)
    if distill_ns:                          # eval: plain lrtb distances, no DFL (effidehead_distill_ns.py:105-150)
        return decode_eval(dict(cfg, use_dfl=False, reg_max=0), cls, reg, sizes)
    if train_outputs:
        return cls, reg, sizes
    return decode_eval(cfg, cls, reg, sizes)
