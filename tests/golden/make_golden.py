"""tests/golden/make_golden.py -- mint golden vectors by running the UNMODIFIED reference.

The reference ships no tests, fixtures or golden vectors, so parity is pinned here:
this script imports meituan/YOLOv6 from /root/reference (read-only; never copied), runs its own
PyTorch CPU path on seeded synthetic inputs and stores small outputs under tests/golden/.
It only runs in the build container (the GPU box has no /root/reference); the stored fixtures travel.

    PYTHONPATH=tests/golden/refshim:/root/reference:. python tests/golden/make_golden.py

Inputs are NOT stored: tests regenerate them from seeds with oracle/fabricate.py and verify the
stored checksums first, so RNG drift is reported as such rather than as a parity failure.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]

torch.cuda.is_available = lambda: False
nn.Module.cuda = lambda self, *a, **k: self  # ComputeLoss.__init__ calls .cuda() on parameter-less modules

from yolov6.layers.common import RepVGGBlock  # noqa: E402
from yolov6.models.losses.loss import ComputeLoss  # noqa: E402
from yolov6.models.yolo import build_model  # noqa: E402
from yolov6.utils.config import Config  # noqa: E402
from yolov6.utils.nms import non_max_suppression  # noqa: E402
from yolov6.utils.torch_utils import fuse_model  # noqa: E402

from oracle import fabricate as fab  # noqa: E402
from oracle import loss as oloss  # noqa: E402

MODELS = {"yolov6n": 64, "yolov6s": 64, "yolov6m": 64, "yolov6l6": 128}  # name -> golden input size

NMS_CASES = [  # (B, A, nc, seed, kwargs)
    (2, 500, 80, 0, dict(conf_thres=0.25, iou_thres=0.45)),
    (2, 500, 80, 0, dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)),
    (3, 2000, 20, 1, dict(conf_thres=0.1, iou_thres=0.5, agnostic=True)),
    (3, 2000, 20, 1, dict(conf_thres=0.05, iou_thres=0.45, classes=[0, 3, 5], max_det=50)),
    (1, 8400, 80, 2, dict(conf_thres=0.4, iou_thres=0.45, max_det=1000)),
    (2, 300, 1, 3, dict(conf_thres=0.03, iou_thres=0.65, multi_label=True)),
    (2, 64, 80, 4, dict(conf_thres=0.999, iou_thres=0.45)),  # nothing passes
]

LOSS_CASES = [  # (name, nc, strides, img, use_dfl, reg_max, iou_type, warmup_epoch, epoch, B, seed)
    ("s_tal_giou", 80, [8, 16, 32], 320, False, 0, "giou", 0, 0, 4, 0),
    ("n_tal_siou", 80, [8, 16, 32], 320, False, 0, "siou", 0, 0, 4, 1),
    ("m_tal_dfl", 80, [8, 16, 32], 320, True, 16, "giou", 0, 0, 4, 2),
    ("l6_atss_dfl", 20, [8, 16, 32, 64], 512, True, 16, "giou", 4, 0, 2, 3),
    ("l6_tal_dfl", 20, [8, 16, 32, 64], 512, True, 16, "giou", 4, 5, 2, 3),
    # ragged / empty targets (12th field: which images lose their boxes)
    ("s_tal_one_empty_image", 80, [8, 16, 32], 320, False, 0, "giou", 0, 0, 4, 5, "img1"),
    ("s_tal_no_targets", 80, [8, 16, 32], 320, False, 0, "giou", 0, 0, 2, 6, "all"),
    ("l6_atss_one_empty_image", 20, [8, 16, 32, 64], 512, True, 16, "giou", 4, 0, 2, 7, "img1"),
]


def drop_targets(targets, drop):
    """Test-case helper: remove the boxes of image 1 ("img1") or of every image ("all")."""
    if drop == "img1":
        return targets[targets[:, 0] != 1]
    if drop == "all":
        return targets[:0]
    return targets


def load_cfg(name):
    cfg = Config.fromfile(f"/root/reference/configs/{name}.py")
    if not hasattr(cfg, "training_mode"):
        setattr(cfg, "training_mode", "repvgg")  # tools/train.py:99-100
    return cfg


def golden_models():
    for name, size in MODELS.items():
        m = build_model(load_cfg(name), 80, torch.device("cpu"))
        keys = [(k, list(v.shape)) for k, v in m.state_dict().items()]
        with open(os.path.join(HERE, f"keys_{name}.json"), "w") as f:
            json.dump(keys, f)
        sd = fab.fabricate_state_dict(keys, seed=0)
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = fab.synthetic_images(2, size, size, seed=0)
        with torch.no_grad():
            out_eval = m(x)[0]
            m.detect.training = True           # train branch of Detect.forward with eval-mode BN
            feats = m.neck(m.backbone(x))
            _, cls_t, reg_t = m.detect(list(feats))
            m.detect.training = False
            fuse_model(m)                       # reference deploy order: fuse BN, then re-parameterise
            for layer in m.modules():
                if isinstance(layer, RepVGGBlock):
                    layer.switch_to_deploy()
            out_deploy = m(x)[0]
        np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), eval_out=out_eval.numpy(),
                            cls_train=cls_t.numpy(), reg_train=reg_t.numpy(), deploy_out=out_deploy.numpy(),
                            x_checksum=np.float64(fab.checksum(x)),
                            w_checksum=np.float64(sum(fab.checksum(v) for v in sd.values())))
        print(name, tuple(out_eval.shape), "deploy drift", (out_eval - out_deploy).abs().max().item())


def golden_nms():
    store = {}
    for i, (B, A, nc, seed, kw) in enumerate(NMS_CASES):
        p = fab.synthetic_predictions(B, A, nc, seed)
        out = non_max_suppression(p.clone(), **kw)
        store[f"c{i}_checksum"] = np.float64(fab.checksum(p))
        store[f"c{i}_counts"] = np.array([o.shape[0] for o in out], dtype=np.int64)
        store[f"c{i}_rows"] = torch.cat(out).numpy() if sum(o.shape[0] for o in out) else np.zeros((0, 6), np.float32)
        print("nms case", i, [o.shape[0] for o in out])
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **store)
    with open(os.path.join(HERE, "nms_cases.json"), "w") as f:
        json.dump([[B, A, nc, seed, kw] for B, A, nc, seed, kw in NMS_CASES], f)


def golden_loss():
    store = {}
    for case in LOSS_CASES:
        (name, nc, strides, img, use_dfl, reg_max, iou_type, warm, epoch, B, seed), drop = case[:11], (case[11] if len(case) > 11 else None)
        sizes = [(img // s, img // s) for s in strides]
        ps, pd = fab.synthetic_head_outputs(B, sizes, nc, 4 * (reg_max + 1), seed)
        ps.requires_grad_(True)
        pd.requires_grad_(True)
        targets = drop_targets(oloss.synthetic_targets(B, seed=seed + 1, num_classes=nc), drop)
        feats = [torch.zeros(B, 8, h, w) for h, w in sizes]
        cl = ComputeLoss(fpn_strides=strides, num_classes=nc, ori_img_size=img, warmup_epoch=warm, use_dfl=use_dfl,
                         reg_max=reg_max, iou_type=iou_type)
        # capture the assigner outputs the reference produced
        captured = {}
        for attr in ("warmup_assigner", "formal_assigner"):
            mod = getattr(cl, attr)
            orig = mod.forward

            def wrap(*a, _orig=orig, **k):
                r = _orig(*a, **k)
                captured["out"] = r
                return r
            mod.forward = wrap
        loss, items = cl((feats, ps, pd), targets.clone(), epoch, 1, img, img)
        g_ps, g_pd = torch.autograd.grad(loss, [ps, pd], allow_unused=True)
        tl, tb, ts, fg = captured["out"]
        fg = fg.bool()      # the no-target early return of the assigners yields a float zero mask (tal_assigner.py:41-46)
        nz = ts.nonzero()
        store[f"{name}_loss"] = np.float64(loss.item())
        store[f"{name}_items"] = items.double().numpy()
        store[f"{name}_in_checksum"] = np.float64(fab.checksum(ps) + fab.checksum(pd) + fab.checksum(targets))
        store[f"{name}_labels"] = tl.numpy().astype(np.int16)
        store[f"{name}_fg"] = np.packbits(fg.numpy())
        store[f"{name}_bboxes_fg"] = tb[fg].double().numpy()          # boxes of positives (pixels, pre /stride)
        store[f"{name}_scores_idx"] = nz.numpy().astype(np.int32)
        store[f"{name}_scores_val"] = ts[ts != 0].double().numpy()
        store[f"{name}_grad_scores_sum"] = np.float64(g_ps.double().sum().item())
        store[f"{name}_grad_scores_abs"] = np.float64(g_ps.double().abs().sum().item())
        store[f"{name}_grad_scores_head"] = g_ps.flatten()[:4096].double().numpy()
        store[f"{name}_grad_scores_fg"] = g_ps[fg].double().numpy()
        if g_pd is not None:
            store[f"{name}_grad_distri_abs"] = np.float64(g_pd.double().abs().sum().item())
            store[f"{name}_grad_distri_fg"] = g_pd[fg].double().numpy()
        print(name, "loss", loss.item(), "items", items.tolist(), "pos", int(fg.sum()))
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **store)
    with open(os.path.join(HERE, "loss_cases.json"), "w") as f:
        json.dump(LOSS_CASES, f)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "--loss-only" not in sys.argv:
        golden_models()
        golden_nms()
    golden_loss()
    print("golden vectors written to", HERE)
