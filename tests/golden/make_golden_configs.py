"""tests/golden/make_golden_configs.py -- golden vectors AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs 2, 3, 5),
minted by running the UNMODIFIED reference from /root/reference on CPU (build container only):

  * YOLOv6-S 640x640 batch 4 and YOLOv6-L6 1280x1280 batch 1 eval forwards (A = 8400 / 34000): every 16th / 32nd anchor
    row of the [B, A, 85] output plus float64 column sums over ALL rows (a checksum of everything that is not stored);
  * ComputeLoss (TAL, GIoU, no DFL = YOLOv6-S settings) at 640x640, batch 32, COCO-shaped synthetic targets:
    loss, loss_items, the foreground mask, labels / boxes / scores of the positives and the gradients at the positives;
  * non_max_suppression at B = 32, A = 8400 with the Evaler settings (conf 0.03, iou 0.65, multi_label, max_det 300) on
    sparse predictions (~6000 candidates per image, the regime of the benchmark) and, B = 4, on dense predictions
    (~390 k candidates per image, i.e. the max_nms = 30000 truncation of nms.py:90-91).

    PYTHONPATH=tests/golden/refshim:/root/reference:. python tests/golden/make_golden_configs.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]
torch.cuda.is_available = lambda: False
nn.Module.cuda = lambda self, *a, **k: self

from yolov6.models.losses.loss import ComputeLoss  # noqa: E402
from yolov6.models.yolo import build_model  # noqa: E402
from yolov6.utils.config import Config  # noqa: E402
from yolov6.utils.nms import non_max_suppression  # noqa: E402

from oracle import fabricate as fab  # noqa: E402
from oracle import loss as oloss  # noqa: E402

MODEL_CASES = [("yolov6s", 4, 640, 16), ("yolov6l6", 1, 1280, 32)]   # name, batch, size, stored row stride
EVAL_KW = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)


def load_cfg(name):
    cfg = Config.fromfile(f"/root/reference/configs/{name}.py")
    if not hasattr(cfg, "training_mode"):
        setattr(cfg, "training_mode", "repvgg")
    return cfg


def golden_models(store):
    import json
    for name, B, size, step in MODEL_CASES:
        with open(os.path.join(HERE, f"keys_{name}.json")) as f:
            keys = [(k, tuple(s)) for k, s in json.load(f)]
        m = build_model(load_cfg(name), 80, torch.device("cpu"))
        sd = fab.fabricate_state_dict(keys, seed=0)
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = fab.synthetic_images(B, size, size, seed=40)
        t0 = time.time()
        with torch.no_grad():
            out = m(x)[0]
        print(name, tuple(out.shape), f"{time.time() - t0:.1f}s")
        store[f"{name}_rows"] = out[:, ::step].numpy()
        store[f"{name}_colsum"] = out.double().sum(1).numpy()
        store[f"{name}_abs_colsum"] = out.double().abs().sum(1).numpy()
        store[f"{name}_x_checksum"] = np.float64(fab.checksum(x))


def golden_loss(store):
    B, img, strides, nc = 32, 640, [8, 16, 32], 80
    sizes = [(img // s, img // s) for s in strides]
    ps, pd = fab.synthetic_head_outputs(B, sizes, nc, 4, seed=60)
    ps.requires_grad_(True)
    pd.requires_grad_(True)
    targets = oloss.synthetic_targets(B, seed=61, num_classes=nc)
    feats = [torch.zeros(B, 8, h, w) for h, w in sizes]
    cl = ComputeLoss(fpn_strides=strides, num_classes=nc, ori_img_size=img, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type="giou")
    captured = {}
    orig = cl.formal_assigner.forward

    def wrap(*a, **k):
        r = orig(*a, **k)
        captured["out"] = r
        return r
    cl.formal_assigner.forward = wrap
    t0 = time.time()
    loss, items = cl((feats, ps, pd), targets.clone(), 0, 1, img, img)
    g_ps, g_pd = torch.autograd.grad(loss, [ps, pd])
    print("loss640", loss.item(), items.tolist(), f"{time.time() - t0:.1f}s")
    tl, tb, ts, fg = captured["out"]
    fg = fg.bool()
    store["loss640_loss"] = np.float64(loss.item())
    store["loss640_items"] = items.double().numpy()
    store["loss640_in_checksum"] = np.float64(fab.checksum(ps) + fab.checksum(pd) + fab.checksum(targets))
    store["loss640_fg"] = np.packbits(fg.numpy())
    store["loss640_labels_fg"] = tl[fg].numpy().astype(np.int16)
    store["loss640_bboxes_fg"] = tb[fg].float().numpy()
    score_fg = ts[fg]                                                 # [npos, nc]; one non-zero per row
    store["loss640_score_fg"] = score_fg.max(1).values.double().numpy()
    store["loss640_grad_scores_abs"] = np.float64(g_ps.double().abs().sum().item())
    store["loss640_grad_scores_sum"] = np.float64(g_ps.double().sum().item())
    store["loss640_grad_scores_fg_cls"] = g_ps[fg].gather(1, tl[fg].long().unsqueeze(1)).squeeze(1).double().numpy()
    store["loss640_grad_scores_rowabs"] = g_ps.double().abs().sum(-1)[:, ::64].numpy()
    store["loss640_grad_distri_fg"] = g_pd[fg].double().numpy()
    store["loss640_grad_distri_abs"] = np.float64(g_pd.double().abs().sum().item())
    print("positives", int(fg.sum()), "targets", targets.shape[0])


def golden_nms(store):
    for tag, B, gen in (("sparse", 32, fab.synthetic_predictions_sparse), ("dense", 4, fab.synthetic_predictions)):
        p = gen(B, 8400, 80, 70)
        ncand = ((p[..., 5:] * p[..., 4:5]) > EVAL_KW["conf_thres"]).sum((1, 2))
        t0 = time.time()
        out = non_max_suppression(p.clone(), **EVAL_KW)
        print("nms", tag, "candidates/img", ncand[:4].tolist(), "kept", [o.shape[0] for o in out][:4], f"{time.time() - t0:.1f}s")
        store[f"nms_{tag}_checksum"] = np.float64(fab.checksum(p))
        store[f"nms_{tag}_counts"] = np.array([o.shape[0] for o in out], dtype=np.int64)
        store[f"nms_{tag}_rows"] = torch.cat(out).numpy()
        store[f"nms_{tag}_candidates"] = ncand.numpy()


if __name__ == "__main__":
    torch.set_num_threads(8)
    store = {}
    golden_nms(store)
    golden_loss(store)
    golden_models(store)
    np.savez_compressed(os.path.join(HERE, "configs.npz"), **store)
    print("written", os.path.join(HERE, "configs.npz"), os.path.getsize(os.path.join(HERE, "configs.npz")) // 1024, "KiB")
