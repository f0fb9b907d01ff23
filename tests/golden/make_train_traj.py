"""Writes tests/golden/train_traj.json: the loss trajectory of 8 SGD steps of the ORACLE (oracle/model.py
train-mode network + oracle/loss.py, float64, torch autograd on CPU) on one fixed synthetic batch, for
the ATSS warm-up branch (epoch 0) and the TAL branch (epoch 5) of ComputeLoss (loss.py:86-123).

The GPU test replays the same steps through yolov6_b200's Model / ComputeLoss / TrainEngine
(tests/test_gpu_train.py::test_training_steps_follow_the_oracle_trajectory).

    python tests/golden/make_train_traj.py        (about 3 minutes of CPU)
"""
import json
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import golden_keys          # noqa: E402
from oracle import fabricate as fab      # noqa: E402
from oracle import model as om           # noqa: E402
from oracle.loss import compute_loss, synthetic_targets   # noqa: E402

NAME, SIZE, BATCH, STEPS, LR = "yolov6n", 160, 4, 8, 0.02


def initial_state():
    """Fabricated backbone / neck / head convs; prediction convs as Detect.initialize_biases leaves them
    (effidehead.py:49-65)."""
    sd = fab.fabricate_state_dict(golden_keys(NAME), seed=0)
    for k in sd:
        if k.startswith("detect.") and "_preds." in k:
            if k.endswith("weight"):
                sd[k] = torch.zeros_like(sd[k])
            else:
                sd[k] = torch.full_like(sd[k], -math.log((1 - 1e-2) / 1e-2) if "cls_preds" in k else 1.0)
    return sd


def run(epoch):
    sd = initial_state()
    x = fab.synthetic_images(BATCH, SIZE, SIZE, seed=3).double()
    targets = synthetic_targets(BATCH, seed=2)
    params = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k and "proj" not in k else v)
              for k, v in sd.items()}
    opt = torch.optim.SGD([v for v in params.values() if torch.is_tensor(v) and v.requires_grad], lr=LR, momentum=0.9, nesterov=True)
    cfg = om.CONFIGS[NAME]
    sizes = [(SIZE // s, SIZE // s) for s in cfg["strides"]]
    out = []
    for step in range(STEPS):
        opt.zero_grad()
        with om.train_mode():
            cls, reg, _ = om.forward(params, cfg, x, train_outputs=True)
        loss, items = compute_loss(sizes, cls, reg, targets, strides=cfg["strides"], ori_img_size=SIZE, warmup_epoch=4,
                                   epoch_num=epoch, use_dfl=False, reg_max=0, iou_type="siou")
        loss.backward()
        opt.step()
        out.append(dict(loss=float(loss.detach()), items=[float(i) for i in items]))
        print(epoch, step, out[-1], flush=True)
    return out


if __name__ == "__main__":
    res = dict(name=NAME, size=SIZE, batch=BATCH, lr=LR, momentum=0.9, nesterov=True, image_seed=3, target_seed=2,
               trajectories={str(e): run(e) for e in (0, 5)})
    with open(os.path.join(HERE, "train_traj.json"), "w") as f:
        json.dump(res, f, indent=1)
