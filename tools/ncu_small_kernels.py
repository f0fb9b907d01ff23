"""Eager run of one YOLOv6-S bs32 forward + head-tensor NMS (twice) for `ncu -k regex:...` captures of the non-conv kernels.
usage: ncu --set full --clock-control none -k regex:'stem_mma|sppf_pool|nms_' -s 8 -c 8 -o out python tools/ncu_small_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200.model import build_model  # noqa: E402
from yolov6_b200.nms import nms_batched_head  # noqa: E402
from yolov6_b200.synth import randomize_  # noqa: E402

dev = torch.device("cuda:0")
m = randomize_(build_model("yolov6s", 80, dev), seed=0).eval()
eng = m.engine()
eng.n_lanes = 1
x = torch.rand(32, 3, 640, 640, generator=torch.Generator().manual_seed(1)).to(dev)
kw = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
with torch.no_grad():
    for _ in range(2):
        cls, reg, sizes = eng.forward(x, decode=False)
        nms_batched_head(cls, reg, sizes, m.graph.strides, **kw)
torch.cuda.synchronize()
