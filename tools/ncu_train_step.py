"""Two eager training steps of a model (the second one is the profiled one): the command ncu wraps.
usage: [ncu ...] python tools/ncu_train_step.py [model] [batch] [size]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200.loss import ComputeLoss  # noqa: E402
from yolov6_b200.model import build_model  # noqa: E402
from yolov6_b200.step import TrainStep  # noqa: E402
from yolov6_b200.synth import synthetic_targets  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolov6s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = build_model(name, 80, dev).train()
kw = dict(use_dfl=False, reg_max=0, iou_type="giou") if name in ("yolov6n", "yolov6s") else dict(use_dfl=True, reg_max=16, iou_type="giou")
crit = ComputeLoss(fpn_strides=[int(s) for s in m.graph.strides], num_classes=80, ori_img_size=S, warmup_epoch=0, **kw)
step = TrainStep(m, crit, B, S, S, in_dtype=torch.uint8, graph=False)
x = (torch.rand(B, 3, S, S) * 255).to(torch.uint8)
step.load(x, synthetic_targets(B, seed=1))
for _ in range(2):
    step.run(epoch_num=0)
torch.cuda.synchronize()
f, b = step.eng.launch_counts()
print("launches per step:", f, b)
