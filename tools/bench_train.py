"""Training-step timing: YOLOv6-S train form, one full step = forward + TAL assignment +
VFL/IoU loss + backward (no optimiser, no dataloader) on synthetic COCO-shaped targets.  Prints one JSON line.
usage: python tools/bench_train.py [model] [batch] [size] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200.loss import ComputeLoss  # noqa: E402
from yolov6_b200.model import build_model  # noqa: E402
from yolov6_b200.synth import randomize_  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolov6s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
m = randomize_(build_model(name, 80, dev), seed=0)
m.detect.initialize_biases()
m.train()
g = torch.Generator().manual_seed(1)
x = torch.rand(B, 3, S, S, generator=g).to(dev)
rows = []
for b in range(B):                       # n ~ clip(Poisson(7.3), 1, 60) boxes per image
    n = int(torch.poisson(torch.tensor([7.3]), generator=g).clamp(1, 60).item())
    wh = torch.rand(n, 2, generator=g) * 0.58 + 0.02
    cxy = wh / 2 + torch.rand(n, 2, generator=g) * (1 - wh)
    rows.append(torch.cat([torch.full((n, 1), float(b)), torch.randint(0, 80, (n, 1), generator=g).float(), cxy, wh], 1))
targets = torch.cat(rows).to(dev)
crit = ComputeLoss(num_classes=80, ori_img_size=S, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type="giou")


def step():
    for p in m.parameters():
        p.grad = None
    preds, _ = m(x)
    loss, _ = crit(preds, targets, 1, 0, S, S)
    loss.backward()
    return loss


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"metric": f"train-step images/sec {name} {S} bs{B} (fwd + TAL + loss + bwd, eager launches)", "value": B / (ms * 1e-3),
                  "unit": "images/s", "ms_per_step": ms, "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
