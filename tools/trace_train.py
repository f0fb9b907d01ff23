"""Kernel-time breakdown of one training step (CUPTI via torch.profiler): totals per kernel name, GPU busy vs span.
usage: python tools/trace_train.py [model] [batch] [size]"""
import collections
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200.loss import ComputeLoss  # noqa: E402
from yolov6_b200.model import build_model  # noqa: E402
from yolov6_b200.synth import randomize_  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolov6s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
dev = torch.device("cuda:0")
m = randomize_(build_model(name, 80, dev), seed=0)
m.detect.initialize_biases()
m.train()
g = torch.Generator().manual_seed(1)
x = torch.rand(B, 3, S, S, generator=g).to(dev)
n = 7 * B
wh = torch.rand(n, 2, generator=g) * 0.58 + 0.02
cxy = wh / 2 + torch.rand(n, 2, generator=g) * (1 - wh)
targets = torch.cat([torch.arange(n).remainder(B).float().view(-1, 1), torch.randint(0, 80, (n, 1), generator=g).float(), cxy, wh], 1).to(dev)
crit = ComputeLoss(num_classes=80, ori_img_size=S, warmup_epoch=0, use_dfl=False, reg_max=0, iou_type="giou")


def step():
    for p in m.parameters():
        p.grad = None
    preds, _ = m(x)
    loss, _ = crit(preds, targets, 1, 0, S, S)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace_train.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
agg = collections.OrderedDict()
for e in ev:
    k = e["name"].split("(")[0].replace("void ", "").replace("yv6::", "")[:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e["dur"]
busy = sum(a[1] for a in agg.values())
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
print(f"# {name} bs{B} {S}x{S}: one training step (fwd + TAL + loss + bwd), eager launches\n")
print(f"GPU span {span / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms ({100 * busy / span:.0f} %), {len(ev)} GPU activities\n")
print("| kernel | launches | total ms | share of busy |")
print("|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"| {k} | {a[0]} | {a[1] / 1e3:.2f} | {100 * a[1] / busy:.1f}% |")
