"""In-situ kernel timeline of one CUDA-graph step (CUPTI via torch.profiler; no cache flush, no replay):
per-kernel duration and the idle gap before it.  usage: python tools/trace_step.py [model] [batch] [size]"""
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200.model import build_model  # noqa: E402
from yolov6_b200.pipeline import DetectPipeline  # noqa: E402
from yolov6_b200.synth import randomize_  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolov6s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
dev = torch.device("cuda:0")
m = randomize_(build_model(name, 80, dev)).eval()
pipe = DetectPipeline(m, B, S, S, host_input=False, conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
pipe.x_dev.copy_(torch.rand(B, 3, S, S, device=dev))
for _ in range(5):
    pipe.launch()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        pipe.launch()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# last step = last third of the events
n = len(ev) // 3
step = ev[2 * n:]
names = []
g = m.graph
for op in g.ops:
    if op.kind == "convT":
        names += [f"{op.name}[{q}] {op.cin}->{op.cout}" for q in range(4)]
    elif op.kind == "pool":
        names.append(op.name + " pool")
    else:
        names.append(f"{op.name} {op.cin}->{op.cout} k{op.k}s{op.s}")
t0 = step[0]["ts"]
prev_end = t0
busy = 0.0
print(f"# {name} bs{B} {S}x{S}: in-situ kernel timeline of one graph step (CUPTI)\n")
print("| # | kernel | start us | dur us | gap before us | layer |")
print("|---|---|---|---|---|---|")
ki = 0
for i, e in enumerate(step):
    nm = e["name"].split("(")[0].replace("void ", "").replace("yv6::", "")
    is_model = e.get("cat") == "kernel" and ki < len(names) and ("conv_igemm" in nm or "stem" in nm or "sppf" in nm)
    layer = names[ki] if is_model else ""
    if is_model:
        ki += 1
    print(f"| {i} | {nm[:28]} | {e['ts'] - t0:.1f} | {e['dur']:.1f} | {e['ts'] - prev_end:.1f} | {layer} |")
    busy += e["dur"]
    prev_end = e["ts"] + e["dur"]
print(f"\nstep span {prev_end - t0:.1f} us, busy {busy:.1f} us, idle {prev_end - t0 - busy:.1f} us over {len(step)} events")
