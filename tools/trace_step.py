"""In-situ kernel timeline of one CUDA-graph step (CUPTI via torch.profiler; no cache flush, no replay): per-kernel duration and
the idle gap before it.  The per-layer table is taken with the engine's launches on ONE stream (YV6_LANES=1), where start order =
launch order and every kernel can be named from the engine's call list; the span of the shipped multi-stream schedule (graph
branches for BiFusion / CSPSPPF / head levels) is measured next to it.  usage: python tools/trace_step.py [model] [batch] [size]"""
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


def trace(lanes):
    """Kernel / memset events of the last of three replays of the serving graph, with the engine spread over `lanes` streams."""
    if lanes is None:
        os.environ.pop("YV6_LANES", None)
    else:
        os.environ["YV6_LANES"] = str(lanes)
    torch.manual_seed(0)
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
    path = os.path.join(tempfile.gettempdir(), f"trace_{lanes}.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    step = ev[2 * (len(ev) // 3):]                      # last step = last third of the events
    plan = m.engine()._plan(B, S, S, torch.float32)
    names, ci = [], iter(plan["conv_info"])
    for kind, _ in plan["calls"]:
        if kind == "conv":
            c = next(ci)
            names.append(f"{c['name']} {c['cin']}->{c['cout']} k{c['k']}s{c['s']} @{c['ho']}x{c['wo']}")
        elif kind == "stem":
            names.append("backbone.stem k3s2")
        elif kind == "pool":
            names.append("cspsppf / sppf max-pool x3")
        else:
            names.append(kind)
    return step, names


def span_busy(step):
    t0 = step[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in step)
    return end - t0, sum(e["dur"] for e in step)


step, names = trace(1)
t0 = step[0]["ts"]
prev_end = t0
print(f"# {name} bs{B} {S}x{S}: in-situ kernel timeline of one graph step (CUPTI), launches on one stream\n")
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
    prev_end = e["ts"] + e["dur"]
sp, busy = span_busy(step)
print(f"\nsingle stream: step span {sp:.1f} us, kernel time {busy:.1f} us, idle {sp - busy:.1f} us over {len(step)} events")
step_m, _ = trace(None)
spm, busym = span_busy(step_m)
print(f"shipped schedule (graph branches over 4 streams): step span {spm:.1f} us, summed kernel time {busym:.1f} us "
      f"({busym - spm:.1f} us of it overlapped) over {len(step_m)} events")
