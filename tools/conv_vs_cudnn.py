"""Per-layer comparison of yv6_conv_fwd (bias + activation fused) with cuDNN's bf16 convolution (torch F.conv2d,
channels_last, cudnn.benchmark, conv only -- bias / activation would be extra kernels) on the layer shapes of a model.
TOOLS ONLY: cuDNN is the library baseline the kernels are measured against, it is not on the product path.
usage: python tools/conv_vs_cudnn.py [model] [batch] [size] > profiles/rNN_conv_vs_cudnn.md"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200.model import build_model  # noqa: E402
from yolov6_b200.synth import randomize_  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolov6s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
m = randomize_(build_model(name, 80, dev)).eval()
x = torch.rand(B, 3, S, S, device=dev)
rows = m.engine().profile_layers(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def cudnn_ms(cin, cout, k, s, h, w):
    xi = torch.randn(B, cin, h, w, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, k, k, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        F.conv2d(xi, wt, stride=s, padding=k // 2)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        F.conv2d(xi, wt, stride=s, padding=k // 2)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


plan = m.engine()._plan(B, S, S, torch.float32)
seen = {}
tot_o = tot_c = 0.0
print(f"# yv6_conv_fwd vs cuDNN bf16 (conv only): {name} bs{B} {S}x{S}, L2 flushed before each launch, median of 10\n")
print("| layer | Cin | Cout | k | s | HxW | ours us | TFLOP/s | cuDNN us | ours / cuDNN |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r, ci in zip(rows, plan["conv_info"]):
    key = (ci["cin"], r["cout"], ci["k"], ci["s"], ci["h"], ci["w"])
    if key not in seen:
        seen[key] = cudnn_ms(ci["cin"], r["cout"], ci["k"], ci["s"], ci["h"], ci["w"])
    c = seen[key]
    tot_o += r["ms"]
    tot_c += c
    print(f"| {r['name']} | {ci['cin']} | {r['cout']} | {ci['k']} | {ci['s']} | {r['hw']} | {r['ms'] * 1e3:.1f} | {r['tflops']:.0f} | {c * 1e3:.1f} | {r['ms'] / c:.2f} |")
print(f"\ntotal: ours {tot_o:.3f} ms, cuDNN {tot_c:.3f} ms (ratio {tot_o / tot_c:.2f}); cuDNN figures exclude bias / activation / residual / "
      "concat-slice writes, which ours fuse")
