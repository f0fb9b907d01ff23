"""Per-layer conv timing table for a model (writes markdown to stdout / a file). GPU only.
usage: python tools/profile_layers.py [model] [batch] [size] > profiles/rNN_conv_layers.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200.model import build_model  # noqa: E402
from yolov6_b200.synth import randomize_  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolov6s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
dev = torch.device("cuda:0")
m = randomize_(build_model(name, 80, dev)).eval()
x = torch.rand(B, 3, S, S, device=dev)
rows = m.engine().profile_layers(x)
tot = sum(r["ms"] for r in rows)
print(f"# conv_igemm per-launch timings: {name} bs{B} {S}x{S}, bf16, L2 flushed before each launch, median of 10\n")
print(f"total {tot:.3f} ms over {len(rows)} launches; {sum(r['tflops'] * r['ms'] for r in rows) / tot:.0f} TFLOP/s average\n")
print("| layer | Cin | Cout | k | s | HxW | ms | TFLOP/s | GB/s | tile BWxBHxBI BN KB stg halo |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    p = r["plan"]
    print(f"| {r['name']} | {r['cin']} | {r['cout']} | {r['k']} | {r['s']} | {r['hw']} | {r['ms']:.4f} | {r['tflops']:.0f} | {r['gbs']:.0f} | "
          f"{p[0]}x{p[1]}x{p[2]} {p[3]} {p[4]} {p[5]} {p[8]}/{p[9]} |")
