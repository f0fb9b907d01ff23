"""Pair mode (two M tiles per weight tile, conv kernel mode 3) against the single-tile halo mode on the 3x3 stride-1
shapes of YOLOv6-S/M bs32: cold (L2 flushed before each launch, median of 10) and warm (40 back-to-back launches).
usage (GPU box): python tools/pair_sweep.py > gpurun_out/pair_sweep.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = 32
SHAPES = [(80, 128, 128), (40, 256, 256), (40, 128, 128), (20, 256, 256), (20, 512, 512), (80, 64, 128), (40, 128, 256),
          (20, 256, 512), (80, 192, 192), (40, 384, 384), (160, 64, 64)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_conv(xb, wb, bias, y, force):
    for _ in range(3):
        ops.conv_fwd(xb, wb, bias, y, act="relu", force=force)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_fwd(xb, wb, bias, y, act="relu", force=force)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        ops.conv_fwd(xb, wb, bias, y, act="relu", force=force)
    e1.record()
    torch.cuda.synchronize()
    return ts[len(ts) // 2], e0.elapsed_time(e1) / 40


print("| HxW | Cin | Cout | single cold us | pair cold us | single warm us | pair warm us | warm TFLOP/s single -> pair | auto plan |")
print("|---|---|---|---|---|---|---|---|---|")
for hw, cin, cout in SHAPES:
    xb = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    wb = (torch.randn(cout, 3, 3, cin, device=dev) / (9 * cin) ** 0.5).to(torch.bfloat16)
    bias = ops.pad_bias(torch.zeros(cout, device=dev), cout)
    y = torch.empty(B, hw, hw, cout, dtype=torch.bfloat16, device=dev)
    fl = 2.0 * B * hw * hw * cout * cin * 9
    try:
        c0, w0 = time_conv(xb, wb, bias, y, dict(pair=-1))
        y0 = y.clone()
        c1, w1 = time_conv(xb, wb, bias, y, dict(pair=1, halo=1))
        same = bool((y0 == y).all())
        plan = ops.conv_plan((B, hw, hw, cin), (cout, 3, 3, cin), 1, 1, None)
        print(f"| {hw}x{hw} | {cin} | {cout} | {c0 * 1e3:.1f} | {c1 * 1e3:.1f} | {w0 * 1e3:.1f} | {w1 * 1e3:.1f} | "
              f"{fl / w0 / 1e9:.0f} -> {fl / w1 / 1e9:.0f} | BN {plan['BN']} tiles {plan['tiles']} mode {plan['a_res']} {'same' if same else 'DIFFERENT'} |", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"| {hw}x{hw} | {cin} | {cout} | ERROR {e!r} |", flush=True)
