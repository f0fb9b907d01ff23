"""CTA-pair kernels (tcgen05 cta_group::2) against the single-CTA kernels on the conv shapes of YOLOv6-S bs32:
cold (L2 flushed before each launch, median of 10) and warm (40 back-to-back launches).
usage (GPU box): python tools/pair_sweep.py > gpurun_out/pair_sweep.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = 32
# (H = W of the input, Cin, Cout, k, stride)
SHAPES = [(80, 128, 128, 3, 1), (40, 256, 256, 3, 1), (20, 512, 512, 3, 1), (160, 64, 64, 3, 1), (40, 128, 128, 3, 1),
          (20, 256, 256, 3, 1), (80, 64, 64, 3, 1), (80, 64, 128, 3, 1), (40, 128, 256, 3, 1), (20, 256, 512, 3, 1),
          (160, 64, 128, 3, 2), (80, 128, 256, 3, 2), (40, 256, 512, 3, 2), (160, 64, 64, 1, 1), (80, 128, 128, 1, 1),
          (40, 384, 128, 1, 1), (20, 512, 256, 1, 1), (20, 1024, 256, 1, 1), (80, 64, 80, 1, 1), (20, 128, 128, 1, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_conv(xb, wb, bias, y, st, force):
    for _ in range(3):
        ops.conv_fwd(xb, wb, bias, y, stride=st, act="relu", force=force)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_fwd(xb, wb, bias, y, stride=st, act="relu", force=force)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        ops.conv_fwd(xb, wb, bias, y, stride=st, act="relu", force=force)
    e1.record()
    torch.cuda.synchronize()
    return ts[len(ts) // 2], e0.elapsed_time(e1) / 40


print("| in HxW | Cin | Cout | k | s | single cold us | pair cold us | single warm us | pair warm us | warm TFLOP/s single -> pair | plan (pair) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for hw, cin, cout, k, st in SHAPES:
    xb = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    wb = (torch.randn(cout, k, k, cin, device=dev) / (k * k * cin) ** 0.5).to(torch.bfloat16)
    bias = ops.pad_bias(torch.zeros(cout, device=dev), cout)
    ho = (hw + 2 * (k // 2) - k) // st + 1
    y = torch.empty(B, ho, ho, cout, dtype=torch.bfloat16, device=dev)
    fl = 2.0 * B * ho * ho * cout * cin * k * k
    try:
        c0, w0 = time_conv(xb, wb, bias, y, st, dict(pair=-1))
        y0 = y.clone()
        c1, w1 = time_conv(xb, wb, bias, y, st, dict(pair=1))
        same = bool((y0 == y).all())
        plan = ops.conv_plan((B, hw, hw, cin), (cout, k, k, cin), st, 1, dict(pair=1))
        print(f"| {hw}x{hw} | {cin} | {cout} | {k} | {st} | {c0 * 1e3:.1f} | {c1 * 1e3:.1f} | {w0 * 1e3:.1f} | {w1 * 1e3:.1f} | "
              f"{fl / w0 / 1e9:.0f} -> {fl / w1 / 1e9:.0f} | {plan['BW']}x{plan['BH']}x{plan['BI']} BN {plan['BN']} units {plan['tiles']} grid {plan['grid']} "
              f"mode {plan['a_res']} {'same' if same else 'DIFFERENT'} |", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"| {hw}x{hw} | {cin} | {cout} | {k} | {st} | ERROR {e!r} |", flush=True)
