"""Epilogue warp groups (2 vs 4 TMEM accumulators / epilogue groups per CTA) on the N <= 128 3x3 layers of YOLOv6-S bs32.
cold = L2 flushed before each launch (median of 10), warm = 40 back-to-back launches.
usage (GPU box): python tools/groups_sweep.py > gpurun_out/groups_sweep.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = 32
# (H = W of the input, Cin, Cout, k, stride)
SHAPES = [(160, 64, 64, 3, 1), (80, 64, 64, 3, 1), (80, 128, 128, 3, 1), (40, 128, 128, 3, 1), (80, 64, 128, 3, 1), (160, 64, 64, 3, 2),
          (160, 64, 128, 3, 2), (80, 128, 128, 3, 2), (160, 64, 64, 1, 1), (80, 128, 128, 1, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return ts[len(ts) // 2] * 1e3, e0.elapsed_time(e1) / 40 * 1e3


print("| in HxW | Cin | Cout | k | s | act | G=2 cold us | G=4 cold us | G=2 warm us | G=4 warm us | outputs |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for hw, cin, cout, k, st in SHAPES:
    for act in ("relu", "silu"):
        xb = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
        w = torch.randn(cout, k, k, cin, device=dev) / (k * k * cin) ** 0.5
        bias = ops.pad_bias(torch.zeros(cout, device=dev), cout)
        ho = hw // st
        y = torch.empty(B, ho, ho, cout, dtype=torch.bfloat16, device=dev)
        if st == 2:
            xin, wb = xb.view(B, hw, hw // 2, 2 * cin), ops.pair_view_weights(w).to(torch.bfloat16)
            kw = dict(stride=2, stride_w=1, pad=(1, 1), out_hw=(0, ho), pair_view=1, act=act)
        else:
            xin, wb = xb, w.to(torch.bfloat16)
            kw = dict(stride=1, act=act)
        res = {}
        for g in (2, 4):
            try:
                res[g] = timed(lambda: ops.conv_fwd(xin, wb, bias, y, force=dict(groups=g), **kw))
                res[(g, "y")] = y.clone()
            except Exception as e:  # noqa: BLE001
                res[g] = (float("nan"), float("nan"))
                res[(g, "y")] = None
                print("ERROR", str(e)[:100], file=sys.stderr)
        same = "same" if res[(2, "y")] is not None and res[(4, "y")] is not None and bool((res[(2, "y")] == res[(4, "y")]).all()) else "DIFFERENT"
        print(f"| {hw}x{hw} | {cin} | {cout} | {k} | {st} | {act} | {res[2][0]:.1f} | {res[4][0]:.1f} | {res[2][1]:.1f} | {res[4][1]:.1f} | {same} |", flush=True)
