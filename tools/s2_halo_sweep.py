"""3x3 stride-2 conv layers of YOLOv6-S bs32: plain stride-2 generic mainloop vs the column-pair view (include/yv6.h `pair_view`)
with the halo-reuse mainloop (MODE 3 / 4 of yv6_conv_igemm.cu), single CTA and CTA pairs.  cold = L2 flushed before each launch
(median of 10), warm = 40 back-to-back launches.  usage (GPU box): python tools/s2_halo_sweep.py > gpurun_out/s2_halo_sweep.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = 32
SHAPES = [(320, 32, 64), (160, 64, 128), (80, 128, 256), (40, 256, 512), (160, 64, 64), (80, 128, 128), (80, 64, 64), (40, 128, 128)]
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


print("| in HxW | Cin | Cout | variant | cold us | warm us | warm TFLOP/s | plan |")
print("|---|---|---|---|---|---|---|---|")
for hw, cin, cout in SHAPES:
    xb = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, 3, 3, cin, device=dev) / (9 * cin) ** 0.5
    wb, wv = w.to(torch.bfloat16), ops.pair_view_weights(w).to(torch.bfloat16)
    bias = ops.pad_bias(torch.zeros(cout, device=dev), cout)
    ho = hw // 2
    y = torch.empty(B, ho, ho, cout, dtype=torch.bfloat16, device=dev)
    xv = xb.view(B, hw, hw // 2, 2 * cin)
    fl = 2.0 * B * ho * ho * cout * cin * 9
    view = dict(stride=2, stride_w=1, pad=(1, 1), out_hw=(0, ho), pair_view=1, act="relu")
    variants = [("plain stride 2", lambda f: ops.conv_fwd(xb, wb, bias, y, stride=2, act="relu", force=f), dict(pair=-1), None),
                ("pair view, generic", lambda f: ops.conv_fwd(xv, wv, bias, y, force=f, **view), dict(pair=-1, halo=-1), view),
                ("pair view, halo", lambda f: ops.conv_fwd(xv, wv, bias, y, force=f, **view), dict(pair=-1, halo=1), view),
                ("pair view, halo, CTA pairs", lambda f: ops.conv_fwd(xv, wv, bias, y, force=f, **view), dict(pair=1, halo=1), view),
                ("pair view, auto", lambda f: ops.conv_fwd(xv, wv, bias, y, force=f, **view), dict(), view)]
    y0 = None
    for name, fn, force, v in variants:
        try:
            c, wm = timed(lambda: fn(force))
            if y0 is None:
                y0 = y.clone()
                same = "reference"
            else:
                same = "same" if bool((y0 == y).all()) else f"max diff {float((y0.float() - y.float()).abs().max()):.3g}"
            if v is None:
                plan = ops.conv_plan((B, hw, hw, cin), (cout, 3, 3, cin), 2, 1, force)
            else:
                plan = ops.conv_plan((B, hw, hw // 2, 2 * cin), (cout, 3, 2, 2 * cin), 2, 1, force, stride_w=1, pad=(1, 1), out_hw=(0, ho), pair_view=1)
            print(f"| {hw}x{hw} | {cin} | {cout} | {name} | {c:.1f} | {wm:.1f} | {fl / wm / 1e6:.0f} | {plan['BW']}x{plan['BH']} BN {plan['BN']} "
                  f"tiles {plan['tiles']} halo {plan['halo']} mode {plan['a_res']} {same} |", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"| {hw}x{hw} | {cin} | {cout} | {name} | ERROR {str(e)[:80]} |", flush=True)
