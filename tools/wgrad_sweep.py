"""Per-shape timing of yv6_conv_wgrad for the layer shapes of a training step, with the split-K factor forced to a few
values (separates mainloop time from the cost of the fp32 reduction epilogue).
usage: python tools/wgrad_sweep.py [s32|m8]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import _lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "s32"
print("YV6_WGRAD_FLAGS =", os.environ.get("YV6_WGRAD_FLAGS", "0"))
N = 32 if which == "s32" else 8
SHAPES = {  # (H, Cin, Cout, k, s)
    "s32": [(160, 64, 64, 3, 1), (160, 64, 64, 1, 1), (80, 128, 128, 3, 1), (40, 256, 256, 3, 1), (20, 512, 512, 3, 1), (20, 256, 256, 3, 1),
            (320, 32, 64, 3, 2), (160, 64, 128, 3, 2), (80, 128, 256, 3, 2), (40, 384, 128, 1, 1), (20, 1024, 256, 1, 1), (80, 64, 80, 1, 1)],
    "m8": [(160, 96, 96, 3, 1), (80, 128, 128, 3, 1), (40, 256, 256, 3, 1), (20, 512, 512, 3, 1), (40, 192, 384, 3, 2), (20, 1536, 768, 1, 1),
           (80, 192, 192, 1, 1)],
}[which]
dev = torch.device("cuda:0")
lib, h = _lib.lib(), _lib.handle(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"| H | Cin | Cout | k | s | GFLOP | " + " | ".join(f"ks={k} us (TF/s)" for k in ("auto", 4, 16)) + " |")
print("|" + "---|" * 9)
for (H, ci, co, k, s) in SHAPES:
    x = torch.randn(N, H, H, ci, device=dev).to(torch.bfloat16)
    Ho = (H + 2 * (k // 2) - k) // s + 1
    dy = torch.randn(N, Ho, Ho, co, device=dev).to(torch.bfloat16)
    dw = torch.zeros(co, k, k, ci, dtype=torch.float32, device=dev)
    fl = 2.0 * N * Ho * Ho * co * ci * k * k
    cells = []
    for ks in (0, 4, 16):
        d = _lib.WgradDesc()
        d.x, d.N, d.H, d.W, d.Cin, d.x_c_total = x.data_ptr(), N, H, H, ci, ci
        d.dy, d.Cout, d.dy_c_total = dy.data_ptr(), co, co
        d.kh = d.kw = k
        d.stride, d.pad, d.dw, d.force_ksplit = s, k // 2, dw.data_ptr(), ks
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.yv6_conv_wgrad(h, C.byref(d), _lib.stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        cells.append(f"{ms * 1e3:.0f} ({fl / ms / 1e9:.0f})")
    print(f"| {H} | {ci} | {co} | {k} | {s} | {fl / 1e9:.1f} | " + " | ".join(cells) + " |")
