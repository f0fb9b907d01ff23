"""Phase timeline of conv_igemm's CTA 0 (clock64 stamps written by the kernel's debug trace):
entry -> prologue done -> first TMA issued -> first operands landed -> tile-0 MMAs committed ->
accumulator visible to the epilogue -> epilogue of tile 0 done -> stores drained -> TMEM freed.
usage: python tools/trace_conv.py   (GPU)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [  # name, N, H, W, Cin, Cout, k, stride, act, out_f32
    ("1x1 256->256 @20 relu", 32, 20, 20, 256, 256, 1, 1, "relu", False),
    ("1x1 256->256 @20 silu", 32, 20, 20, 256, 256, 1, 1, "silu", False),
    ("1x1 128->128 @20 none", 32, 20, 20, 128, 128, 1, 1, None, False),
    ("1x1 256->4 @20 f32", 32, 20, 20, 256, 4, 1, 1, None, True),
    ("1x1 256->80 @20 sigmoid f32", 32, 20, 20, 256, 80, 1, 1, "sigmoid", True),
    ("3x3 256->256 @20 relu", 32, 20, 20, 256, 256, 3, 1, "relu", False),
    ("3x3 128->128 @40 relu", 32, 40, 40, 128, 128, 3, 1, "relu", False),
    ("3x3 64->64 @80 relu", 32, 80, 80, 64, 64, 3, 1, "relu", False),
    ("1x1 64->64 @160 relu", 32, 160, 160, 64, 64, 1, 1, "relu", False),
    ("3x3 256->256 @40 relu", 32, 40, 40, 256, 256, 3, 1, "relu", False),
    ("3x3 64->64 @160 relu", 32, 160, 160, 64, 64, 3, 1, "relu", False),
    ("3x3 128->128 @80 relu", 32, 80, 80, 128, 128, 3, 1, "relu", False),
]
LABELS = ["prologue", "first TMA issue", "operands landed", "tile0 MMAs issued", "acc visible", "epilogue tile0",
          "stores drained", "dealloc"]
print("| layer | total us (events) | " + " | ".join(LABELS) + " | plan |")
print("|---|---|" + "---|" * (len(LABELS) + 1))
for name, N, H, W, Cin, Cout, k, st, act, f32 in SHAPES:
    x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, k, k, Cin, device=dev) / (k * k * Cin) ** 0.5).to(torch.bfloat16)
    bias = ops.pad_bias(torch.randn(Cout, device=dev) * 0.1, Cout)
    Ho, Wo = (H + 2 * (k // 2) - k) // st + 1, (W + 2 * (k // 2) - k) // st + 1
    y = torch.empty(N, Ho, Wo, Cout, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
    tr = torch.zeros(16, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.conv_fwd(x, w, bias, y, stride=st, act=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record()
        ops.conv_fwd(x, w, bias, y, stride=st, act=act, force=dict(trace=tr))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = tr.cpu().tolist()
    clk = 1.9e3  # ~MHz -> cycles per us (approximate; stamps are SM cycles)
    d = [(t[i + 1] - t[i]) / clk if t[i + 1] and t[i] else float("nan") for i in range(8)]
    # slots: 0 entry,1 prologue,2 first tma,3 landed,4 mma issued,5 acc visible,6 epi done,7 drained,8 dealloc
    rel = [(t[i] - t[0]) / clk if t[i] else float("nan") for i in range(1, 9)]
    def rel1(i):
        return (t[i] - t[0]) / clk if t[i] else float("nan")
    print(f"    local tile 4 / 8 (halo kernels): A-load issued {rel1(12):.2f} / {rel1(13):.2f}, MMAs committed {rel1(14):.2f} / {rel1(15):.2f}, "
          f"epilogue done {rel1(9):.2f} / {rel1(10):.2f}")
    plan = ops.conv_plan((N, H, W, Cin), (Cout, k, k, Cin), st)
    print(f"| {name} | {sorted(ts)[2]:.1f} | " + " | ".join(f"{v:.2f}" for v in rel) + f" | {plan} |", flush=True)
print("\n(columns = microseconds since kernel entry of CTA 0, at ~1.9 GHz)")
