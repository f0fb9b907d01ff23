"""Bring-up probe for the CTA-pair conv kernels: each configuration runs in its own process (a device fault poisons the
CUDA context) and is compared with the single-CTA kernel.  usage: python tools/pair_probe.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # N, HW, Cin, Cout, k, force
    (2, 40, 128, 128, 3, {}), (4, 40, 128, 128, 3, {}), (8, 40, 128, 128, 3, {}), (16, 40, 128, 128, 3, {}), (32, 40, 128, 128, 3, {}),
    (1, 80, 128, 128, 3, {}), (2, 80, 128, 128, 3, {}), (2, 40, 64, 64, 3, {}), (2, 40, 256, 256, 3, {}), (1, 24, 128, 128, 3, {}),
    (2, 40, 128, 128, 3, {"grid": 4}), (2, 40, 128, 128, 3, {"grid": 16}), (32, 40, 128, 128, 3, {"grid": 60}), (2, 40, 128, 128, 3, {"halo": -1}),
    (2, 40, 128, 128, 3, {"direct": 2}), (2, 40, 128, 128, 3, {"direct": 1}), (2, 40, 128, 128, 1, {}), (2, 40, 128, 128, 3, {"stages": 2}),
]

if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ROOT)
    from yolov6_b200 import ops
    N, hw, cin, cout, k = (int(v) for v in sys.argv[1:6])
    force = eval(sys.argv[6])
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    xb = torch.randn(N, hw, hw, cin, generator=g).to(torch.bfloat16).to(dev)
    wb = (torch.randn(cout, k, k, cin, generator=g) / (k * k * cin) ** 0.5).to(torch.bfloat16).to(dev)
    bias = ops.pad_bias(torch.randn(cout, generator=g).to(dev), cout)
    y0 = torch.zeros(N, hw, hw, cout, dtype=torch.bfloat16, device=dev)
    y1 = torch.zeros_like(y0)
    ops.conv_fwd(xb, wb, bias, y0, act="relu", force=dict(force, pair=-1))
    torch.cuda.synchronize()
    for _ in range(3):
        ops.conv_fwd(xb, wb, bias, y1, act="relu", force=dict(force, pair=1))
        torch.cuda.synchronize()
    plan = ops.conv_plan(tuple(xb.shape), tuple(wb.shape), 1, 1, dict(force, pair=1))
    print("OK" if bool((y0 == y1).all()) else "MISMATCH", plan)
    sys.exit(0)

for c in CASES:
    args = [str(v) for v in c[:5]] + [repr(c[5])]
    r = subprocess.run([sys.executable, os.path.abspath(__file__)] + args, capture_output=True, text=True, timeout=120)
    tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or ["?"])[-1][:160]
    print(c, "->", tail, flush=True)
