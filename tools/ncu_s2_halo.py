"""One layer for an `ncu --set full` capture of the stride-2 halo mainloop (conv_igemm_kernel<4, 4, false>): ERBlock_2.0 of
YOLOv6-S at bs32 (32 -> 64 channels, 320x320 -> 160x160) on the column-pair view.  usage (GPU box):
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_igemm -s 2 -c 1 -o out python tools/ncu_s2_halo.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, hw, cin, cout = 32, 320, 32, 64
xb = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
w = torch.randn(cout, 3, 3, cin, device=dev) / (9 * cin) ** 0.5
wv = ops.pair_view_weights(w).to(torch.bfloat16)
bias = ops.pad_bias(torch.zeros(cout, device=dev), cout)
y = torch.empty(B, hw // 2, hw // 2, cout, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    ops.conv_fwd(xb.view(B, hw, hw // 2, 2 * cin), wv, bias, y, stride=2, stride_w=1, pad=(1, 1), out_hw=(0, hw // 2), pair_view=1, act="relu")
torch.cuda.synchronize()
print("plan", ops.conv_plan((B, hw, hw // 2, 2 * cin), tuple(wv.shape), 2, 1, None, stride_w=1, pad=(1, 1), out_hw=(0, hw // 2), pair_view=1))
