"""Host-side check of the column-pair view used for 3x3 stride-2 convs (include/yv6.h `stride_w` / `pair_view`, yolov6_b200/ops.py
`pair_view_weights`, engine.py): the 3x2 / stride (2, 1) / pad (1, 1) conv over [N, H, W/2, 2*Cin] with the rearranged weights is the
original conv (reference ConvModule / RepVGGBlock with stride 2, yolov6/layers/common.py:26-60, 197-319), and the weights promise
what the kernel's zero-block skipping relies on.  Pure torch on the CPU; the CUDA path is tests/test_gpu_conv.py."""
import pytest
import torch
import torch.nn.functional as F

from yolov6_b200 import ops


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 8, 12, 4, 6), (1, 16, 16, 8, 8), (3, 6, 4, 16, 5)])
def test_pair_view_is_the_same_convolution(N, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(N * 100 + Cin)
    x = torch.randn(N, H, W, Cin, generator=g, dtype=torch.float64)          # NHWC, as the kernels see it
    w = torch.randn(Cout, 3, 3, Cin, generator=g, dtype=torch.float64)       # KRSC
    ref = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), stride=2, padding=1).permute(0, 2, 3, 1)
    wv = ops.pair_view_weights(w)                                            # [Cout, 3, 2, 2*Cin]
    xv = x.reshape(N, H, W // 2, 2 * Cin)                                    # same memory, pixel pairs as channels
    xv_p = F.pad(xv.permute(0, 3, 1, 2), (1, 0, 1, 1))                       # pad_w = 1 on the left only (out_w = W/2), rows 1 / 1
    got = F.conv2d(xv_p, wv.permute(0, 3, 1, 2), stride=(2, 1)).permute(0, 2, 3, 1)
    assert got.shape == ref.shape == (N, H // 2, W // 2, Cout)
    assert torch.allclose(got, ref, rtol=0, atol=1e-12)


def test_pair_view_weights_zero_block_promise():
    """`pair_view = 1` tells the kernel that w[:, :, 0, 0:Cin] (left tap, even pixel of the pair) is zero: those 64-channel blocks
    are never loaded or multiplied."""
    w = torch.randn(8, 3, 3, 64)
    wv = ops.pair_view_weights(w)
    assert wv.shape == (8, 3, 2, 128)
    assert bool((wv[:, :, 0, :64] == 0).all())
    assert torch.equal(wv[:, :, 0, 64:], w[:, :, 0]) and torch.equal(wv[:, :, 1, :64], w[:, :, 1]) and torch.equal(wv[:, :, 1, 64:], w[:, :, 2])
