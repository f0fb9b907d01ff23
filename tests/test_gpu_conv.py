"""GPU parity of the tcgen05 implicit-GEMM conv kernel (through the C ABI) against an fp64 CPU
convolution of the same (bf16-rounded or bf16x3-split) operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # name, N, H, W, Cin, Cout, k, stride, act, out_f32, nsplit, res, x_extra, y_extra, force
    ("1x1_c64", 2, 20, 20, 64, 64, 1, 1, "relu", False, 1, False, 0, 0, None),
    ("3x3_c128_40", 2, 40, 40, 128, 128, 3, 1, "relu", False, 1, False, 0, 0, None),
    ("3x3_c512_ntiles", 2, 20, 20, 512, 512, 3, 1, "relu", False, 1, False, 0, 0, None),
    ("3x3_s2_odd", 1, 23, 17, 64, 64, 3, 2, "relu", False, 1, False, 0, 0, None),
    ("3x3_s1_odd_cout96", 3, 23, 17, 64, 96, 3, 1, "silu", False, 1, False, 0, 0, None),
    ("3x3_cin32_sw64_s2", 2, 32, 32, 32, 64, 3, 2, "relu", False, 1, False, 0, 0, None),
    ("1x1_cin48_sw32", 2, 16, 16, 48, 96, 1, 1, "relu", False, 1, False, 0, 0, None),
    ("1x1_cout80_sigmoid_f32", 2, 20, 20, 128, 80, 1, 1, "sigmoid", True, 1, False, 0, 0, None),
    ("1x1_cout4_f32", 2, 20, 20, 64, 4, 1, 1, None, True, 1, False, 0, 0, None),
    ("1x1_cout68_f32", 2, 20, 20, 64, 68, 1, 1, None, True, 1, False, 0, 0, None),
    ("3x3_residual", 2, 20, 20, 64, 64, 3, 1, "relu", False, 1, True, 0, 0, None),
    ("3x3_slices", 2, 20, 20, 64, 64, 3, 1, "relu", False, 1, False, 64, 128, None),
    ("3x3_persistent", 4, 40, 40, 64, 64, 3, 1, "relu", False, 1, False, 0, 0, dict(grid=8)),
    ("3x3_direct_store", 2, 40, 40, 128, 128, 3, 1, "relu", False, 1, False, 0, 0, dict(direct=1)),
    ("3x3_x3", 2, 20, 20, 64, 64, 3, 1, "relu", False, 3, False, 0, 0, None),
    ("3x3_x3_s2", 2, 20, 20, 64, 128, 3, 2, "relu", False, 3, False, 0, 0, None),
    ("1x1_x3_f32out", 2, 20, 20, 64, 80, 1, 1, "sigmoid", True, 3, False, 0, 0, None),
    ("3x3_x3_residual", 2, 20, 20, 64, 64, 3, 1, "relu", False, 3, True, 0, 0, None),
    ("3x3_bi_batch5", 5, 10, 10, 64, 64, 3, 1, "relu", False, 1, False, 0, 0, None),
]


def ref_conv(x, w, b, stride, act, res, alpha):
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(), stride=stride,
                 padding=w.shape[1] // 2)
    y = {"relu": torch.relu, "silu": lambda t: t * torch.sigmoid(t), "sigmoid": torch.sigmoid, None: lambda t: t}[act](y)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + alpha * res.double()
    return y


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_fwd(case):
    from yolov6_b200 import ops
    name, N, H, W, Cin, Cout, k, stride, act, out_f32, nsplit, use_res, x_extra, y_extra, force = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    Ct = Cin + x_extra
    xfull = torch.randn(N, H, W, Ct, generator=g)
    w = torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    res = torch.randn(N, Ho, Wo, Cout, generator=g) if use_res else None
    xoff, yoff = x_extra // 2, y_extra // 2
    bias = ops.pad_bias(b.to(dev), Cout)
    ydt = torch.float32 if out_f32 else torch.bfloat16
    if nsplit == 1:
        xb, wb = xfull.to(torch.bfloat16), w.to(torch.bfloat16)
        resb = res.to(torch.bfloat16) if use_res else None
        y = torch.full((N, Ho, Wo, Cout + y_extra), 7.0, dtype=ydt, device=dev)
        ops.conv_fwd(xb.to(dev), wb.to(dev), bias, y, x_c_offset=xoff, stride=stride, act=act, y_c_offset=yoff,
                     res=resb.to(dev) if use_res else None, alpha=0.5, force=force)
        ref = ref_conv(xb.float()[..., xoff:xoff + Cin], wb.float(), b, stride, act, resb.float() if use_res else None, 0.5)
        got = y.float().cpu()
        assert bool((got[..., :yoff] == 7).all() and (got[..., yoff + Cout:] == 7).all()), "wrote outside its slice"
        got = got[..., yoff:yoff + Cout].double()
        tol = 2e-6 if out_f32 else 2.0 ** -8      # fp32 accumulate; bf16 output rounding = 2^-9 relative
    else:
        x3, w3 = ops.split3(xfull), ops.split3(w)
        res3 = ops.split3(res) if use_res else None
        y = torch.zeros((N, Ho, Wo, Cout) if out_f32 else (3, N, Ho, Wo, Cout), dtype=ydt, device=dev)
        ops.conv_fwd(x3.to(dev), w3.to(dev), bias, y, stride=stride, act=act, res=res3.to(dev) if use_res else None,
                     alpha=0.5, nsplit=3, force=force)
        ref = ref_conv(xfull, w, b, stride, act, res, 0.5)
        got = (y if out_f32 else y.float().sum(0)).cpu().double()
        tol = 5e-6                                  # fp32-equivalent mode
    err = ((got - ref).abs() / (1.0 + ref.abs())).max().item()
    assert err <= tol, f"{name}: rel err {err:.3e} > {tol:.1e}"


def test_conv_rejects_bad_arguments():
    from yolov6_b200 import ops
    dev = torch.device("cuda:0")
    x = torch.zeros(1, 8, 8, 24, dtype=torch.bfloat16, device=dev)      # Cin not a multiple of 16
    w = torch.zeros(16, 1, 1, 24, dtype=torch.bfloat16, device=dev)
    y = torch.zeros(1, 8, 8, 16, dtype=torch.bfloat16, device=dev)
    with pytest.raises(RuntimeError):
        ops.conv_fwd(x, w, None, y)
