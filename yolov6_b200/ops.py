"""Thin Python wrappers over the C ABI: one function per kernel entry point.

Tensors are torch CUDA tensors used purely as device memory; all layout decisions (NHWC, channel
slices, bf16x3 planes) are made by the caller (`yolov6_b200.engine`).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_CODES, DT_BF16, DT_F32, ConvDesc


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def pad_bias(bias, cout):
    """fp32 bias padded with zeros to a multiple of 256 (the kernel reads 16 floats at a time)."""
    n = (cout + 255) // 256 * 256
    out = torch.zeros(n, dtype=torch.float32, device=bias.device)
    out[:cout] = bias.float()
    return out


def split3(t):
    """fp32 -> stacked bf16 planes [3, ...] with hi + mid + lo == t to ~2^-24 relative."""
    t = t.float()
    p0 = t.to(torch.bfloat16)
    r1 = t - p0.float()
    p1 = r1.to(torch.bfloat16)
    r2 = r1 - p1.float()
    p2 = r2.to(torch.bfloat16)
    return torch.stack([p0, p1, p2])


def pair_view_weights(w33):
    """[Cout, 3, 3, Cin] weights of a 3x3 stride-2 conv -> [Cout, 3, 2, 2*Cin] weights of the same conv on the column-pair view
    [N, H, W/2, 2*Cin] of its input (3x2 kernel, stride (2, 1), pad (1, 1)): tap 0 = input columns (2j-2 | 2j-1), of which only the
    odd one is under the 3x3 window; tap 1 = (2j | 2j+1).  See include/yv6.h `stride_w` / `pair_view`."""
    cout, _, _, cin = w33.shape
    wf = torch.zeros(cout, 3, 2, 2 * cin, dtype=w33.dtype, device=w33.device)
    wf[:, :, 0, cin:] = w33[:, :, 0]
    wf[:, :, 1, :cin] = w33[:, :, 1]
    wf[:, :, 1, cin:] = w33[:, :, 2]
    return wf


def conv_fwd(x, w, bias, y, *, cin=None, x_c_offset=0, stride=1, act=None, y_c_offset=0,
             y_img_stride=None, y_h_stride=None, y_w_stride=None, y_elem_offset=0,
             res=None, res_c_offset=0, alpha=1.0, nsplit=1, force=None, stream=None, pad=None, out_hw=None, stride_w=0, pair_view=0):
    """y[..., y_c_offset:+Cout] = act(conv(x[..., x_c_offset:+Cin], w) + bias) (+ alpha*res).

    x: [N,H,W,Ct] bf16 (nsplit=1) or [3,N,H,W,Ct] (nsplit=3); w: [Cout,kh,kw,Cin] bf16 (or [3,...]);
    bias: padded fp32 (see pad_bias) or None; y: NHWC buffer, bf16 ([3,...] when nsplit=3) or fp32.
    """
    planes = nsplit == 3
    xs = x.shape[1:] if planes else x.shape
    ws = w.shape[1:] if planes else w.shape
    N, H, W, Ct = xs
    Cout, kh, kw, Cin = ws
    if cin is None:
        cin = Cin
    assert cin == Cin, (cin, Cin)
    d = ConvDesc()
    esz = x.element_size()
    d.x = x.data_ptr() + x_c_offset * esz
    d.N, d.H, d.W, d.Cin, d.x_c_total = N, H, W, Cin, Ct
    d.x_plane_stride = x.stride(0) if planes else 0
    d.w = w.data_ptr()
    d.w_plane_stride = w.stride(0) if planes else 0
    d.bias = bias.data_ptr() if bias is not None else 0
    d.Cout, d.kh, d.kw, d.stride, d.pad = Cout, kh, kw, stride, (kh // 2 if pad is None else pad[0])
    d.pad_w = _lib.PAD_SAME if pad is None else pad[1]
    if out_hw is not None:
        d.out_h, d.out_w = out_hw
    d.stride_w, d.pair_view = int(stride_w), int(pair_view)
    d.act = ACT_CODES[act]
    y_planes = planes and y.dtype == torch.bfloat16
    ysh = y.shape[1:] if y_planes else y.shape
    yst = y.stride()[1:] if y_planes else y.stride()
    d.y = y.data_ptr() + (y_c_offset + y_elem_offset) * y.element_size()
    d.y_dtype = DT_BF16 if y.dtype == torch.bfloat16 else DT_F32
    d.y_img_stride = yst[0] if y_img_stride is None else y_img_stride
    d.y_h_stride = yst[1] if y_h_stride is None else y_h_stride
    d.y_w_stride = yst[2] if y_w_stride is None else y_w_stride
    d.y_plane_stride = y.stride(0) if y_planes else 0
    if res is not None:
        rst = res.stride()[1:] if planes else res.stride()
        d.res = res.data_ptr() + res_c_offset * res.element_size()
        d.res_img_stride, d.res_h_stride, d.res_w_stride = rst[0], rst[1], rst[2]
        d.res_plane_stride = res.stride(0) if planes else 0
    d.alpha = float(alpha)
    d.nsplit = nsplit
    if force:
        for k, v in force.items():
            if k == "trace":                      # debug: uint64[16] device tensor for the kernel's phase stamps
                d.trace = v.data_ptr()
            else:
                setattr(d, "force_" + k, int(v))
    dev = x.device.index or 0
    _lib.check(_lib.lib().yv6_conv_fwd(_lib.handle(dev), C.byref(d), _lib.stream_ptr(stream)))
    return y


def conv_plan(x_shape, w_shape, stride=1, nsplit=1, force=None, device=0, stride_w=0, pad=None, out_hw=None, pair_view=0):
    """Tile plan (BW,BH,BI,BN,KB,stages,grid,tiles) the kernel would use for a shape."""
    N, H, W, Ct = x_shape
    Cout, kh, kw, Cin = w_shape
    d = ConvDesc()
    d.x = d.w = d.y = 16  # plan only: non-null, aligned
    d.N, d.H, d.W, d.Cin, d.x_c_total = N, H, W, Cin, Ct
    d.Cout, d.kh, d.kw, d.stride, d.pad = Cout, kh, kw, stride, kh // 2
    d.pad_w = _lib.PAD_SAME
    if pad is not None:
        d.pad, d.pad_w = pad
    if out_hw is not None:
        d.out_h, d.out_w = out_hw
    d.stride_w, d.pair_view = int(stride_w), int(pair_view)
    d.nsplit = nsplit
    if force:
        for k, v in force.items():
            setattr(d, "force_" + k, int(v))
    out = (C.c_int32 * 10)()
    _lib.check(_lib.lib().yv6_conv_plan(_lib.handle(device), C.byref(d), out))
    return dict(zip(("BW", "BH", "BI", "BN", "KB", "stages", "grid", "tiles", "halo", "a_res"), list(out)))
