"""GPU bring-up check for yv6_conv_fwd: many shapes vs torch fp32 conv, then timings.
Run on the B200 box:  python tools/gpu_conv_check.py  (writes gpurun_out/conv_check.json)"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov6_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
results = []


def ref_conv(x_nhwc_f32, w_krsc_f32, bias, stride, act, res=None, alpha=1.0):
    x = x_nhwc_f32.permute(0, 3, 1, 2).double()
    w = w_krsc_f32.permute(0, 3, 1, 2).double()
    y = F.conv2d(x, w, bias.double() if bias is not None else None, stride=stride, padding=w.shape[-1] // 2)
    if act == "relu":
        y = y.relu()
    elif act == "silu":
        y = y * torch.sigmoid(y)
    elif act == "sigmoid":
        y = torch.sigmoid(y)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + alpha * res.double()
    return y.float()


def run_case(name, N, H, W, Cin, Cout, k, stride, act="relu", out_f32=False, nsplit=1, use_res=False,
             x_extra=0, y_extra=0, force=None, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    Ct = Cin + x_extra
    xfull = torch.randn(N, H, W, Ct, generator=g).to(dev)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev) * 0.1
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    xoff = x_extra // 2
    res = torch.randn(N, Ho, Wo, Cout, generator=g).to(dev) if use_res else None
    bias = ops.pad_bias(b, Cout)
    ydt = torch.float32 if out_f32 else torch.bfloat16
    Cyt = Cout + y_extra
    yoff = y_extra // 2
    try:
        if nsplit == 1:
            xb = xfull.to(torch.bfloat16)
            wb = w.to(torch.bfloat16)
            resb = res.to(torch.bfloat16) if use_res else None
            y = torch.full((N, Ho, Wo, Cyt), 7.0, dtype=ydt, device=dev)
            ops.conv_fwd(xb, wb, bias, y, x_c_offset=xoff, stride=stride, act=act, y_c_offset=yoff, res=resb,
                         alpha=0.5, force=force)
            torch.cuda.synchronize()
            ref = ref_conv(xb.float()[..., xoff:xoff + Cin], wb.float(), b, stride, act,
                           resb.float() if use_res else None, 0.5)
            got = y.float()[..., yoff:yoff + Cout]
            untouched = bool((y.float()[..., :yoff] == 7).all() and (y.float()[..., yoff + Cout:] == 7).all())
            tol = 1e-5 if out_f32 else 2.0 ** -8
        else:
            x3 = ops.split3(xfull)
            w3 = ops.split3(w)
            res3 = ops.split3(res) if use_res else None
            if out_f32:
                y = torch.full((N, Ho, Wo, Cyt), 7.0, dtype=ydt, device=dev)
            else:
                y = torch.full((3, N, Ho, Wo, Cyt), 7.0, dtype=ydt, device=dev)
            ops.conv_fwd(x3, w3, bias, y, x_c_offset=xoff, stride=stride, act=act, y_c_offset=yoff, res=res3,
                         alpha=0.5, nsplit=3, force=force)
            torch.cuda.synchronize()
            ref = ref_conv(xfull[..., xoff:xoff + Cin], w, b, stride, act, res if use_res else None, 0.5)
            got = (y if out_f32 else y.float().sum(0))[..., yoff:yoff + Cout]
            untouched = True
            tol = 1e-5
        err = (got - ref).abs()
        rel = (err / (ref.abs() + 1.0)).max().item()
        ok = rel <= tol and untouched
        plan = ops.conv_plan((N, H, W, Ct), (Cout, k, k, Cin), stride, nsplit, force)
        results.append(dict(name=name, ok=bool(ok), max_rel=rel, max_abs=err.max().item(), untouched=untouched, plan=plan))
        print(f"{'PASS' if ok else 'FAIL'} {name}: rel={rel:.3e} abs={err.max().item():.3e} untouched={untouched} plan={plan}", flush=True)
    except Exception as e:  # noqa: BLE001
        results.append(dict(name=name, ok=False, error=repr(e)))
        print(f"ERROR {name}: {e!r}", flush=True)


def bench_case(name, N, H, W, Cin, Cout, k, stride, nsplit=1, force=None, iters=20, act="relu", out_f32=False):
    xb = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
    wb = (torch.randn(Cout, k, k, Cin, device=dev) / (k * k * Cin) ** 0.5).to(torch.bfloat16)
    if nsplit == 3:
        xb = torch.stack([xb, xb, xb])
        wb = torch.stack([wb, wb, wb])
    bias = ops.pad_bias(torch.zeros(Cout, device=dev), Cout)
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    y = torch.empty((3, N, Ho, Wo, Cout) if nsplit == 3 else (N, Ho, Wo, Cout), dtype=torch.float32 if out_f32 else torch.bfloat16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    try:
        for _ in range(3):
            ops.conv_fwd(xb, wb, bias, y, stride=stride, act=act, nsplit=nsplit, force=force)
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_fwd(xb, wb, bias, y, stride=stride, act=act, nsplit=nsplit, force=force)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        flops = 2.0 * N * Ho * Wo * Cout * Cin * k * k * (6 if nsplit == 3 else 1)
        byts = (N * H * W * Cin + N * Ho * Wo * Cout) * 2 * (3 if nsplit == 3 else 1)
        # cudnn bf16 for comparison
        xc = xb[0] if nsplit == 3 else xb
        xn = xc.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wn = (wb[0] if nsplit == 3 else wb).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        for _ in range(3):
            F.conv2d(xn, wn, stride=stride, padding=k // 2)
        tc = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            F.conv2d(xn, wn, stride=stride, padding=k // 2)
            e1.record()
            torch.cuda.synchronize()
            tc.append(e0.elapsed_time(e1))
        tc.sort()
        r = dict(name=name, ms=ms, tflops=flops / ms / 1e9, gbs=byts / ms / 1e6, cudnn_ms=tc[len(tc) // 2],
                 plan=ops.conv_plan(tuple(xc.shape), tuple((wb[0] if nsplit == 3 else wb).shape), stride, nsplit, force))
        results.append(r)
        print(f"BENCH {name}: {ms:.4f} ms  {r['tflops']:.1f} TFLOP/s  {r['gbs']:.0f} GB/s  (cudnn bf16 {r['cudnn_ms']:.4f} ms) plan={r['plan']}", flush=True)
    except Exception as e:  # noqa: BLE001
        results.append(dict(name=name, error=repr(e)))
        print(f"ERROR {name}: {e!r}", flush=True)


BENCH_SHAPES = {
    "s1_64@160": (160, 160, 64, 64, 3, 1), "s1_128@80": (80, 80, 128, 128, 3, 1),
    "s1_256@40": (40, 40, 256, 256, 3, 1), "s1_512@20": (20, 20, 512, 512, 3, 1),
    "s1_128@40": (40, 40, 128, 128, 3, 1), "s1_256@20": (20, 20, 256, 256, 3, 1),
    "s1_64@80": (80, 80, 64, 64, 3, 1), "s2_32_64@160": (320, 320, 32, 64, 3, 2),
    "s2_64_128@80": (160, 160, 64, 128, 3, 2), "s2_128_256@40": (80, 80, 128, 256, 3, 2),
    "s2_256_512@20": (40, 40, 256, 512, 3, 2), "1x1_512_256@20": (20, 20, 512, 256, 1, 1),
    "1x1_384_128@40": (40, 40, 384, 128, 1, 1), "1x1_128_128@80": (80, 80, 128, 128, 1, 1),
    "1x1_256_256@20": (20, 20, 256, 256, 1, 1), "1x1_64_64@160": (160, 160, 64, 64, 1, 1),
    "1x1_128_128@20": (20, 20, 128, 128, 1, 1),
}

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--profile":  # few launches of one shape, for ncu
        H, W, Cin, Cout, k, st = BENCH_SHAPES[sys.argv[2]]
        xb = torch.randn(32, H, W, Cin, device=dev).to(torch.bfloat16)
        wb = (torch.randn(Cout, k, k, Cin, device=dev) / (k * k * Cin) ** 0.5).to(torch.bfloat16)
        bias = ops.pad_bias(torch.zeros(Cout, device=dev), Cout)
        Ho, Wo = (H + 2 * (k // 2) - k) // st + 1, (W + 2 * (k // 2) - k) // st + 1
        y = torch.empty(32, Ho, Wo, Cout, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            ops.conv_fwd(xb, wb, bias, y, stride=st, act=sys.argv[3] if len(sys.argv) > 3 else "relu")
        torch.cuda.synchronize()
        sys.exit(0)
    # --- correctness ---
    run_case("1x1_c64_small", 1, 8, 16, 64, 64, 1, 1, force=dict(bw=16, bh=8))
    run_case("1x1_c64_auto", 2, 20, 20, 64, 64, 1, 1)
    run_case("3x3_c64_s1", 2, 20, 20, 64, 64, 3, 1)
    run_case("3x3_c128_s1_40", 2, 40, 40, 128, 128, 3, 1)
    run_case("3x3_c256_s1_bn256", 2, 20, 20, 256, 256, 3, 1)
    run_case("3x3_c512_s1_ntiles2", 2, 20, 20, 512, 512, 3, 1)
    run_case("3x3_s2_c64_128", 2, 40, 40, 64, 128, 3, 2)
    run_case("3x3_s2_odd", 1, 23, 17, 64, 64, 3, 2)
    run_case("3x3_s1_odd", 3, 23, 17, 64, 96, 3, 1)
    run_case("3x3_cin32_sw64_s2", 2, 32, 32, 32, 64, 3, 2)
    run_case("3x3_cin16_sw32", 2, 16, 16, 16, 32, 3, 1)
    run_case("1x1_cin48_sw32", 2, 16, 16, 48, 96, 1, 1)
    run_case("1x1_cout80_sigmoid_f32", 2, 20, 20, 128, 80, 1, 1, act="sigmoid", out_f32=True)
    run_case("1x1_cout4_f32", 2, 20, 20, 64, 4, 1, 1, act=None, out_f32=True)
    run_case("1x1_cout68_f32", 2, 20, 20, 64, 68, 1, 1, act=None, out_f32=True)
    run_case("3x3_silu", 2, 20, 20, 64, 64, 3, 1, act="silu")
    run_case("3x3_residual", 2, 20, 20, 64, 64, 3, 1, use_res=True)
    run_case("3x3_slices", 2, 20, 20, 64, 64, 3, 1, x_extra=64, y_extra=128)
    run_case("3x3_persistent_grid8", 4, 40, 40, 64, 64, 3, 1, force=dict(grid=8))
    run_case("3x3_stages2", 2, 40, 40, 128, 128, 3, 1, force=dict(stages=2))
    run_case("3x3_x3", 2, 20, 20, 64, 64, 3, 1, nsplit=3)
    run_case("3x3_x3_s2_res", 2, 20, 20, 64, 128, 3, 2, nsplit=3)
    run_case("1x1_x3_f32out", 2, 20, 20, 64, 80, 1, 1, nsplit=3, act="sigmoid", out_f32=True)
    run_case("3x3_x3_residual", 2, 20, 20, 64, 64, 3, 1, nsplit=3, use_res=True)
    run_case("3x3_big_bs8_160", 8, 160, 160, 64, 64, 3, 1)
    run_case("3x3_s2_stem_like", 2, 64, 64, 16, 32, 3, 2)
    run_case("nohalo_3x3_c128_40", 2, 40, 40, 128, 128, 3, 1, force=dict(halo=-1))
    run_case("halo_3x3_c64_s1_odd", 3, 23, 17, 64, 96, 3, 1, force=dict(halo=1))
    run_case("halo_3x3_c256_bn256", 2, 32, 32, 256, 256, 3, 1, force=dict(halo=1))
    run_case("halo_3x3_c512_ntiles", 2, 16, 16, 512, 512, 3, 1, force=dict(halo=1))
    run_case("halo_3x3_x3", 2, 24, 24, 64, 64, 3, 1, nsplit=3, force=dict(halo=1))
    run_case("halo_3x3_res_slices", 2, 24, 24, 64, 64, 3, 1, use_res=True, x_extra=64, y_extra=64, force=dict(halo=1))
    run_case("halo_3x3_silu_c128_cout80_f32", 2, 24, 24, 128, 80, 3, 1, act="silu", out_f32=True, force=dict(halo=1))
    run_case("direct_3x3_c128", 2, 40, 40, 128, 128, 3, 1, force=dict(direct=1))
    run_case("blockstore_3x3_c128", 2, 40, 40, 128, 128, 3, 1, force=dict(direct=2))
    run_case("blockstore_x3_cout80_f32", 2, 24, 24, 64, 80, 1, 1, nsplit=3, act="sigmoid", out_f32=True, force=dict(direct=2))
    run_case("warpstore_bw32", 2, 64, 64, 64, 64, 3, 1, force=dict(bw=32, bh=4, halo=-1))
    run_case("warpstore_bw128_1x1", 1, 4, 256, 64, 64, 1, 1, force=dict(bw=128, bh=1))
    run_case("warpstore_bw4_s2", 2, 40, 40, 64, 64, 3, 2, force=dict(bw=4, bh=32))
    run_case("persistent_many_tiles_grid3", 4, 64, 64, 64, 64, 3, 1, force=dict(grid=3))
    run_case("persistent_grid2_nohalo_odd_tiles", 3, 48, 40, 64, 96, 3, 1, force=dict(grid=2, halo=-1))
    run_case("direct_1x1_cout80_f32", 2, 20, 20, 128, 80, 1, 1, act="sigmoid", out_f32=True, force=dict(direct=1))
    run_case("3x3_cout96_partial_chunk", 2, 20, 20, 64, 96, 3, 1)
    run_case("3x3_cout32", 2, 20, 20, 64, 32, 3, 1)
    run_case("1x1_x3_cout68_f32", 2, 20, 20, 64, 68, 1, 1, nsplit=3, act=None, out_f32=True)
    run_case("3x3_bi2_batch5", 5, 10, 10, 64, 64, 3, 1)
    run_case("3x3_c384_ntiles", 2, 20, 20, 128, 384, 3, 1)
    nfail = sum(1 for r in results if not r.get("ok", True))
    print(f"correctness: {len(results) - nfail} pass, {nfail} fail", flush=True)
    # --- timings (YOLOv6-S bs32 main shapes) ---
    B = 32
    bench_case("s1_64@160", B, 160, 160, 64, 64, 3, 1)
    bench_case("s1_128@80", B, 80, 80, 128, 128, 3, 1)
    bench_case("s1_256@40", B, 40, 40, 256, 256, 3, 1)
    bench_case("s1_512@20", B, 20, 20, 512, 512, 3, 1)
    bench_case("s1_128@40", B, 40, 40, 128, 128, 3, 1)
    bench_case("s1_256@20", B, 20, 20, 256, 256, 3, 1)
    bench_case("s1_64@80", B, 80, 80, 64, 64, 3, 1)
    bench_case("s2_32_64@160", B, 320, 320, 32, 64, 3, 2)
    bench_case("s2_64_128@80", B, 160, 160, 64, 128, 3, 2)
    bench_case("s2_128_256@40", B, 80, 80, 128, 256, 3, 2)
    bench_case("s2_256_512@20", B, 40, 40, 256, 512, 3, 2)
    bench_case("1x1_512_256@20", B, 20, 20, 512, 256, 1, 1)
    bench_case("1x1_384_128@40", B, 40, 40, 384, 128, 1, 1)
    bench_case("1x1_128_128@80", B, 80, 80, 128, 128, 1, 1)
    for nm in ("s1_64@160", "s1_128@80", "s1_256@40", "s1_512@20", "s1_128@40", "s1_256@20", "s1_64@80"):
        H_, W_, ci, co, k_, st = BENCH_SHAPES[nm]
        bench_case(nm + "_nohalo", B, H_, W_, ci, co, k_, st, force=dict(halo=-1))
    bench_case("s1_512@20_halo", B, 20, 20, 512, 512, 3, 1, force=dict(halo=1))
    bench_case("s1_256@20_halo", B, 20, 20, 256, 256, 3, 1, force=dict(halo=1))
    bench_case("1x1_64_64@80_silu", B, 80, 80, 64, 64, 1, 1, act="silu")
    bench_case("3x3_64_64@80_silu", B, 80, 80, 64, 64, 3, 1, act="silu")
    bench_case("1x1_64_80@80_sigmoid_f32", B, 80, 80, 64, 80, 1, 1, act="sigmoid", out_f32=True)
    bench_case("1x1_64_64@160", B, 160, 160, 64, 64, 1, 1)
    bench_case("s1_256@40_bn128", B, 40, 40, 256, 256, 3, 1, force=dict(bn=128))
    bench_case("s1_256@40_bn256", B, 40, 40, 256, 256, 3, 1, force=dict(bn=256))
    bench_case("s1_256@40_direct", B, 40, 40, 256, 256, 3, 1, force=dict(direct=1))
    bench_case("s1_128@80_bn64", B, 80, 80, 128, 128, 3, 1, force=dict(bn=64))
    bench_case("s1_256@40_x3", B, 40, 40, 256, 256, 3, 1, nsplit=3)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/conv_check.json", "w") as f:
        json.dump(results, f, indent=1)
    sys.exit(1 if nfail else 0)
