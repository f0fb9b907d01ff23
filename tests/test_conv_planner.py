"""CPU checks of the conv tile planner (`plan_conv` of yolov6_b200/csrc/yv6_conv_igemm.cu) through the host-only C-ABI entry
`yv6_conv_plan_host`, with the B200's device properties stated explicitly (148 SMs, 232448 bytes of opt-in shared memory, 74
co-resident CTA pairs): every conv of every supported model (reference configs/yolov6{n,s,m}.py, yolov6l6.py) at the BASELINE.json
configurations gets a plan that fits the SM, and the YOLOv6-S plans are the ones the committed B200 profile shows
(profiles/r02_conv_layers_yolov6s.md).  No compute, no GPU."""
import ctypes as C

import pytest

from yolov6_b200 import _lib
from yolov6_b200.arch import build_graph
from yolov6_b200.configs import get_config

B200 = (148, 232448, 74)


def plan(d):
    out = (C.c_int32 * 12)()
    rc = _lib.lib().yv6_conv_plan_host(*B200, C.byref(d), out)
    assert rc == 0, _lib.lib().yv6_last_error().decode()
    keys = ("BW", "BH", "BI", "BN", "KB", "stages", "grid", "tiles", "halo", "mode", "smem", "tmem")
    return dict(zip(keys, list(out)))


def layer_descs(name, batch, size):
    """(op name, descriptor) of every conv launch of a model's forward, shaped as engine.InferEngine._plan shapes them."""
    g = build_graph(get_config(name), 80, name)
    for op in g.ops:
        if op.kind not in ("conv", "pred", "convT"):
            continue
        sb = g.bufs[op.src.buf]
        h = w = size >> sb.level
        for q in range(4 if op.kind == "convT" else 1):
            d = _lib.ConvDesc()
            d.x = 4096 + op.src.c_off * 2
            d.w = d.y = 4096
            d.N, d.H, d.W, d.Cin, d.x_c_total = batch, h, w, op.cin, sb.c_total
            d.Cout = op.cout
            d.kh = d.kw = 1 if op.kind == "convT" else op.k
            d.stride = 1 if op.kind == "convT" else op.s
            d.pad, d.pad_w, d.nsplit = d.kh // 2, _lib.PAD_SAME, 1
            if op.kind == "pred":
                d.y_dtype = _lib.DT_F32
                d.y_img_stride, d.y_h_stride, d.y_w_stride = 8400 * op.cout, w * op.cout, op.cout
            else:
                db = g.bufs[op.dst.buf]
                oh = h * 2 if op.kind == "convT" else h // d.stride
                d.y_dtype = _lib.DT_BF16
                d.y = 4096 + op.dst.c_off * 2
                if op.kind == "convT":
                    d.y_img_stride, d.y_h_stride, d.y_w_stride = oh * oh * db.c_total, 2 * oh * db.c_total, 2 * db.c_total
                else:
                    d.y_img_stride, d.y_h_stride, d.y_w_stride = oh * oh * db.c_total, oh * db.c_total, db.c_total
            if op.kind == "conv" and op.k == 3 and op.s == 2 and (op.cin <= 32 or op.cin in (64, 128)) and op.src.c_off == 0 \
                    and sb.c_total == op.cin and w % 2 == 0:
                # engine.py: column-pair view, kept for > 32 channels only where the halo mainloop takes it
                keep = (d.W, d.Cin, d.x_c_total, d.kw, d.stride_w, d.pad_w, d.out_w)
                d.W, d.Cin, d.x_c_total, d.kw, d.stride_w, d.pad_w, d.out_w, d.pair_view = w // 2, 2 * op.cin, 2 * op.cin, 2, 1, 1, w // 2, 1
                if op.cin > 32 and plan(d)["halo"] != 2:
                    d.W, d.Cin, d.x_c_total, d.kw, d.stride_w, d.pad_w, d.out_w = keep
                    d.pair_view = 0
            yield op.name, d


@pytest.mark.parametrize("name,batch,size", [("yolov6n", 32, 640), ("yolov6s", 32, 640), ("yolov6s", 1, 64), ("yolov6m", 8, 640),
                                             ("yolov6m", 64, 640), ("yolov6l6", 2, 1280), ("yolov6l6", 16, 1280), ("yolov6l6", 1, 128),
                                             ("yolov6s", 4, 416), ("yolov6n", 2, 96)])
def test_every_layer_of_every_model_gets_a_plan_that_fits_the_sm(name, batch, size):
    n = 0
    for lname, d in layer_descs(name, batch, size):
        p = plan(d)              # asserts rc == 0: no YV6_REQUIRE of the planner fires, shared memory and TMEM fit
        assert 0 < p["smem"] <= B200[1] and p["tmem"] in (32, 64, 128, 256, 512), (lname, p)
        assert p["BW"] * p["BH"] * p["BI"] <= 128 and p["BN"] % 16 == 0 and 16 <= p["BN"] <= 256, (lname, p)
        assert 1 <= p["grid"] <= B200[0] and p["stages"] >= 2, (lname, p)
        if p["mode"] // 10 % 10:                                # CTA pairs: an even grid of at most 74 clusters
            assert p["grid"] % 2 == 0 and p["grid"] <= 2 * B200[2], (lname, p)
        n += 1
    assert n >= 60


def test_yolov6s_bench_plans_match_the_b200_profile():
    """The plans of the benchmark configuration as measured on the B200 (profiles/r02_conv_layers_yolov6s.md, column
    `tile BWxBHxBI BN KB stg halo`): stride-2 halo mainloop on six of the eight 3x3 stride-2 layers, CTA pairs from 128 output
    channels there and on the 3x3 stride-1 layers over >= 128 channels, resident weights for the 64-channel layers."""
    got = {ln: plan(d) for ln, d in layer_descs("yolov6s", 32, 640)}
    want = {   # name: (BW, BH, BN, stages, halo, mode)
        "backbone.ERBlock_2.0": (8, 16, 64, 6, 2, 201), "backbone.ERBlock_3.0": (8, 16, 128, 9, 2, 211),
        "backbone.ERBlock_4.0": (8, 16, 256, 5, 2, 210), "backbone.ERBlock_5.0": (20, 5, 256, 3, 0, 0),
        "neck.Bifusion0.downsample": (8, 16, 128, 10, 2, 210), "neck.Bifusion1.downsample": (8, 16, 64, 9, 2, 201),
        "neck.downsample2": (8, 16, 64, 9, 2, 201), "neck.downsample1": (20, 5, 128, 4, 0, 0),
        "backbone.ERBlock_2.1.conv1": (8, 16, 64, 9, 1, 301), "backbone.ERBlock_3.1.conv1": (8, 16, 128, 11, 1, 310),
        "backbone.ERBlock_4.1.conv1": (8, 16, 256, 5, 1, 310), "neck.Bifusion1.cv2": (32, 4, 64, 6, 0, 0),
    }
    for ln, w in want.items():
        p = got[ln]
        assert (p["BW"], p["BH"], p["BN"], p["stages"], p["halo"], p["mode"]) == w, (ln, p)


def test_planner_rejects_what_the_kernel_cannot_run():
    d = _lib.ConvDesc()
    d.x = d.w = d.y = 4096
    d.N, d.H, d.W, d.Cin, d.x_c_total, d.Cout, d.kh, d.kw, d.stride, d.pad, d.nsplit = 1, 8, 8, 24, 24, 16, 1, 1, 1, 0, 1
    d.pad_w = _lib.PAD_SAME
    out = (C.c_int32 * 12)()
    assert _lib.lib().yv6_conv_plan_host(*B200, C.byref(d), out) != 0          # Cin must be a multiple of 16
    assert b"Cin" in _lib.lib().yv6_last_error()
    d.Cin = d.x_c_total = 32
    d.kh = 5
    assert _lib.lib().yv6_conv_plan_host(*B200, C.byref(d), out) != 0          # kernel sizes 1..3 only


def test_host_planner_reproduces_every_plan_of_the_b200_profile():
    """`yv6_conv_plan_host` with the B200's properties against the plans `yv6_conv_plan` reported ON the B200 for every conv launch
    of the benchmark step (profiles/r02_conv_layers_yolov6s.md, written by tools/profile_layers.py on the GPU box).  The fused
    cls | reg head convs (engine.py sibling fusion: one launch with 2 x Cout) are the only launches this mirror does not shape."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_conv_layers_yolov6s.md")
    prof = {}
    for line in open(path):
        c = [x.strip() for x in line.split("|")]
        if len(c) > 10 and c[1] not in ("layer", "---"):
            prof[c[1]] = c[10]
    count, compared = {}, 0
    for ln, d in layer_descs("yolov6s", 32, 640):
        q = count.get(ln, 0)
        count[ln] = q + 1
        key = ln if ln in prof else f"{ln}[{q}]"            # transposed convs: one launch per quadrant
        if key not in prof or ".cls_convs." in ln or ".reg_convs." in ln:
            continue
        p = plan(d)
        assert prof[key] == f"{p['BW']}x{p['BH']}x{p['BI']} {p['BN']} {p['KB']} {p['stages']} {p['halo']}/{p['mode']}", (key, prof[key], p)
        compared += 1
    assert compared >= 65
