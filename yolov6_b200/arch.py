"""Declarative layer graph of the YOLOv6 networks on the hot path.

The reference builds its models as nested nn.Modules (yolov6/models/yolo.py:55-133 -> efficientrep.py,
reppan.py, effidehead.py, layers/common.py).  Here a network is a flat list of ops over NHWC
activation buffers, emitted once per (config, num_classes):

  * every op names the reference parameter prefix it owns (`backbone.ERBlock_3.1.block.0`, ...), so
    the parameter container (model.py) exposes exactly the reference's `state_dict` keys and released
    checkpoints load unchanged;
  * `torch.cat` never happens: a concat is one buffer and its producers write channel slices
    (reppan.py:228,232, common.py:650,718 -> `dst=T(buf, c_off, c)`);
  * the engine (engine.py) walks the list and issues one kernel per op through the C ABI.

Op kinds: stem | conv (rep / cba / plain parameter layouts) | convT | pool.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass(frozen=True)
class T:
    """Channel slice [c_off, c_off + c) of activation buffer `buf`."""
    buf: int
    c_off: int
    c: int


@dataclass
class Buf:
    level: int      # spatial size = input / 2**level
    c_total: int
    name: str = ""


@dataclass
class Op:
    kind: str                 # 'stem' | 'conv' | 'convT' | 'pool' | 'pred'
    name: str                 # reference parameter prefix
    layout: str = ""          # 'rep' | 'cba' | 'plain' | 'convT'
    src: Optional[T] = None
    dst: Optional[T] = None
    cin: int = 0
    cout: int = 0
    k: int = 1
    s: int = 1
    act: Optional[str] = None
    res: Optional[T] = None
    alpha: Optional[str] = None   # parameter name of BottleRep.alpha (common.py:600-603)
    head: Optional[tuple] = None  # ('cls' | 'reg', level index) for the prediction convs


@dataclass
class Graph:
    name: str
    num_classes: int
    strides: List[int]
    use_dfl: bool
    reg_max: int
    mode: str
    bufs: List[Buf] = field(default_factory=list)
    ops: List[Op] = field(default_factory=list)
    feat: List[T] = field(default_factory=list)   # neck outputs (reference `featmaps`, yolo.py:37-39)
    fuse_ab: bool = False
    anchors_init: Optional[list] = None           # fuse_ab: per level [w0, h0, w1, h1, w2, h2] in pixels
    distill_ns: bool = False
    dist_reg_ch: int = 0                          # distill_ns: channels of the DFL (reg_preds_dist) branch

    # -- builders ---------------------------------------------------------------------------
    def buf(self, level, c_total, name=""):
        self.bufs.append(Buf(level, c_total, name))
        return len(self.bufs) - 1

    def new(self, level, c, name=""):
        return T(self.buf(level, c, name), 0, c)

    def level(self, t):
        return self.bufs[t.buf].level

    def conv(self, name, layout, src, cout, k=1, s=1, act="relu", dst=None, res=None, alpha=None):
        lvl = self.level(src) + (1 if s == 2 else 0)
        if dst is None:
            dst = self.new(lvl, cout, name)
        assert dst.c == cout and self.level(dst) == lvl, (name, dst, cout, lvl)
        self.ops.append(Op("conv", name, layout, src, dst, src.c, cout, k, s, act, res, alpha))
        return dst

    def block(self, name, src, cout, s=1, dst=None, res=None, alpha=None):
        """get_block(training_mode) of common.py:721-737: RepVGGBlock (relu) or ConvBNSiLU / ConvBNReLU."""
        if self.mode == "repvgg":
            return self.conv(name, "rep", src, cout, 3, s, "relu", dst, res, alpha)
        return self.conv(name, "cba", src, cout, 3, s, "silu" if self.mode == "conv_silu" else "relu", dst, res, alpha)

    @property
    def act(self):
        return "silu" if self.mode == "conv_silu" else "relu"


DETECT_DEFAULT_REG_MAX = 16  # effidehead.py:16
AB_ANCHORS = 3                # build_network passes num_anchors = 3 to the fuse_ab head (yolo.py:125)


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


def _rep_block(g, name, src, cout, n, dst=None):
    """RepBlock with plain basic blocks (common.py:569-588)."""
    for i in range(n):
        last = i == n - 1
        src = g.block(f"{name}.conv1" if i == 0 else f"{name}.block.{i - 1}", src, cout, dst=dst if last else None)
    return src


def _bottle_stage(g, name, src, c, n, dst=None):
    """RepBlock(block=BottleRep, weight=True): n // 2 BottleReps, each conv2(conv1(x)) + alpha * x
    (common.py:579-582, 591-608).  The shortcut is fused into conv2's epilogue."""
    nb = max(n // 2, 1)
    for i in range(nb):
        p = f"{name}.conv1" if i == 0 else f"{name}.block.{i - 1}"
        mid = g.block(p + ".conv1", src, c)
        shortcut = src.c == c
        src = g.block(p + ".conv2", mid, c, dst=dst if i == nb - 1 else None,
                      res=src if shortcut else None, alpha=p + ".alpha")
    return src


def _bepc3(g, name, src, cout, n, e, dst=None):
    """BepC3 (common.py:634-650): cv3(cat(m(cv1 x), cv2 x)); the cat is one buffer with two slices."""
    c_ = int(cout * e)
    lvl = g.level(src)
    cat = g.buf(lvl, 2 * c_, name + ".cat")
    a = g.conv(name + ".cv1", "cba", src, c_, 1, 1, g.act)
    _bottle_stage(g, name + ".m", a, c_, n, dst=T(cat, 0, c_))
    g.conv(name + ".cv2", "cba", src, c_, 1, 1, g.act, dst=T(cat, c_, c_))
    return g.conv(name + ".cv3", "cba", T(cat, 0, 2 * c_), cout, 1, 1, g.act, dst=dst)


def _sppf(g, name, src, cout):
    """SimSPPF / SPPF (common.py:97-133)."""
    p = name + ".sppf"
    c_ = src.c // 2
    lvl = g.level(src)
    cat = g.buf(lvl, 4 * c_, p + ".cat")
    g.conv(p + ".cv1", "cba", src, c_, 1, 1, g.act, dst=T(cat, 0, c_))
    g.ops.append(Op("pool", p + ".m", src=T(cat, 0, c_), dst=T(cat, 0, 4 * c_), cin=c_, cout=c_))
    return g.conv(p + ".cv2", "cba", T(cat, 0, 4 * c_), cout, 1, 1, g.act)


def _cspsppf(g, name, src, cout):
    """SimCSPSPPF / CSPSPPF (common.py:135-178), e = 0.5."""
    p = name + ".cspsppf"
    c_ = int(cout * 0.5)
    lvl = g.level(src)
    cat4 = g.buf(lvl, 4 * c_, p + ".cat4")
    cat2 = g.buf(lvl, 2 * c_, p + ".cat2")
    a = g.conv(p + ".cv1", "cba", src, c_, 1, 1, g.act)
    a = g.conv(p + ".cv3", "cba", a, c_, 3, 1, g.act)
    g.conv(p + ".cv4", "cba", a, c_, 1, 1, g.act, dst=T(cat4, 0, c_))
    g.conv(p + ".cv2", "cba", src, c_, 1, 1, g.act, dst=T(cat2, 0, c_))
    g.ops.append(Op("pool", p + ".m", src=T(cat4, 0, c_), dst=T(cat4, 0, 4 * c_), cin=c_, cout=c_))
    b = g.conv(p + ".cv5", "cba", T(cat4, 0, 4 * c_), c_, 1, 1, g.act)
    g.conv(p + ".cv6", "cba", b, c_, 3, 1, g.act, dst=T(cat2, c_, c_))
    return g.conv(p + ".cv7", "cba", T(cat2, 0, 2 * c_), cout, 1, 1, g.act)


def _bifusion(g, name, x0, x1, x2, cout):
    """BiFusion (common.py:695-718), always ReLU: cv3(cat(up2x(x0), cv1(x1), down(cv2(x2))))."""
    lvl = g.level(x1)
    cat = g.buf(lvl, 3 * cout, name + ".cat")
    g.ops.append(Op("convT", name + ".upsample", "convT", x0, T(cat, 0, cout), x0.c, cout, 2, 2, None))
    g.conv(name + ".cv1", "cba", x1, cout, 1, 1, "relu", dst=T(cat, cout, cout))
    t = g.conv(name + ".cv2", "cba", x2, cout, 1, 1, "relu")
    g.conv(name + ".downsample", "cba", t, cout, 3, 2, "relu", dst=T(cat, 2 * cout, cout))
    return g.conv(name + ".cv3", "cba", T(cat, 0, 3 * cout), cout, 1, 1, "relu")


def build_graph(cfg, num_classes=80, name="yolov6", fuse_ab=False, distill_ns=False):
    """cfg: dict with the fields of the reference's `config.model` (see configs.py / config_from_reference).
    fuse_ab: add the anchor-aided training branch of effidehead_fuseab.py (two more 1x1 pred convs per level).
    distill_ns: the N / S student head of effidehead_distill_ns.py -- `reg_preds` emits the 4 (l, r, t, b) distances used at
    inference, `reg_preds_dist` the 4 * (reg_max + 1) DFL logits used by the distillation loss (training only)."""
    depth, width = cfg["depth_multiple"], cfg["width_multiple"]
    bb, nk, hd = cfg["backbone"], cfg["neck"], cfg["head"]
    reps = [(max(round(i * depth), 1) if i > 1 else i) for i in bb["num_repeats"] + nk["num_repeats"]]   # yolo.py:66
    ch = [make_divisible(i * width, 8) for i in bb["out_channels"] + nk["out_channels"]]                # yolo.py:67
    nl = hd["num_layers"]
    g = Graph(name, num_classes, list(hd["strides"]), bool(hd["use_dfl"]) and not distill_ns, 0 if distill_ns else int(hd["reg_max"]),
              cfg["training_mode"])
    if distill_ns and (fuse_ab or nl != 3):
        raise ValueError("distill_ns is the 3-level N / S student head (yolo.py:113-120); it excludes fuse_ab")
    csp = "CSP" in bb["type"]
    p6 = bb["type"].endswith("P6")
    nstage = 6 if p6 else 5
    if bb["type"] not in ("EfficientRep", "CSPBepBackbone", "CSPBepBackbone_P6"):
        raise NotImplementedError(f"backbone {bb['type']} is outside the hot-path scope (SURVEY.md section 2)")
    if nk["type"] not in ("RepBiFPANNeck", "CSPRepBiFPANNeck", "CSPRepBiFPANNeck_P6"):
        raise NotImplementedError(f"neck {nk['type']} is outside the hot-path scope (SURVEY.md section 2)")

    # ---- backbone (efficientrep.py:7-118, 250-374, 377-516) ----
    layout = "rep" if g.mode == "repvgg" else "cba"
    stem = T(g.buf(1, ch[0], "stem"), 0, ch[0])
    g.ops.append(Op("stem", "backbone.stem", layout, None, stem, 3, ch[0], 3, 2, "relu" if g.mode == "repvgg" else g.act))
    x = stem
    outs = []
    for s in range(2, nstage + 1):
        p = f"backbone.ERBlock_{s}"
        x = g.block(p + ".0", x, ch[s - 1], s=2)
        if csp:
            x = _bepc3(g, p + ".1", x, ch[s - 1], reps[s - 1], bb["csp_e"])
        else:
            x = _rep_block(g, p + ".1", x, ch[s - 1], reps[s - 1])
        if s == nstage:
            x = _cspsppf(g, p + ".2", x, ch[s - 1]) if bb.get("cspsppf") else _sppf(g, p + ".2", x, ch[s - 1])
        if s > 2 or bb.get("fuse_P2"):
            outs.append(x)
    if not bb.get("fuse_P2"):
        raise NotImplementedError("the Rep-BiFPAN necks on the hot path need fuse_P2=True backbones")

    # ---- neck (reppan.py:132-237, 666-785, 955-1116) ----
    nb = len(bb["num_repeats"])

    def stage(name, src, cout, n, dst=None):
        if csp:
            return _bepc3(g, name, src, cout, n, nk["csp_e"], dst)
        return _rep_block(g, name, src, cout, n, dst)

    if not p6:
        x3, x2, x1, x0 = outs
        lvl1, lvl0 = g.level(x1), g.level(x0)
        cat_n4 = g.buf(lvl0, ch[9] + ch[5], "neck.cat_n4")          # [down_feat0, fpn_out0]
        cat_n3 = g.buf(lvl1, ch[7] + ch[6], "neck.cat_n3")          # [down_feat1, fpn_out1]
        fpn0 = g.conv("neck.reduce_layer0", "cba", x0, ch[5], 1, 1, "relu", dst=T(cat_n4, ch[9], ch[5]))
        f0 = stage("neck.Rep_p4", _bifusion(g, "neck.Bifusion0", fpn0, x1, x2, ch[5]), ch[5], reps[nb + 0])
        fpn1 = g.conv("neck.reduce_layer1", "cba", f0, ch[6], 1, 1, "relu", dst=T(cat_n3, ch[7], ch[6]))
        pan2 = stage("neck.Rep_p3", _bifusion(g, "neck.Bifusion1", fpn1, x2, x3, ch[6]), ch[6], reps[nb + 1])
        g.conv("neck.downsample2", "cba", pan2, ch[7], 3, 2, "relu", dst=T(cat_n3, 0, ch[7]))
        pan1 = stage("neck.Rep_n3", T(cat_n3, 0, ch[7] + ch[6]), ch[8], reps[nb + 2])
        g.conv("neck.downsample1", "cba", pan1, ch[9], 3, 2, "relu", dst=T(cat_n4, 0, ch[9]))
        pan0 = stage("neck.Rep_n4", T(cat_n4, 0, ch[9] + ch[5]), ch[10], reps[nb + 3])
        feats = [pan2, pan1, pan0]
        head_ch = [ch[6], ch[8], ch[10]]                              # effidehead.py:144 chx = [6, 8, 10]
    else:
        x4, x3, x2, x1, x0 = outs
        cat_n6 = g.buf(g.level(x0), ch[10] + ch[6], "neck.cat_n6")
        cat_n5 = g.buf(g.level(x1), ch[9] + ch[7], "neck.cat_n5")
        cat_n4 = g.buf(g.level(x2), ch[8] + ch[8], "neck.cat_n4")
        fpn0 = g.conv("neck.reduce_layer0", "cba", x0, ch[6], 1, 1, "relu", dst=T(cat_n6, ch[10], ch[6]))
        f0 = stage("neck.Rep_p5", _bifusion(g, "neck.Bifusion0", fpn0, x1, x2, ch[6]), ch[6], reps[nb + 0])
        fpn1 = g.conv("neck.reduce_layer1", "cba", f0, ch[7], 1, 1, "relu", dst=T(cat_n5, ch[9], ch[7]))
        f1 = stage("neck.Rep_p4", _bifusion(g, "neck.Bifusion1", fpn1, x2, x3, ch[7]), ch[7], reps[nb + 1])
        fpn2 = g.conv("neck.reduce_layer2", "cba", f1, ch[8], 1, 1, "relu", dst=T(cat_n4, ch[8], ch[8]))
        pan3 = stage("neck.Rep_p3", _bifusion(g, "neck.Bifusion2", fpn2, x3, x4, ch[8]), ch[8], reps[nb + 2])
        g.conv("neck.downsample2", "cba", pan3, ch[8], 3, 2, "relu", dst=T(cat_n4, 0, ch[8]))
        pan2 = stage("neck.Rep_n4", T(cat_n4, 0, 2 * ch[8]), ch[9], reps[nb + 3])
        g.conv("neck.downsample1", "cba", pan2, ch[9], 3, 2, "relu", dst=T(cat_n5, 0, ch[9]))
        pan1 = stage("neck.Rep_n5", T(cat_n5, 0, ch[9] + ch[7]), ch[10], reps[nb + 4])
        g.conv("neck.downsample0", "cba", pan1, ch[10], 3, 2, "relu", dst=T(cat_n6, 0, ch[10]))
        pan0 = stage("neck.Rep_n6", T(cat_n6, 0, ch[10] + ch[6]), ch[11], reps[nb + 5])
        feats = [pan3, pan2, pan1, pan0]
        head_ch = [ch[8], ch[9], ch[10], ch[11]]                      # effidehead.py:144 chx = [8, 9, 10, 11]
    g.feat = feats
    assert len(feats) == nl

    # ---- head (effidehead.py:10-139, 142-293): stem 1x1 -> {cls 3x3 -> pred}, {reg 3x3 -> pred} ----
    reg_ch = 4 * (g.reg_max + 1)
    for i, (f, c) in enumerate(zip(feats, head_ch)):
        assert f.c == c
        st = g.conv(f"detect.stems.{i}", "cba", f, c, 1, 1, "silu")
        cf = g.conv(f"detect.cls_convs.{i}", "cba", st, c, 3, 1, "silu")
        rf = g.conv(f"detect.reg_convs.{i}", "cba", st, c, 3, 1, "silu")
        g.ops.append(Op("pred", f"detect.cls_preds.{i}", "plain", cf, None, c, num_classes, 1, 1, "sigmoid", head=("cls", i)))
        if distill_ns:      # effidehead_distill_ns.py:36-46,87-96: module order cls_preds, reg_preds_dist, reg_preds
            g.ops.append(Op("pred", f"detect.reg_preds_dist.{i}", "plain", rf, None, c, 4 * (int(hd["reg_max"]) + 1), 1, 1, None, head=("reg_dist", i)))
        g.ops.append(Op("pred", f"detect.reg_preds.{i}", "plain", rf, None, c, reg_ch, 1, 1, None, head=("reg", i)))
        if fuse_ab:     # effidehead_fuseab.py:44-55,112-118: num_anchors = 3 class / box predictions per pixel, training only
            g.ops.append(Op("pred", f"detect.cls_preds_ab.{i}", "plain", cf, None, c, num_classes * AB_ANCHORS, 1, 1, "sigmoid", head=("cls_ab", i)))
            g.ops.append(Op("pred", f"detect.reg_preds_ab.{i}", "plain", rf, None, c, 4 * AB_ANCHORS, 1, 1, None, head=("reg_ab", i)))
    g.fuse_ab = bool(fuse_ab)
    g.distill_ns = bool(distill_ns)
    g.dist_reg_ch = 4 * (int(hd["reg_max"]) + 1) if distill_ns else 0
    if fuse_ab:
        ai = hd.get("anchors_init")
        if ai is None or len(ai) != nl or any(len(a) != 2 * AB_ANCHORS for a in ai):
            raise ValueError("fuse_ab needs cfg.model.head.anchors_init with 3 (w, h) pairs per level (configs/yolov6*.py)")
        g.anchors_init = [[float(v) for v in a] for a in ai]
    return g


def param_specs(g):
    """(name, shape, kind) of every tensor in the reference state_dict that this graph owns.
    kind in {'conv', 'bn', 'bias', 'alpha', 'buffer', 'const'}; BN expands to its five tensors."""
    specs = []

    def bn(prefix, c):
        specs.append((prefix + ".weight", (c,), "bn_w"))
        specs.append((prefix + ".bias", (c,), "bn_b"))
        specs.append((prefix + ".running_mean", (c,), "bn_mean"))
        specs.append((prefix + ".running_var", (c,), "bn_var"))
        specs.append((prefix + ".num_batches_tracked", (), "bn_count"))

    seen_alpha = set()
    for op in g.ops:
        n = op.name
        if op.kind == "pool":
            continue
        if op.alpha and op.alpha not in seen_alpha:
            seen_alpha.add(op.alpha)
            specs.append((op.alpha, (1,), "alpha"))
        if op.layout == "rep":
            if op.cin == op.cout and op.s == 1:
                bn(n + ".rbr_identity", op.cin)
            specs.append((n + ".rbr_dense.conv.weight", (op.cout, op.cin, 3, 3), "conv"))
            bn(n + ".rbr_dense.bn", op.cout)
            specs.append((n + ".rbr_1x1.conv.weight", (op.cout, op.cin, 1, 1), "conv"))
            bn(n + ".rbr_1x1.bn", op.cout)
        elif op.layout == "cba":
            specs.append((n + ".block.conv.weight", (op.cout, op.cin, op.k, op.k), "conv"))
            bn(n + ".block.bn", op.cout)
        elif op.layout == "plain":
            specs.append((n + ".weight", (op.cout, op.cin, 1, 1), "conv"))
            specs.append((n + ".bias", (op.cout,), "bias"))
        elif op.layout == "convT":
            specs.append((n + ".upsample_transpose.weight", (op.cin, op.cout, 2, 2), "conv"))
            specs.append((n + ".upsample_transpose.bias", (op.cout,), "bias"))
    # build_network does not forward reg_max to Detect (yolo.py:130-131), so its DFL projection always
    # has Detect's default reg_max = 16 entries even for the N/S models that predict 4 reg channels.
    specs.append(("detect.proj", (DETECT_DEFAULT_REG_MAX + 1,), "const"))
    specs.append(("detect.proj_conv.weight", (1, DETECT_DEFAULT_REG_MAX + 1, 1, 1), "const"))
    return specs
