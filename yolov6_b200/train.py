"""Training engine: train-form forward + backward of the conv stack on sm_100a kernels.

What autograd + cuDNN + ~130 BatchNorm / activation kernels do for the reference's
`Trainer.train_in_steps` (core/engine.py:142-176) over `Model.forward` in train mode
(models/yolo.py:33-41; ConvModule conv->BN->act, layers/common.py:46-49; RepVGGBlock's three BN-ed
branches, common.py:245-255), this engine does by walking the same layer graph as the inference
engine (arch.py):

  forward : raw convs on tcgen05 (yv6_conv_fwd, bf16 operands, fp32 accumulate, no bias/act)
            -> yv6_bn_stats / yv6_bn_finalize (batch statistics, running-stat update, eps 1e-3,
            momentum 0.03 as set by initialize_weights, torch_utils.py:38-48)
            -> yv6_bn_apply_fwd (sum of the BN-ed branches + activation, written into concat slices);
  backward: yv6_bn_bwd (activation + BatchNorm backward of all branches of a block in two passes)
            -> dgrad = yv6_conv_fwd with rotated / transposed weights, accumulating into the input
               gradient through the residual epilogue (stride-2 convs: four parity sub-convolutions)
            -> yv6_conv_wgrad (MN-major tcgen05 GEMM over pixels, fp32 split-K accumulation).

Parameters stay fp32 `nn.Parameter`s of the Model (master weights); every step they are cast to bf16
KRSC for the kernels, and gradients are written to `.grad` in the reference's tensor layouts so that the
reference's optimizer / EMA / DDP all-reduce apply unchanged.  BottleRep shortcuts (M / L6,
common.py:600-617) ride in the BN apply / backward kernels (y = act(z) + alpha * x).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_CODES, DT_BF16, DT_F32, DT_U8, BnDesc, ConvDesc, StemDesc, WgradDesc

BN_EPS, BN_MOMENTUM = 1e-3, 0.03


def _p(t):
    return t.data_ptr() if t is not None else 0


class TrainEngine:
    def __init__(self, model):
        self.model = model
        self.g = model.graph
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("yolov6_b200 training runs on sm_100a CUDA kernels only (no CPU fallback)")
        self.lib = _lib.lib()
        self.h = _lib.handle(self.dev.index or 0)
        self.params = dict(model.named_parameters())
        self.buffers_ = dict(model.named_buffers())
        self._shape = None
        self.debug = False      # tests: snapshot the incoming gradient of every op into self.dbg[op index]
        self.dbg = {}

    # ------------------------------------------------------------------ helpers
    def _conv(self, x, x_off, cin, w, y, y_off, cout, k, stride, *, pad=None, out_hw=None, bias=None, act=None,
              y_strides=None, y_elem_off=0, accumulate=False, y_f32=False):
        """y[..., y_off:+cout] (+)= conv(x[..., x_off:+cin], w) via the C ABI.  x: [N,H,W,Ct] bf16,
        w: [cout,kh,kw,cin] bf16 KRSC, y: [N,Ho,Wo,Cyt] bf16 (or fp32 head tensor with explicit strides)."""
        d = ConvDesc()
        N, H, W, Ct = x.shape
        d.x = x.data_ptr() + x_off * 2
        d.N, d.H, d.W, d.Cin, d.x_c_total = N, H, W, cin, Ct
        d.w = w.data_ptr()
        d.bias = _p(bias)
        d.Cout, d.kh, d.kw, d.stride = cout, w.shape[1], w.shape[2], stride
        d.pad, d.pad_w = (k // 2, _lib.PAD_SAME) if pad is None else pad
        if out_hw is not None:
            d.out_h, d.out_w = out_hw
        d.act = ACT_CODES[act]
        d.nsplit = 1
        es = 4 if y_f32 else 2
        d.y = y.data_ptr() + (y_off + y_elem_off) * es
        d.y_dtype = DT_F32 if y_f32 else DT_BF16
        if y_strides is None:
            _, Ho, Wo, Cyt = y.shape
            y_strides = (Ho * Wo * Cyt, Wo * Cyt, Cyt)
        d.y_img_stride, d.y_h_stride, d.y_w_stride = y_strides
        if accumulate:   # y += conv(...): the residual epilogue reads the old value of the same element
            d.res = d.y
            d.alpha = 1.0
            d.res_img_stride, d.res_h_stride, d.res_w_stride = y_strides
        _lib.check(self.lib.yv6_conv_fwd(self.h, C.byref(d), _lib.stream_ptr()))

    def _wgrad(self, x, x_off, cin, dy, dy_off, cout, k, stride, dw):
        d = WgradDesc()
        N, H, W, Ct = x.shape
        d.x = x.data_ptr() + x_off * 2
        d.N, d.H, d.W, d.Cin, d.x_c_total = N, H, W, cin, Ct
        d.dy = dy.data_ptr() + dy_off * 2
        d.Cout, d.dy_c_total = cout, dy.shape[3]
        d.kh = d.kw = k
        d.stride, d.pad = stride, k // 2
        d.dw = dw.data_ptr()
        _lib.check(self.lib.yv6_conv_wgrad(self.h, C.byref(d), _lib.stream_ptr()))

    def _stats(self, t, c_off, c):
        s = torch.empty(2, c, dtype=torch.float64, device=self.dev)
        N, H, W, Ct = t.shape
        _lib.check(self.lib.yv6_bn_stats(self.h, t.data_ptr() + c_off * 2, N * H * W, c, Ct, s[0].data_ptr(), s[1].data_ptr(),
                                         _lib.stream_ptr()))
        return s

    def _finalize(self, stats, count, prefix, c):
        """-> dict(mean, invstd, scale, shift) fp32 [C]; updates the running stats of BatchNorm `prefix`."""
        out = torch.empty(4, c, dtype=torch.float32, device=self.dev)
        g, b = self.params[prefix + ".weight"], self.params[prefix + ".bias"]
        rm, rv = self.buffers_[prefix + ".running_mean"], self.buffers_[prefix + ".running_var"]
        _lib.check(self.lib.yv6_bn_finalize(self.h, stats[0].data_ptr(), stats[1].data_ptr(), float(count), g.data_ptr(), b.data_ptr(),
                                            BN_EPS, BN_MOMENTUM, rm.data_ptr(), rv.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                            out[2].data_ptr(), out[3].data_ptr(), c, _lib.stream_ptr()))
        self.buffers_[prefix + ".num_batches_tracked"].add_(1)
        return out

    @staticmethod
    def _krsc(w):
        """torch conv weight [Cout,Cin,kh,kw] fp32 -> bf16 KRSC."""
        return w.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()

    # ------------------------------------------------------------------ buffers
    def _alloc(self, N, H, W):
        if self._shape == (N, H, W):
            return
        self._shape = (N, H, W)
        g, dev = self.g, self.dev
        self.bufs = [torch.zeros(N, H >> b.level, W >> b.level, b.c_total, dtype=torch.bfloat16, device=dev) for b in g.bufs]
        self.gbufs = [torch.zeros_like(t) for t in self.bufs]
        self.sizes = [(H // s, W // s) for s in g.strides]
        self.offs = [0]
        for h, w in self.sizes:
            self.offs.append(self.offs[-1] + h * w)
        A = self.offs[-1]
        self.cls = torch.empty(N, A, g.num_classes, dtype=torch.float32, device=dev)
        self.reg = torch.empty(N, A, 4 * (g.reg_max + 1), dtype=torch.float32, device=dev)

    def _view(self, t):
        return self.bufs[t.buf], self.gbufs[t.buf]

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        """x [N,3,H,W] fp32 in [0,1] (or uint8).  Returns (cls [N,A,nc] post-sigmoid, reg [N,A,R]) fp32."""
        x = x.contiguous()
        N, _, H, W = x.shape
        self._alloc(N, H, W)
        self.x = x
        self.ctx = []
        P = self.params
        for i, op in enumerate(self.g.ops):
            if op.kind == "pool":
                buf, _ = self._view(op.dst)
                n, h, w, ct = buf.shape
                _lib.check(self.lib.yv6_sppf_pool(self.h, buf.data_ptr(), n, h, w, op.cin, ct, 1, 0, _lib.stream_ptr()))
                self.ctx.append(None)
                continue
            if op.kind == "pred":
                src, _ = self._view(op.src)
                which, lvl = op.head
                out = self.cls if which == "cls" else self.reg
                ch = out.shape[2]
                lh, lw = self.sizes[lvl]
                A = self.offs[-1]
                w = self._krsc(P[op.name + ".weight"])
                bias = torch.zeros((op.cout + 255) // 256 * 256, dtype=torch.float32, device=self.dev)
                bias[:op.cout] = P[op.name + ".bias"].detach()
                self._conv(src, op.src.c_off, op.cin, w, out, 0, op.cout, 1, 1, bias=bias, act=op.act, y_f32=True,
                           y_strides=(A * ch, lw * ch, ch), y_elem_off=self.offs[lvl] * ch)
                self.ctx.append(dict(w=w))
                continue
            if op.kind == "convT":
                src, _ = self._view(op.src)
                dst, _ = self._view(op.dst)
                wt = P[op.name + ".upsample_transpose.weight"].detach()          # [Cin, Cout, 2, 2]
                bias = torch.zeros((op.cout + 255) // 256 * 256, dtype=torch.float32, device=self.dev)
                bias[:op.cout] = P[op.name + ".upsample_transpose.bias"].detach()
                _, dh, dw, dct = dst.shape
                ws = []
                for q in range(4):
                    dy, dx = q // 2, q % 2
                    wq = wt[:, :, dy, dx].t().reshape(op.cout, 1, 1, op.cin).to(torch.bfloat16).contiguous()
                    ws.append(wq)
                    self._conv(src, op.src.c_off, op.cin, wq, dst, op.dst.c_off, op.cout, 1, 1, bias=bias,
                               y_strides=(dh * dw * dct, 2 * dw * dct, 2 * dct), y_elem_off=(dy * dw + dx) * dct)
                self.ctx.append(dict(w=ws))
                continue
            # ---- BN-ed blocks: stem / rep / cba ----
            dst, _ = self._view(op.dst)
            n, ho, wo, _ = dst.shape
            count = n * ho * wo
            branches = []   # (raw tensor [N,Ho,Wo,C], c_off, bn prefix, conv info)
            if op.layout == "rep":
                specs = [(op.name + ".rbr_dense", 3), (op.name + ".rbr_1x1", 1)]
            else:
                specs = [(op.name + ".block", op.k)]
            for prefix, k in specs:
                raw = torch.empty(n, ho, wo, op.cout, dtype=torch.bfloat16, device=self.dev)
                wt = P[prefix + ".conv.weight"].detach()
                if op.kind == "stem":
                    w33 = wt if k == 3 else torch.nn.functional.pad(wt, [1, 1, 1, 1])      # 1x1 s2 = centre tap of a 3x3 s2
                    wdev = w33.permute(2, 3, 1, 0).contiguous().float()                     # [3][3][3][Cout]
                    d = StemDesc()
                    d.x, d.x_dtype, d.in_scale = self.x.data_ptr(), (DT_U8 if self.x.dtype == torch.uint8 else DT_F32), 1.0 / 255.0
                    d.N, d.H, d.W = N, H, W
                    d.w, d.bias, d.Cout, d.act = wdev.data_ptr(), 0, op.cout, 0
                    d.y, d.y_plane_stride, d.nsplit, d.fp32_math = raw.data_ptr(), 0, 1, 1
                    _lib.check(self.lib.yv6_stem_fwd(self.h, C.byref(d), _lib.stream_ptr()))
                    wk = wdev
                else:
                    src, _ = self._view(op.src)
                    wk = self._krsc(wt)
                    self._conv(src, op.src.c_off, op.cin, wk, raw, 0, op.cout, k, op.s)
                st = self._finalize(self._stats(raw, 0, op.cout), count, prefix + ".bn", op.cout)
                branches.append(dict(x=raw, off=0, st=st, prefix=prefix, k=k, w=wk))
            if op.layout == "rep" and op.cin == op.cout and op.s == 1:
                src, _ = self._view(op.src)
                st = self._finalize(self._stats(src, op.src.c_off, op.cin), count, op.name + ".rbr_identity", op.cin)
                branches.append(dict(x=src, off=op.src.c_off, st=st, prefix=op.name + ".rbr_identity", k=0, w=None))
            d = BnDesc()
            d.nb, d.act, d.C, d.pixels = len(branches), ACT_CODES[op.act], op.cout, count
            for b, br in enumerate(branches):
                d.x[b] = br["x"].data_ptr() + br["off"] * 2
                d.x_pitch[b] = br["x"].shape[3]
                d.scale[b], d.shift[b] = br["st"][2].data_ptr(), br["st"][3].data_ptr()
            d.y, d.y_pitch = dst.data_ptr() + op.dst.c_off * 2, dst.shape[3]
            alpha = 1.0
            if op.res is not None:
                rbuf, _ = self._view(op.res)
                alpha = float(P[op.alpha].detach()) if op.alpha in P else 1.0
                d.res, d.res_pitch, d.res_alpha = rbuf.data_ptr() + op.res.c_off * 2, rbuf.shape[3], alpha
            _lib.check(self.lib.yv6_bn_apply_fwd(self.h, C.byref(d), _lib.stream_ptr()))
            self.ctx.append(dict(branches=branches, count=count, alpha=alpha))
        return self.cls, self.reg

    # ------------------------------------------------------------------ backward
    def _add_grad(self, name, value):
        p = self.params[name]
        value = value.to(p.dtype).reshape(p.shape)
        p.grad = value.clone() if p.grad is None else p.grad + value

    def _dgrad(self, dc, w_krsc, k, stride, gsrc, g_off, cin):
        """g(src)[..., g_off:+cin] += conv_transpose(dc, w).  dc [N,Ho,Wo,Cout] bf16, w [Cout,k,k,Cin] bf16."""
        cout = w_krsc.shape[0]
        if stride == 1:
            wt = w_krsc.flip(1, 2).permute(3, 1, 2, 0).contiguous()            # [Cin, k, k, Cout], rotated 180 degrees
            self._conv(dc, 0, cout, wt, gsrc, g_off, cin, k, 1, accumulate=True)
            return
        # stride 2: the input gradient at parity (ph, pw) is a 1- or 2-tap stride-1 conv of dc
        n, hs, ws, gct = gsrc.shape
        ho, wo = dc.shape[1], dc.shape[2]
        for ph in range(2):
            for pw in range(2):
                if k == 1:
                    if ph or pw:
                        continue                                               # a 1x1 s2 conv only touches even positions
                    wt = w_krsc.permute(3, 1, 2, 0).contiguous()
                else:
                    rows = [1] if ph == 0 else [2, 0]                          # tap t reads dc[i + t]: W[2] at t=0, W[0] at t=1
                    cols = [1] if pw == 0 else [2, 0]
                    wt = w_krsc[:, rows][:, :, cols].permute(3, 1, 2, 0).contiguous()   # [Cin, kh', kw', Cout]
                self._conv(dc, 0, cout, wt, gsrc, g_off, cin, k, 1, pad=(0, 0), out_hw=(ho, wo), accumulate=True,
                           y_strides=(hs * ws * gct, 2 * ws * gct, 2 * gct), y_elem_off=(ph * ws + pw) * gct)

    def backward(self, grad_cls, grad_reg):
        """Accumulates d(loss)/d(parameter) into `.grad` given d(loss)/d(cls), d(loss)/d(reg) ([N,A,*] fp32)."""
        P, dev = self.params, self.dev
        for gb in self.gbufs:
            gb.zero_()
        N = self.cls.shape[0]
        A = self.offs[-1]
        for i in range(len(self.g.ops) - 1, -1, -1):
            op, ctx = self.g.ops[i], self.ctx[i]
            if op.kind == "pool":
                buf, gbuf = self._view(op.dst)
                n, h, w, ct = buf.shape
                c = op.cin
                if self.debug:
                    self.dbg[i] = dict(gdst=gbuf[..., :4 * c].clone())
                scratch = torch.empty(n, h, w, c, dtype=torch.float32, device=dev)
                for j in (3, 2, 1):   # y_j = pool(y_{j-1}); slice j of the concat
                    _lib.check(self.lib.yv6_maxpool5_bwd(self.h, buf.data_ptr() + (j - 1) * c * 2, ct, gbuf.data_ptr() + j * c * 2, ct,
                                                         n, h, w, c, scratch.data_ptr(), gbuf.data_ptr() + (j - 1) * c * 2, ct, 1,
                                                         _lib.stream_ptr()))
                continue
            if op.kind == "pred":
                which, lvl = op.head
                grad, scores = (grad_cls, self.cls) if which == "cls" else (grad_reg, None)
                ch = grad.shape[2]
                ch_pad = (ch + 15) // 16 * 16
                lh, lw = self.sizes[lvl]
                dl = torch.empty(N, lh, lw, ch_pad, dtype=torch.bfloat16, device=dev)
                _lib.check(self.lib.yv6_head_grad_prep(self.h, grad.data_ptr(), _p(scores), N, A, ch, self.offs[lvl], lh * lw, ch_pad,
                                                       dl.data_ptr(), _lib.stream_ptr()))
                src, gsrc = self._view(op.src)
                dw = torch.zeros(ch, 1, 1, op.cin, dtype=torch.float32, device=dev)
                self._wgrad(src, op.src.c_off, op.cin, dl, 0, ch, 1, 1, dw)
                self._add_grad(op.name + ".weight", dw.permute(0, 3, 1, 2))
                self._add_grad(op.name + ".bias", self._stats(dl, 0, ch_pad)[0][:ch])
                wpad = torch.zeros(ch_pad, 1, 1, op.cin, dtype=torch.bfloat16, device=dev)
                wpad[:ch] = ctx["w"]
                self._dgrad(dl, wpad, 1, 1, gsrc, op.src.c_off, op.cin)
                continue
            if op.kind == "convT":
                src, gsrc = self._view(op.src)
                _, gdst = self._view(op.dst)
                gd = gdst[..., op.dst.c_off:op.dst.c_off + op.cout]
                if self.debug:
                    self.dbg[i] = dict(gdst=gd.clone())
                dwt = torch.zeros(op.cin, op.cout, 2, 2, dtype=torch.float32, device=dev)
                db = torch.zeros(op.cout, dtype=torch.float64, device=dev)
                for q in range(4):
                    dy, dx = q // 2, q % 2
                    dq = gd[:, dy::2, dx::2, :].contiguous()                      # gradient of quadrant q, dense [N,H,W,Cout]
                    dw = torch.zeros(op.cout, 1, 1, op.cin, dtype=torch.float32, device=dev)
                    self._wgrad(src, op.src.c_off, op.cin, dq, 0, op.cout, 1, 1, dw)
                    dwt[:, :, dy, dx] = dw.reshape(op.cout, op.cin).t()
                    db += self._stats(dq, 0, op.cout)[0]
                    self._dgrad(dq, ctx["w"][q], 1, 1, gsrc, op.src.c_off, op.cin)
                self._add_grad(op.name + ".upsample_transpose.weight", dwt)
                self._add_grad(op.name + ".upsample_transpose.bias", db)
                continue
            # ---- BN-ed blocks ----
            dst, gdst = self._view(op.dst)
            brs = ctx["branches"]
            n, ho, wo, dct = dst.shape
            d = BnDesc()
            d.nb, d.act, d.C, d.pixels = len(brs), ACT_CODES[op.act], op.cout, ctx["count"]
            sums = torch.empty(4, op.cout, dtype=torch.float64, device=dev)
            d.s1 = sums[0].data_ptr()
            dcs = []
            for b, br in enumerate(brs):
                d.x[b], d.x_pitch[b] = br["x"].data_ptr() + br["off"] * 2, br["x"].shape[3]
                d.mean[b], d.invstd[b] = br["st"][0].data_ptr(), br["st"][1].data_ptr()
                d.scale[b], d.shift[b] = br["st"][2].data_ptr(), br["st"][3].data_ptr()
                d.s2[b] = sums[1 + b].data_ptr()
                if br["k"] == 0:      # identity branch: its input gradient goes straight into g(src)
                    _, gsrc = self._view(op.src)
                    d.dx[b], d.dx_pitch[b], d.accumulate[b] = gsrc.data_ptr() + op.src.c_off * 2, gsrc.shape[3], 1
                    dcs.append(None)
                else:
                    dc = torch.empty(n, ho, wo, op.cout, dtype=torch.bfloat16, device=dev)
                    d.dx[b], d.dx_pitch[b], d.accumulate[b] = dc.data_ptr(), op.cout, 0
                    dcs.append(dc)
            d.y, d.y_pitch = dst.data_ptr() + op.dst.c_off * 2, dct
            d.dy, d.dy_pitch = gdst.data_ptr() + op.dst.c_off * 2, dct
            if op.res is not None:
                rbuf, grbuf = self._view(op.res)
                dalpha = torch.empty(1, dtype=torch.float64, device=dev)
                d.res, d.res_pitch, d.res_alpha = rbuf.data_ptr() + op.res.c_off * 2, rbuf.shape[3], ctx["alpha"]
                d.dres, d.dres_pitch, d.dalpha = grbuf.data_ptr() + op.res.c_off * 2, grbuf.shape[3], dalpha.data_ptr()
            _lib.check(self.lib.yv6_bn_bwd(self.h, C.byref(d), _lib.stream_ptr()))
            if op.res is not None and op.alpha in P and P[op.alpha].requires_grad:
                self._add_grad(op.alpha, dalpha)
            if self.debug:
                self.dbg[i] = dict(gdst=gdst[..., op.dst.c_off:op.dst.c_off + op.cout].clone(), dcs=dcs)
            for b, br in enumerate(brs):
                bnp = br["prefix"] + (".bn" if br["k"] else "")
                self._add_grad(bnp + ".weight", sums[1 + b])        # dgamma = sum dz * xhat
                self._add_grad(bnp + ".bias", sums[0])              # dbeta  = sum dz
                if br["k"] == 0:
                    continue
                dc, k = dcs[b], br["k"]
                if op.kind == "stem":
                    dw = torch.empty(op.cout, 3, 3, 3, dtype=torch.float32, device=dev)
                    _lib.check(self.lib.yv6_stem_wgrad(self.h, self.x.data_ptr(), DT_U8 if self.x.dtype == torch.uint8 else DT_F32,
                                                       1.0 / 255.0, dc.data_ptr(), op.cout, N, self.x.shape[2], self.x.shape[3], op.cout,
                                                       dw.data_ptr(), _lib.stream_ptr()))
                    dw = dw.permute(0, 3, 1, 2)                     # [Cout][r][s][c] -> [Cout, c, r, s]
                    self._add_grad(br["prefix"] + ".conv.weight", dw if k == 3 else dw[:, :, 1:2, 1:2])
                    continue
                src, gsrc = self._view(op.src)
                dw = torch.zeros(op.cout, k, k, op.cin, dtype=torch.float32, device=dev)
                self._wgrad(src, op.src.c_off, op.cin, dc, 0, op.cout, k, op.s, dw)
                self._add_grad(br["prefix"] + ".conv.weight", dw.permute(0, 3, 1, 2))
                self._dgrad(dc, br["w"], k, op.s, gsrc, op.src.c_off, op.cin)


class _HeadFn(torch.autograd.Function):
    """Connects the engine to autograd: forward returns the head tensors, backward runs the engine's
    backward pass with the incoming gradients (parameter gradients land in `.grad`)."""

    @staticmethod
    def forward(ctx, engine, x, token):
        ctx.engine = engine
        cls, reg = engine.forward(x)
        return cls.clone(), reg.clone()

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        ctx.engine.backward(g_cls.contiguous().float(), g_reg.contiguous().float())
        return None, None, None


def train_forward(engine, x):
    token = torch.zeros(1, device=engine.dev, requires_grad=True)   # makes the outputs part of the autograd graph
    return _HeadFn.apply(engine, x, token)
