"""Inference engine: walks the layer graph (arch.py) and issues one sm_100a kernel per op.

What the reference does with nested nn.Module.forward calls, cuDNN and ~15 eager ops for the head
decode (Model.forward yolo.py:33-41, Detect.forward effidehead.py:93-139), this does with:
  * folded deploy-form weights (fold.py) packed once to KRSC bf16 (or three bf16 planes),
  * pre-allocated NHWC activation buffers (concats are channel slices),
  * a prepared list of C-ABI descriptors per input shape, replayed per batch -- optionally as one
    captured CUDA graph (no Python / launch overhead in steady state).
Precision modes: "bf16" (bf16 operands, fp32 accumulate) and "fp32" (bf16x3 split operands: fp32-
equivalent products, used for the 1e-4 parity bar of BASELINE.json).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ACT_CODES, DT_BF16, DT_F32, DT_U8, ConvDesc, StemDesc
from .fold import fold_op


def _split3(t):
    t = t.float()
    p0 = t.to(torch.bfloat16)
    r1 = t - p0.float()
    p1 = r1.to(torch.bfloat16)
    p2 = (r1 - p1.float()).to(torch.bfloat16)
    return torch.stack([p0, p1, p2])


class InferEngine:
    def __init__(self, graph, state_dict, device, precision="bf16"):
        if device.type != "cuda":
            raise RuntimeError("yolov6_b200 runs its networks on sm_100a CUDA kernels only (no CPU fallback); "
                               f"got device {device}")
        assert precision in ("bf16", "fp32")
        self.g = graph
        self.device = device
        self.precision = precision
        self.nsplit = 3 if precision == "fp32" else 1
        self.handle = _lib.handle(device.index or 0)
        self.lib = _lib.lib()
        self.weights = {}     # op index -> dict(w=..., bias=..., alpha=...)
        self._plans = {}      # (N, H, W, dtype) -> plan, least recently used first; bounded (rect-shaped evaluation
        self.max_plans = 4    # would otherwise keep one full activation set per distinct shape)
        self._pinned = set()  # shapes captured into CUDA graphs (their buffers must outlive the graph)
        self._pack(state_dict)

    # ------------------------------------------------------------------ weight packing
    def _pack(self, sd):
        sd = {k: v.detach().cpu() for k, v in sd.items()}
        dev = self.device
        for i, op in enumerate(self.g.ops):
            if op.kind == "pool":
                continue
            w, b = fold_op(sd, op)
            ent = {}
            if op.kind == "stem":
                ent["w_dev"] = w.float().permute(1, 2, 3, 0).contiguous().to(dev)       # [3][3][3][Cout]
                ent["b_dev"] = b.float().contiguous().to(dev)
            else:
                ws = w if isinstance(w, list) else [w]
                packed = []
                for wi in ws:
                    wi = wi.float().to(dev)
                    packed.append(_split3(wi).contiguous() if self.nsplit == 3 else wi.to(torch.bfloat16).contiguous())
                ent["w"] = packed
                bias = torch.zeros((op.cout + 255) // 256 * 256, dtype=torch.float32, device=dev)
                bias[:op.cout] = b.float().to(dev)
                ent["bias"] = bias
            ent["alpha"] = float(sd[op.alpha]) if (op.alpha and op.res is not None) else 1.0
            self.weights[i] = ent

    # ------------------------------------------------------------------ per-shape plan
    def _plan(self, N, H, W, in_dtype):
        key = (N, H, W, in_dtype)
        if key in self._plans:
            plan = self._plans.pop(key)
            self._plans[key] = plan          # most recently used last
            return plan
        while len(self._plans) >= self.max_plans:
            victim = next((k for k in self._plans if k not in self._pinned), None)
            if victim is None:
                break
            del self._plans[victim]
        g, dev, P = self.g, self.device, self.nsplit
        maxs = max(g.strides)
        if H % maxs or W % maxs:
            raise RuntimeError(f"input {H}x{W} must be a multiple of the largest stride {maxs}")
        plan = {"bufs": [], "calls": []}
        for b in g.bufs:
            h, w = H >> b.level, W >> b.level
            shape = (P, N, h, w, b.c_total) if P == 3 else (N, h, w, b.c_total)
            plan["bufs"].append(torch.zeros(shape, dtype=torch.bfloat16, device=dev))
        sizes = [(H // s, W // s) for s in g.strides]
        offs = np.concatenate([[0], np.cumsum([h * w for h, w in sizes])]).astype(int)
        A = int(offs[-1])
        nc, R = g.num_classes, 4 * (g.reg_max + 1)
        plan["cls"] = torch.empty(N, A, nc, dtype=torch.float32, device=dev)
        plan["reg"] = torch.empty(N, A, R, dtype=torch.float32, device=dev)
        plan["pred"] = torch.empty(N, A, 5 + nc, dtype=torch.float32, device=dev)
        plan["sizes"], plan["A"] = sizes, A
        plan["lvl_h"] = (C.c_int32 * len(sizes))(*[h for h, _ in sizes])
        plan["lvl_w"] = (C.c_int32 * len(sizes))(*[w for _, w in sizes])
        plan["lvl_s"] = (C.c_float * len(sizes))(*[float(s) for s in g.strides])
        plan["image"] = None

        def view(t):
            buf = plan["bufs"][t.buf]
            b = g.bufs[t.buf]
            h, w = H >> b.level, W >> b.level
            return buf, h, w, b.c_total

        for i, op in enumerate(g.ops):
            ent = self.weights.get(i)
            if op.kind == "stem":
                buf, h, w, ct = view(op.dst)
                d = StemDesc()
                d.x_dtype = DT_U8 if in_dtype == torch.uint8 else DT_F32
                d.in_scale = 1.0 / 255.0
                d.N, d.H, d.W = N, H, W
                d.w = ent["w_dev"].data_ptr()
                d.bias = ent["b_dev"].data_ptr()
                d.Cout, d.act = op.cout, ACT_CODES[op.act]
                d.y = buf.data_ptr()
                d.y_plane_stride = buf.stride(0) if P == 3 else 0
                d.nsplit = P
                plan["stem"] = d
                plan["calls"].append(("stem", d))
            elif op.kind in ("conv", "pred", "convT"):
                sbuf, sh, sw, sct = view(op.src)
                quads = ent["w"] if op.kind == "convT" else [ent["w"][0]]
                for q, wq in enumerate(quads):
                    d = ConvDesc()
                    d.x = sbuf.data_ptr() + op.src.c_off * 2
                    d.N, d.H, d.W, d.Cin, d.x_c_total = N, sh, sw, op.cin, sct
                    d.x_plane_stride = sbuf.stride(0) if P == 3 else 0
                    d.w = wq.data_ptr()
                    d.w_plane_stride = wq.stride(0) if P == 3 else 0
                    d.bias = ent["bias"].data_ptr()
                    d.Cout = op.cout
                    d.kh = d.kw = 1 if op.kind == "convT" else op.k
                    d.stride = 1 if op.kind == "convT" else op.s
                    d.pad = d.kh // 2
                    d.pad_w = _lib.PAD_SAME
                    d.act = ACT_CODES[op.act]
                    d.nsplit = P
                    if op.kind == "pred":
                        which, lvl = op.head
                        out = plan[which]
                        ch = out.shape[2]
                        lh, lw = sizes[lvl]
                        d.y = out.data_ptr() + int(offs[lvl]) * ch * 4
                        d.y_dtype = DT_F32
                        d.y_img_stride, d.y_h_stride, d.y_w_stride = A * ch, lw * ch, ch
                    else:
                        dbuf, dh, dw, dct = view(op.dst)
                        d.y_dtype = DT_BF16
                        d.y_plane_stride = dbuf.stride(0) if P == 3 else 0
                        if op.kind == "convT":   # scatter quadrant (dy, dx) of the 2x upsample
                            dy, dx = q // 2, q % 2
                            d.y = dbuf.data_ptr() + ((dy * dw + dx) * dct + op.dst.c_off) * 2
                            d.y_img_stride, d.y_h_stride, d.y_w_stride = dh * dw * dct, 2 * dw * dct, 2 * dct
                        else:
                            d.y = dbuf.data_ptr() + op.dst.c_off * 2
                            d.y_img_stride, d.y_h_stride, d.y_w_stride = dh * dw * dct, dw * dct, dct
                    if op.res is not None:
                        rbuf, rh, rw, rct = view(op.res)
                        d.res = rbuf.data_ptr() + op.res.c_off * 2
                        d.res_img_stride, d.res_h_stride, d.res_w_stride = rh * rw * rct, rw * rct, rct
                        d.res_plane_stride = rbuf.stride(0) if P == 3 else 0
                        d.alpha = ent["alpha"]
                    plan["calls"].append(("conv", d))
            elif op.kind == "pool":
                buf, h, w, ct = view(op.dst)
                plan["calls"].append(("pool", (buf.data_ptr(), N, h, w, op.cin, ct, P, buf.stride(0) if P == 3 else 0)))
        self._plans[key] = plan
        return plan

    def pin(self, N, H, W, in_dtype=torch.float32):
        """Keep the buffers of this shape for the engine's lifetime (they are referenced by a captured CUDA graph)."""
        self._plan(N, H, W, in_dtype)
        self._pinned.add((N, H, W, in_dtype))

    # ------------------------------------------------------------------ execution
    def launch_count(self, N, H, W, in_dtype=torch.float32):
        """Kernels launched per forward for this shape (convs + stem + pools + decode)."""
        return len(self._plan(N, H, W, in_dtype)["calls"]) + 1

    def forward(self, x, stream=None):
        """x: [N,3,H,W] CUDA tensor, fp32 in [0,1] or uint8.  Returns pred [N,A,5+nc] fp32 (a buffer owned
        by the engine, overwritten by the next call with the same shape)."""
        if x.device != self.device:
            raise RuntimeError(f"input on {x.device}, engine on {self.device}")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        x = x.contiguous()
        N, Cin, H, W = x.shape
        assert Cin == 3
        plan = self._plan(N, H, W, x.dtype)
        plan["image"] = x  # keep alive while kernels are in flight
        sp = _lib.stream_ptr(stream)
        lib, h, chk = self.lib, self.handle, _lib.check
        plan["stem"].x = x.data_ptr()
        for kind, d in plan["calls"]:
            if kind == "conv":
                chk(lib.yv6_conv_fwd(h, C.byref(d), sp))
            elif kind == "stem":
                chk(lib.yv6_stem_fwd(h, C.byref(d), sp))
            else:
                chk(lib.yv6_sppf_pool(h, C.c_void_p(d[0]), d[1], d[2], d[3], d[4], d[5], d[6], d[7], sp))
        g = self.g
        chk(lib.yv6_head_decode(h, C.c_void_p(plan["cls"].data_ptr()), C.c_void_p(plan["reg"].data_ptr()),
                                C.c_void_p(plan["pred"].data_ptr()), N, g.num_classes, 4 * (g.reg_max + 1),
                                len(g.strides), plan["lvl_h"], plan["lvl_w"], plan["lvl_s"], sp))
        return plan["pred"]

    def head_outputs(self, N, H, W, in_dtype=torch.float32):
        """(cls [N,A,nc] post-sigmoid, reg [N,A,R] raw) of the last forward with this shape."""
        plan = self._plan(N, H, W, in_dtype)
        return plan["cls"], plan["reg"]

    def feature_maps(self, N, H, W, in_dtype=torch.float32):
        """Neck outputs as NCHW views (reference `featmaps`, yolo.py:37-39); bf16 (hi plane in fp32 mode)."""
        plan = self._plan(N, H, W, in_dtype)
        out = []
        for t in self.g.feat:
            buf = plan["bufs"][t.buf]
            buf = buf[0] if self.nsplit == 3 else buf
            out.append(buf[..., t.c_off:t.c_off + t.c].permute(0, 3, 1, 2))
        return out

    def profile_convs(self, x, steps=5, stream=None):
        """Roofline instrumentation for bench.py: CUDA events around every conv_igemm launch.
        Returns (conv-kernel ms per forward, algorithmic conv FLOPs per forward, launches per forward)."""
        x = x.contiguous()
        N, _, H, W = x.shape
        plan = self._plan(N, H, W, x.dtype if x.dtype == torch.uint8 else torch.float32)
        self.forward(x, stream)
        torch.cuda.synchronize()
        sp = _lib.stream_ptr(stream)
        convs = [d for kind, d in plan["calls"] if kind == "conv"]
        flops = 0.0
        for d in convs:
            ho = (d.H + 2 * d.pad - d.kh) // d.stride + 1
            wo = (d.W + 2 * d.pad - d.kw) // d.stride + 1
            flops += 2.0 * d.N * ho * wo * d.Cout * d.Cin * d.kh * d.kw
        # one event pair around the back-to-back conv launches of a forward (launch gaps between the
        # kernels are part of the step, so they stay in the denominator)
        total_ms = 0.0
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for d in convs:
                _lib.check(self.lib.yv6_conv_fwd(self.handle, C.byref(d), sp))
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms / steps, flops, len(convs)

    def conv_bytes_per_launch(self, N, H, W):
        """Algorithmic HBM bytes of an average conv launch: each conv reads its input slice and writes its output
        slice once (bf16), weights once."""
        plan = self._plan(N, H, W, torch.float32)
        tot, n = 0.0, 0
        for kind, d in plan["calls"]:
            if kind != "conv":
                continue
            ho = (d.H + 2 * d.pad - d.kh) // d.stride + 1
            wo = (d.W + 2 * d.pad - d.kw) // d.stride + 1
            ysz = 4 if d.y_dtype == DT_F32 else 2
            tot += d.N * (d.H * d.W * d.Cin * 2 + ho * wo * d.Cout * ysz) + d.Cout * d.kh * d.kw * d.Cin * 2
            n += 1
        return tot / max(n, 1)

    def profile_layers(self, x, iters=10, stream=None):
        """Per-launch timing table (name, shape, ms, TFLOP/s) with an L2 flush before every launch."""
        x = x.contiguous()
        N, _, H, W = x.shape
        plan = self._plan(N, H, W, x.dtype if x.dtype == torch.uint8 else torch.float32)
        self.forward(x, stream)
        torch.cuda.synchronize()
        sp = _lib.stream_ptr(stream)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)
        names = []
        for op in self.g.ops:
            if op.kind in ("conv", "pred"):
                names.append(op.name)
            elif op.kind == "convT":
                names.extend([f"{op.name}[{q}]" for q in range(4)])
        rows = []
        convs = [d for kind, d in plan["calls"] if kind == "conv"]
        for name, d in zip(names, convs):
            ts = []
            for _ in range(iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(self.lib.yv6_conv_fwd(self.handle, C.byref(d), sp))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            ho = (d.H + 2 * d.pad - d.kh) // d.stride + 1
            wo = (d.W + 2 * d.pad - d.kw) // d.stride + 1
            fl = 2.0 * d.N * ho * wo * d.Cout * d.Cin * d.kh * d.kw
            by = 2.0 * d.N * (d.H * d.W * d.Cin + ho * wo * d.Cout)
            out = (C.c_int32 * 10)()
            _lib.check(self.lib.yv6_conv_plan(self.handle, C.byref(d), out))
            rows.append(dict(name=name, cin=d.Cin, cout=d.Cout, k=d.kh, s=d.stride, hw=f"{ho}x{wo}", ms=ms,
                             tflops=fl / ms / 1e9, gbs=by / ms / 1e6, plan=list(out)))
        return rows
