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
import os

import numpy as np
import torch

from . import _lib, ops
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
        self.fuse_siblings = True
        self.n_lanes = int(os.environ.get("YV6_LANES", "4"))   # streams the launches of one forward are spread over (1 = serial)
        self.force_pair = -1 if os.environ.get("YV6_PAIR", "1") == "0" else 0   # A/B switch: single-CTA kernels only
        self._side = None
        self.weights = {}     # op index -> dict(w=..., bias=..., alpha=...)
        self._plans = {}      # (N, H, W, dtype) -> plan, least recently used first; bounded (rect-shaped evaluation
        self.max_plans = 4    # would otherwise keep one full activation set per distinct shape)
        self._pinned = set()  # shapes captured into CUDA graphs (their buffers must outlive the graph)
        self._pack(state_dict)

    # ------------------------------------------------------------------ weight packing
    def _siblings(self):
        """Pairs of convs that read the same tensor with the same geometry and can run as ONE launch over the
        concatenated output channels: the cls / reg branches of the decoupled head (effidehead.py:79-84)."""
        pairs = {}
        ops = self.g.ops
        for i, a in enumerate(ops):
            if a.kind != "conv" or not a.name.startswith("detect.cls_convs."):
                continue
            for j in range(i + 1, min(i + 3, len(ops))):
                b = ops[j]
                if (b.kind == "conv" and b.name == a.name.replace("cls_convs", "reg_convs") and b.src == a.src and
                        (b.k, b.s, b.act, b.cout, b.cin) == (a.k, a.s, a.act, a.cout, a.cin) and a.res is None and b.res is None and
                        a.dst.c_off == 0 and b.dst.c_off == 0 and a.cout % 64 == 0):
                    pairs[i] = j
        return pairs

    def _pack(self, sd):
        sd = {k: v.detach().cpu() for k, v in sd.items()}
        dev = self.device
        self.sibling = self._siblings() if self.fuse_siblings else {}
        self.sibling_second = {j: i for i, j in self.sibling.items()}
        for i, op in enumerate(self.g.ops):
            if op.kind == "pool" or (op.kind == "pred" and op.head[0] not in ("cls", "reg")):   # fuse_ab / distillation preds: training only
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
                if op.kind == "conv" and op.k == 3 and op.s == 2 and (op.cin <= 32 or op.cin in (64, 128)):     # (64 | 128: the view's zero blocks are skipped)
                    # column-pair view (see _plan): [Cout][3][2][2*Cin]; tap 0 = input columns (2j-2 | 2j-1), tap 1 = (2j | 2j+1)
                    wf = ops.pair_view_weights(ws[0].float()).to(dev)
                    ent["w_pair"] = _split3(wf).contiguous() if self.nsplit == 3 else wf.to(torch.bfloat16).contiguous()
                bias = torch.zeros((op.cout + 255) // 256 * 256, dtype=torch.float32, device=dev)
                bias[:op.cout] = b.float().to(dev)
                ent["bias"] = bias
            ent["alpha"] = float(sd[op.alpha]) if (op.alpha and op.res is not None) else 1.0
            self.weights[i] = ent
        for i, j in self.sibling.items():      # one conv with 2 x Cout output channels: [cls branch | reg branch]
            (wa, ba), (wb, bb) = fold_op(sd, self.g.ops[i]), fold_op(sd, self.g.ops[j])
            wf = torch.cat([wa, wb], 0).float().to(dev)
            c2 = wf.shape[0]
            bias = torch.zeros((c2 + 255) // 256 * 256, dtype=torch.float32, device=dev)
            bias[:c2] = torch.cat([ba, bb]).float().to(dev)
            self.weights[i]["w_fused"] = _split3(wf).contiguous() if self.nsplit == 3 else wf.to(torch.bfloat16).contiguous()
            self.weights[i]["bias_fused"] = bias

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
        plan = {"bufs": [], "calls": [], "conv_info": [], "deps": [], "pred_descs": [], "head_set": 0}
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

        # sibling convs write one [.., 2C] buffer; their consumers read its two channel halves
        redirect = {}       # graph buffer index -> (fused tensor, channel offset, channel pitch)
        plan["fused_bufs"] = []
        for i, j in self.sibling.items():
            a, b2 = g.ops[i], g.ops[j]
            lvl = g.bufs[a.dst.buf].level
            h, w = H >> lvl, W >> lvl
            shape = (P, N, h, w, 2 * a.cout) if P == 3 else (N, h, w, 2 * a.cout)
            fb = torch.zeros(shape, dtype=torch.bfloat16, device=dev)
            plan["fused_bufs"].append(fb)
            redirect[a.dst.buf] = (fb, 0, 2 * a.cout)
            redirect[b2.dst.buf] = (fb, a.cout, 2 * a.cout)

        def view(t):
            b = g.bufs[t.buf]
            h, w = H >> b.level, W >> b.level
            if t.buf in redirect:
                fb, coff, pitch = redirect[t.buf]
                return fb, h, w, pitch
            return plan["bufs"][t.buf], h, w, b.c_total

        def coff(t):
            return t.c_off + (redirect[t.buf][1] if t.buf in redirect else 0)

        def span(t):
            """(tensor identity, first channel, end channel) of a graph tensor slice: the unit of the dependency analysis."""
            return (id(view(t)[0]), coff(t), coff(t) + t.c)

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
                plan["deps"].append(dict(reads=[], writes=[span(op.dst)]))
            elif op.kind in ("conv", "pred", "convT"):
                if op.kind == "pred" and op.head[0] not in ("cls", "reg"):
                    continue                      # eval forward: anchor-free (cls, reg) branches only (effidehead_fuseab.py:141-199, _distill_ns.py:105-150)
                if i in self.sibling_second:
                    continue                      # computed by its sibling's launch
                sbuf, sh, sw, sct = view(op.src)
                quads = ent["w"] if op.kind == "convT" else [ent["w"][0]]
                fused = i in self.sibling
                for q, wq in enumerate(quads):
                    d = ConvDesc()
                    d.x = sbuf.data_ptr() + coff(op.src) * 2
                    d.N, d.H, d.W, d.Cin, d.x_c_total = N, sh, sw, op.cin, sct
                    d.x_plane_stride = sbuf.stride(0) if P == 3 else 0
                    if fused:
                        wq = ent["w_fused"]
                    d.w = wq.data_ptr()
                    d.w_plane_stride = wq.stride(0) if P == 3 else 0
                    d.bias = (ent["bias_fused"] if fused else ent["bias"]).data_ptr()
                    d.Cout = op.cout * (2 if fused else 1)
                    d.kh = d.kw = 1 if op.kind == "convT" else op.k
                    d.stride = 1 if op.kind == "convT" else op.s
                    d.pad = d.kh // 2
                    d.pad_w = _lib.PAD_SAME
                    d.act = ACT_CODES[op.act]
                    d.nsplit = P
                    d.force_pair = self.force_pair
                    oh, ow = (sh + 2 * d.pad - d.kh) // d.stride + 1, (sw + 2 * d.pad - d.kw) // d.stride + 1
                    plan["conv_info"].append(dict(name=op.name if op.kind != "convT" else f"{op.name}[{q}]", cin=op.cin, cout=int(d.Cout),
                                                  k=d.kh, s=d.stride, ho=oh, wo=ow, h=sh, w=sw,
                                                  flops=2.0 * N * oh * ow * d.Cout * op.cin * d.kh * d.kw,
                                                  y_f32=op.kind == "pred"))
                    if op.kind == "pred":
                        which, lvl = op.head
                        out = plan[which]
                        ch = out.shape[2]
                        lh, lw = sizes[lvl]
                        d.y = out.data_ptr() + int(offs[lvl]) * ch * 4
                        plan["pred_descs"].append((d, which, int(offs[lvl]) * ch * 4))
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
                            d.y = dbuf.data_ptr() + (coff(op.dst) if not fused else 0) * 2
                            d.y_img_stride, d.y_h_stride, d.y_w_stride = dh * dw * dct, dw * dct, dct
                    if op.res is not None:
                        rbuf, rh, rw, rct = view(op.res)
                        d.res = rbuf.data_ptr() + coff(op.res) * 2
                        d.res_img_stride, d.res_h_stride, d.res_w_stride = rh * rw * rct, rw * rct, rct
                        d.res_plane_stride = rbuf.stride(0) if P == 3 else 0
                        d.alpha = ent["alpha"]
                    if "w_pair" in ent and op.src.c_off == 0 and sct == op.cin and sw % 2 == 0:
                        # 3x3 stride-2 conv over <= 128 channels.  On the column-pair view of the same memory, [N, H, W/2, 2*Cin],
                        # it is a 3x2 conv with stride (2, 1) over contiguous 2*Cin-channel rows.  <= 32 channels: 64-byte pixels
                        # fetched with an element stride of 2 keep the TMA unit, not the tensor pipe, busy (ERBlock_2.0 of
                        # YOLOv6-S: 170 -> 135 us).  And where the library's halo-reuse mainloop takes the view (one 9 x 33 input
                        # box per 8 x 16 output tile instead of one box per tap, include/yv6.h `pair_view`) the input operand
                        # crosses L2 -> shared memory 2.6x less often, which is what bounds these small-K layers.
                        keep = (d.W, d.Cin, d.x_c_total, d.w, d.w_plane_stride, d.kw, d.stride_w, d.pad_w, d.out_w)
                        d.W, d.Cin, d.x_c_total = sw // 2, 2 * op.cin, 2 * op.cin
                        d.w = ent["w_pair"].data_ptr()
                        d.w_plane_stride = ent["w_pair"].stride(0) if P == 3 else 0
                        d.kw, d.stride_w, d.pad_w, d.out_w = 2, 1, 1, sw // 2
                        d.pair_view = 1
                        out10 = (C.c_int32 * 10)()
                        _lib.check(self.lib.yv6_conv_plan(self.handle, C.byref(d), out10))
                        if op.cin > 32 and out10[8] != 2:       # generic mainloop: the plain stride-2 conv moves fewer bytes
                            d.W, d.Cin, d.x_c_total, d.w, d.w_plane_stride, d.kw, d.stride_w, d.pad_w, d.out_w = keep
                            d.pair_view = 0
                    if "neck_start" not in plan and op.name.startswith("neck."):
                        plan["neck_start"] = len(plan["calls"])     # first launch after the backbone (pipeline.DetectStream forks here)
                    plan["calls"].append(("conv", d))
                    reads = [span(op.src)] + ([span(op.res)] if op.res is not None else [])
                    if op.kind == "pred":
                        writes = [(("head",) + tuple(op.head), 0, 1)]
                    elif fused:
                        writes = [(id(dbuf), 0, 2 * op.cout)]
                    else:
                        writes = [span(op.dst)]
                    plan["deps"].append(dict(reads=reads, writes=writes))
            elif op.kind == "pool":
                buf, h, w, ct = view(op.dst)
                plan["calls"].append(("pool", (buf.data_ptr(), N, h, w, op.cin, ct, P, buf.stride(0) if P == 3 else 0)))
                plan["deps"].append(dict(reads=[(id(buf), 0, op.cin)], writes=[(id(buf), op.cin, 4 * op.cin)]))
        self._schedule(plan)
        self._plans[key] = plan
        return plan

    def _schedule(self, plan):
        """Assigns every launch to one of a few streams from its data dependencies (channel slices of the activation buffers):
        a chain keeps its stream, an op whose producers' streams have moved on takes the least recently used one and waits on
        events.  The backbone stays one chain; the branches of BiFusion (transpose-conv quadrants, cv1, cv2 + downsample), of
        CSPSPPF and the three head levels run side by side -- these launches are tens of CTAs and a few microseconds each, and
        their launch-to-first-MMA latency and drain overlap instead of adding up.  Inside a captured CUDA graph the events
        become plain graph edges."""
        K = self.n_lanes
        deps = plan["deps"]
        n = len(deps)
        writers = {}
        lane, waits = [0] * n, [[] for _ in range(n)]
        last_on = [-1] * K
        for i, dp in enumerate(deps):
            prod = set()
            for key, lo, hi in dp["reads"]:
                for wlo, whi, j in writers.get(key, ()):
                    if wlo < hi and lo < whi:
                        prod.add(j)
            chain = [lane[j] for j in prod if last_on[lane[j]] == j]
            if chain:
                s = min(chain)
            elif not prod:
                s = 0
            else:
                s = min(range(K), key=lambda t: last_on[t])
            lane[i] = s
            waits[i] = sorted(j for j in prod if lane[j] != s)
            last_on[s] = i
            for key, lo, hi in dp["writes"]:
                writers.setdefault(key, []).append((lo, hi, i))
        need = set(j for w in waits for j in w)
        tails = [last_on[t] for t in range(1, K) if last_on[t] >= 0]
        need.update(tails)
        plan["lane"], plan["waits"], plan["tails"] = lane, waits, tails
        plan["events"] = {j: torch.cuda.Event() for j in need}
        plan["fork"] = torch.cuda.Event()

    def pin(self, N, H, W, in_dtype=torch.float32):
        """Keep the buffers of this shape for the engine's lifetime (they are referenced by a captured CUDA graph)."""
        self._plan(N, H, W, in_dtype)
        self._pinned.add((N, H, W, in_dtype))

    # ------------------------------------------------------------------ execution
    def launch_count(self, N, H, W, in_dtype=torch.float32):
        """Kernels launched per forward for this shape (convs + stem + pools + decode)."""
        return len(self._plan(N, H, W, in_dtype)["calls"]) + 1

    def _select_head_set(self, plan, k):
        """Point the pred convs at head-output set k (0 = the default tensors; 1 = a second pair, allocated on first use).
        The software-pipelined serving loop (pipeline.DetectStream) alternates between the two so that the NMS of batch
        i - 1 can read one set while the network of batch i writes the other."""
        if k == plan["head_set"]:
            return
        if k == 1 and "cls_alt" not in plan:
            plan["cls_alt"], plan["reg_alt"] = torch.empty_like(plan["cls"]), torch.empty_like(plan["reg"])
        for d, which, off in plan["pred_descs"]:
            d.y = plan[which + ("_alt" if k == 1 else "")].data_ptr() + off
        plan["head_set"] = k

    def forward(self, x, stream=None, decode=True, head_set=0, hook=None):
        """x: [N,3,H,W] CUDA tensor, fp32 in [0,1] or uint8.  Returns pred [N,A,5+nc] fp32 (a buffer owned
        by the engine, overwritten by the next call with the same shape).  decode=False stops after the head convs and
        returns (cls [N,A,nc], reg [N,A,R], level sizes): the serving pipeline feeds them to the NMS kernels directly;
        head_set = 1 makes the head convs write the alternate (cls, reg) pair (decode=False only).  hook = (i, fn): fn() is
        called just before launch number i is enqueued, with the main stream current (used to fork a side branch there)."""
        if x.device != self.device:
            raise RuntimeError(f"input on {x.device}, engine on {self.device}")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        x = x.contiguous()
        N, Cin, H, W = x.shape
        assert Cin == 3
        plan = self._plan(N, H, W, x.dtype)
        plan["image"] = x  # keep alive while kernels are in flight
        assert head_set == 0 or not decode, "the alternate head set is for decode=False callers"
        self._select_head_set(plan, head_set)
        main = stream if stream is not None else torch.cuda.current_stream(self.device)
        sp = _lib.stream_ptr(main)
        lib, h, chk = self.lib, self.handle, _lib.check
        plan["stem"].x = x.data_ptr()
        lane, waits, events = plan["lane"], plan["waits"], plan["events"]
        multi = self.n_lanes > 1 and max(lane) > 0
        if multi:
            if self._side is None:
                self._side = [torch.cuda.Stream(device=self.device) for _ in range(self.n_lanes - 1)]
            streams = [main] + self._side
            sps = [_lib.stream_ptr(t) for t in streams]
            plan["fork"].record(main)
            forked = set()
        for i, (kind, d) in enumerate(plan["calls"]):
            if hook is not None and i == hook[0]:
                hook[1]()
            if multi:
                t = lane[i]
                if t and t not in forked:        # a side stream starts after everything the caller queued before this forward
                    streams[t].wait_event(plan["fork"])
                    forked.add(t)
                for j in waits[i]:
                    streams[t].wait_event(events[j])
                sp = sps[t]
            if kind == "conv":
                chk(lib.yv6_conv_fwd(h, C.byref(d), sp))
            elif kind == "stem":
                chk(lib.yv6_stem_fwd(h, C.byref(d), sp))
            else:
                chk(lib.yv6_sppf_pool(h, C.c_void_p(d[0]), d[1], d[2], d[3], d[4], d[5], d[6], d[7], sp))
            if multi and i in events:
                events[i].record(streams[lane[i]])
        if multi:                                  # join: the caller's stream continues after every lane
            for j in plan["tails"]:
                main.wait_event(events[j])
            sp = sps[0]
        g = self.g
        if not decode:
            if head_set == 1:
                return plan["cls_alt"], plan["reg_alt"], plan["sizes"]
            return plan["cls"], plan["reg"], plan["sizes"]
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
        flops = sum(ci["flops"] for ci in plan["conv_info"])     # algorithmic (the column-pair view's zero taps do not count)
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
        for ci in plan["conv_info"]:
            ysz = 4 if ci["y_f32"] else 2
            tot += N * (ci["h"] * ci["w"] * ci["cin"] * 2 + ci["ho"] * ci["wo"] * ci["cout"] * ysz) + ci["cout"] * ci["k"] ** 2 * ci["cin"] * 2
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
        rows = []
        convs = [d for kind, d in plan["calls"] if kind == "conv"]
        for ci, d in zip(plan["conv_info"], convs):
            name = ci["name"]
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
            ho, wo = ci["ho"], ci["wo"]
            fl = ci["flops"]
            by = 2.0 * d.N * (ci["h"] * ci["w"] * ci["cin"] + ho * wo * ci["cout"])
            out = (C.c_int32 * 10)()
            _lib.check(self.lib.yv6_conv_plan(self.handle, C.byref(d), out))
            rows.append(dict(name=name, cin=ci["cin"], cout=ci["cout"], k=ci["k"], s=ci["s"], hw=f"{ho}x{wo}", ms=ms,
                             tflops=fl / ms / 1e9, gbs=by / ms / 1e6, plan=list(out)))
        return rows
