"""CPU checks of the training engine's table-driven repacks (train.py): the yv6_xform segment tables are executed by a
torch emulation of the kernel's index arithmetic (include/yv6.h: yv6_xform_seg) and compared with plain permutes of
the reference-layout parameters -- forward KRSC weights, rotated / parity-split dgrad weights, ConvTranspose
quadrants, padded prediction weights, and the way back from the fp32 KRSC / float64 gradient arenas to the flat
gradient buffer.  Also: the flat state keeps `state_dict()` / optimizer parameter groups of the reference intact."""
import ctypes as C

import numpy as np
import pytest
import torch

from yolov6_b200 import _lib
from yolov6_b200.flat import GROUP_B, GROUP_BNW, GROUP_EMA_ONLY, GROUP_W
from yolov6_b200.model import build_model
from yolov6_b200.synth import randomize_
from yolov6_b200.train import TrainEngine, op_branches


def dry_engine(model, n_buckets=1):
    """TrainEngine without a CUDA device: only the shape-independent state (flat buffers, arenas, tables)."""
    eng = TrainEngine.__new__(TrainEngine)
    eng.model, eng.g = model, model.graph
    eng.dev = torch.device("cpu")
    eng.n_buckets = n_buckets
    eng.debug, eng.dbg, eng._shape, eng.bucket_hook = False, {}, None, None
    eng._build_state()
    return eng


class Memory:
    """Maps raw addresses back to torch tensors (the emulator's 'device memory')."""

    def __init__(self):
        self.blocks = []

    def add(self, t):
        flat = t.reshape(-1) if t.is_contiguous() else None
        assert flat is not None
        self.blocks.append((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), flat))

    def find(self, ptr):
        for lo, hi, t in self.blocks:
            if lo <= ptr < hi:
                return t, (ptr - lo)
        raise KeyError(hex(ptr))


def run_table(table, mem, arena_bytes, accumulate=False):
    raw = table.segs.numpy().tobytes()
    segs = (_lib.XformSeg * max(table.n, 1)).from_buffer_copy(raw)
    for s in segs[:table.n]:
        n, ds, ss = list(s.n), list(s.ds), list(s.ss)
        idx = np.indices(n).reshape(4, -1)
        so = sum(idx[i].astype(np.int64) * ss[i] for i in range(4))
        do = sum(idx[i].astype(np.int64) * ds[i] for i in range(4))
        src_t, src_b = mem.find(s.src)
        if s.src_dtype == _lib.XF_F64:
            view = src_t.view(torch.float64) if src_t.dtype == torch.uint8 else src_t
            vals = view[src_b // 8 + torch.from_numpy(so)].float()
        else:
            view = src_t.view(torch.float32) if src_t.dtype == torch.uint8 else src_t
            vals = view[src_b // 4 + torch.from_numpy(so)]
        dst_t, dst_b = mem.find(s.dst)
        es = 2 if s.dst_dtype == _lib.XF_BF16 else 4
        assert dst_t.element_size() == es
        pos = dst_b // es + torch.from_numpy(do)
        assert len(torch.unique(pos)) == len(pos), "a segment writes an element twice"
        if accumulate and es == 4:
            dst_t[pos] += vals
        else:
            dst_t[pos] = vals.to(dst_t.dtype)


@pytest.fixture(scope="module", params=["yolov6n", "yolov6m"])
def engine(request):
    torch.manual_seed(0)
    m = randomize_(build_model(request.param, 80, torch.device("cpu")), seed=1)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    eng = dry_engine(m, n_buckets=3)
    return m, eng, before


def test_flat_state_keeps_the_reference_surface(engine):
    m, eng, before = engine
    after = m.state_dict()
    assert list(after) == list(before)
    for k in before:
        assert torch.equal(after[k], before[k]), k
    fl = eng.flat
    # every trainable parameter is a view of the flat buffer, gradients precede everything else
    for n, p in m.named_parameters():
        o, k, shape = fl.slots[n]
        assert p.data_ptr() == fl.pflat.data_ptr() + 4 * o and tuple(p.shape) == shape
        assert (o + k <= fl.n_train) == p.requires_grad
    # optimizer groups of the reference's build_optimizer (solver/build.py:12-19)
    groups = {0: 0, 1: 0, 2: 0}
    import torch.nn as nn
    g_bnw, g_w, g_b = [], [], []
    for v in m.modules():
        if hasattr(v, 'bias') and isinstance(v.bias, nn.Parameter):
            g_b.append(v.bias)
        if isinstance(v, nn.BatchNorm2d):
            g_bnw.append(v.weight)
        elif hasattr(v, 'weight') and isinstance(v.weight, nn.Parameter):
            g_w.append(v.weight)
    grp = fl.group.cpu()
    for want, plist in ((GROUP_BNW, g_bnw), (GROUP_W, g_w), (GROUP_B, g_b)):
        for p in plist:
            if not p.requires_grad:
                continue
            o = (p.data_ptr() - fl.pflat.data_ptr()) // 4
            assert int(grp[o // 4]) == want
            groups[want] += 1
    assert groups[0] > 50 and groups[1] > 50 and groups[2] > 50
    o = fl.slots["detect.proj"][0]
    assert int(grp[o // 4]) == GROUP_EMA_ONLY
    # load_state_dict writes through the views
    sd = {k: (v + 1 if v.dtype.is_floating_point else v) for k, v in before.items()}
    m.load_state_dict(sd)
    o, k, _ = fl.slots["backbone.stem.rbr_dense.conv.weight" if "backbone.stem.rbr_dense.conv.weight" in fl.slots else "backbone.stem.block.conv.weight"]
    assert torch.allclose(fl.pflat[o:o + k], sd[[n for n in sd if n.startswith("backbone.stem") and n.endswith("conv.weight")][0]].reshape(-1))
    m.load_state_dict(before)


def test_weight_pack_table_matches_plain_permutes(engine):
    m, eng, _ = engine
    mem = Memory()
    mem.add(eng.flat.pflat)
    for W in eng.wts.values():
        for key in ("w", "wt", "bias"):
            if key in W:
                mem.add(W[key])
        for ent in W.get("br", []):
            if ent.get("w") is not None:
                mem.add(ent["w"])
            for t in ent.get("wt", []) or []:
                mem.add(t)
    run_table(eng.pack_table, mem, None)
    P = dict(m.named_parameters())
    bf = lambda t: t.to(torch.bfloat16)   # noqa: E731
    checked = 0
    for i, op in enumerate(m.graph.ops):
        if op.kind == "pool":
            continue
        W = eng.wts[i]
        if op.kind == "pred":
            w = P[op.name + ".weight"].detach()
            assert torch.equal(W["w"], bf(w.permute(0, 2, 3, 1)))
            assert torch.equal(W["bias"][:op.cout], P[op.name + ".bias"].detach()) and float(W["bias"][op.cout:].abs().sum()) == 0
            chp = W["wt"].shape[3]
            want = torch.zeros(op.cin, 1, 1, chp)
            want[:, 0, 0, :op.cout] = w[:, :, 0, 0].t()
            assert torch.equal(W["wt"], bf(want))
        elif op.kind == "convT":
            wt = P[op.name + ".upsample_transpose.weight"].detach()        # [Cin, Cout, 2, 2]
            for q in range(4):
                dy, dx = q // 2, q % 2
                assert torch.equal(W["w"][q], bf(wt[:, :, dy, dx].t().reshape(op.cout, 1, 1, op.cin)))
                assert torch.equal(W["wt"][q], bf(wt[:, :, dy, dx].reshape(op.cin, 1, 1, op.cout)))
            assert torch.equal(W["bias"][:op.cout], P[op.name + ".upsample_transpose.bias"].detach())
        else:
            for ent in W["br"]:
                k = ent["k"]
                if k == 0:
                    continue
                w = P[ent["prefix"] + ".conv.weight"].detach()
                if op.kind == "stem":
                    w33 = w if k == 3 else torch.nn.functional.pad(w, [1, 1, 1, 1])
                    assert torch.equal(ent["w"], w33.permute(2, 3, 1, 0).contiguous())
                    continue
                krsc = w.permute(0, 2, 3, 1)
                assert torch.equal(ent["w"], bf(krsc))
                if op.s == 1:
                    assert torch.equal(ent["wt"][0], bf(krsc.flip(1, 2).permute(3, 1, 2, 0)))
                elif k == 1:
                    assert torch.equal(ent["wt"][0], bf(krsc.permute(3, 1, 2, 0)))
                else:
                    j = 0
                    for ph in range(2):
                        for pw in range(2):
                            rows = [1] if ph == 0 else [2, 0]
                            cols = [1] if pw == 0 else [2, 0]
                            assert torch.equal(ent["wt"][j], bf(krsc[:, rows][:, :, cols].permute(3, 1, 2, 0)))
                            j += 1
                checked += 1
    assert checked > 40


def test_gradient_unpack_table_and_buckets(engine):
    m, eng, _ = engine
    fl = eng.flat
    g = torch.Generator().manual_seed(3)
    # fill the arena with recognisable data: fp32 where weight gradients live, float64 where sums live
    arena = eng.zero_arena
    arena.view(torch.float32)[:] = torch.randn(arena.numel() // 4, generator=g)
    mem = Memory()
    mem.add(arena)
    mem.add(fl.gflat)
    base = arena.data_ptr()
    ops = m.graph.ops
    dbl = arena.view(torch.float64)
    for i, op in enumerate(ops):     # float64 slots: write clean doubles
        if op.kind in ("pool",):
            continue
        keys = ["bsum"] if op.kind in ("pred", "convT") else ["s1", "s2", "dalpha"]
        for key in keys:
            o = eng._z(i, key) - base
            n = {"bsum": 2 * ((op.cout + 15) // 16 * 16 if op.kind == "pred" else op.cout), "s1": op.cout,
                 "s2": len(op_branches(op)) * op.cout if op.kind not in ("pred", "convT") else 0, "dalpha": 1}[key]
            dbl[o // 8:o // 8 + n] = torch.randn(n, generator=g, dtype=torch.float64)
    fl.gflat.zero_()
    for t in eng.grad_tables:
        run_table(t, mem, None)
    f32 = arena.view(torch.float32)
    seen = 0
    for i, op in enumerate(ops):
        if op.kind == "pool":
            continue
        if op.kind == "pred":
            dw = f32[(eng._z(i, "dw") - base) // 4:][:op.cout * op.cin].view(op.cout, op.cin, 1, 1)
            assert torch.equal(fl.grad_view(op.name + ".weight"), dw)
            bs = dbl[(eng._z(i, "bsum") - base) // 8:][:op.cout].float()
            assert torch.equal(fl.grad_view(op.name + ".bias"), bs)
        elif op.kind == "convT":
            dw = f32[(eng._z(i, "dw") - base) // 4:][:4 * op.cout * op.cin].view(2, 2, op.cout, op.cin)
            assert torch.equal(fl.grad_view(op.name + ".upsample_transpose.weight"), dw.permute(3, 2, 0, 1))
        else:
            s1 = dbl[(eng._z(i, "s1") - base) // 8:][:op.cout].float()
            for b, (prefix, k) in enumerate(op_branches(op)):
                bn = prefix + (".bn" if k else "")
                s2 = dbl[(eng._z(i, "s2") - base) // 8 + b * op.cout:][:op.cout].float()
                assert torch.equal(fl.grad_view(bn + ".weight"), s2) and torch.equal(fl.grad_view(bn + ".bias"), s1)
                if k == 0:
                    continue
                got = fl.grad_view(prefix + ".conv.weight")
                if op.kind == "stem":
                    dw = f32[(eng._z(i, "dw", b) - base) // 4:][:op.cout * 32].view(op.cout, 32)       # [co][(r*3+s)*3 + c | pad]
                    if k == 3:
                        assert torch.equal(got, dw[:, :27].reshape(op.cout, 3, 3, 3).permute(0, 3, 1, 2))
                    else:
                        assert torch.equal(got, dw[:, 12:15].reshape(op.cout, 3, 1, 1))
                else:
                    dw = f32[(eng._z(i, "dw", b) - base) // 4:][:op.cout * k * k * op.cin].view(op.cout, k, k, op.cin)
                    assert torch.equal(got, dw.permute(0, 3, 1, 2))
                seen += 1
            if op.alpha:
                da = dbl[(eng._z(i, "dalpha") - base) // 8].float()
                assert float(fl.grad_view(op.alpha)) == float(da)
    assert seen > 40
    # buckets: contiguous, ordered, exact cover of the gradient buffer; each table writes only inside its own bucket
    rng = eng.bucket_range
    assert rng[0][0] == 0 and rng[-1][1] == fl.n_train and all(a[1] == b[0] for a, b in zip(rng, rng[1:]))
    assert sum(hi > lo for lo, hi in rng) >= 2
    for k, t in enumerate(eng.grad_tables):
        fl.gflat.fill_(float("nan"))
        run_table(t, mem, None)
        written = ~torch.isnan(fl.gflat)
        lo, hi = rng[k]
        assert not written[:lo].any() and not written[hi:].any() and written[lo:hi].any()
    # accumulate mode adds
    fl.gflat.zero_()
    for t in eng.grad_tables:
        run_table(t, mem, None)
    once = fl.gflat.clone()
    for t in eng.grad_tables:
        run_table(t, mem, None, accumulate=True)
    assert torch.allclose(fl.gflat, 2 * once)
