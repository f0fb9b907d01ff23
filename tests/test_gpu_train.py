"""GPU parity of the training engine (train-form forward + backward through conv / wgrad / BN kernels)
against the oracle's train-mode network differentiated by torch autograd on CPU (float64).

The kernels compute in bf16 (operands and stored activations) with fp32 accumulation, like the
reference's own GPU training which runs convs under fp16 autocast (core/engine.py:150).  Bars, per
parameter tensor: cosine similarity of the gradient >= 0.99 and relative L2 error <= 8e-2 against the
float64 oracle; forward head outputs within 3e-2.  Measured values are printed."""
import numpy as np
import pytest
import torch

from conftest import golden_keys
from oracle import fabricate as fab
from oracle import model as om

pytestmark = pytest.mark.gpu


def oracle_grads(name, sd, x, wc, wr):
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    with om.train_mode():
        cls, reg, _ = om.forward(sd64, om.CONFIGS[name], x.double(), train_outputs=True)
    loss = (cls * wc.double()).sum() + (reg * wr.double()).sum()
    loss.backward()
    return cls.detach(), reg.detach(), {k: v.grad for k, v in sd64.items() if v.is_floating_point() and v.grad is not None}


@pytest.mark.parametrize("name,size,batch", [("yolov6n", 128, 4), ("yolov6s", 96, 2)])
def test_train_step_gradients_match_oracle(name, size, batch):
    from yolov6_b200.model import build_model
    dev = torch.device("cuda:0")
    sd = fab.fabricate_state_dict(golden_keys(name), seed=0)
    m = build_model(name, 80, dev)
    m.load_state_dict(sd)
    m.train()
    x = fab.synthetic_images(batch, size, size, seed=11)
    (feats, cls, reg), _ = m(x.to(dev))
    g = torch.Generator().manual_seed(5)
    wc, wr = torch.randn(cls.shape, generator=g), torch.randn(reg.shape, generator=g)
    loss = (cls * wc.to(dev)).sum() + (reg * wr.to(dev)).sum()
    loss.backward()
    ocls, oreg, ograds = oracle_grads(name, sd, x, wc, wr)
    e_cls = float((cls.detach().cpu().double() - ocls).abs().max())
    e_reg = float((reg.detach().cpu().double() - oreg).abs().max() / (1 + oreg.abs().max()))
    print(f"{name}: forward |dcls| {e_cls:.2e}  rel |dreg| {e_reg:.2e}")
    assert e_cls < 3e-2 and e_reg < 3e-2
    assert [tuple(f.shape[2:]) for f in feats] == [(size // s, size // s) for s in om.CONFIGS[name]["strides"]]
    worst_cos, worst_rel, nchecked = 1.0, 0.0, 0
    for k, p in m.named_parameters():
        if k not in ograds:
            continue
        assert p.grad is not None, f"no gradient for {k}"
        a, b = p.grad.detach().cpu().double().flatten(), ograds[k].flatten()
        if b.norm() < 1e-12:
            continue
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))
        rel = float((a - b).norm() / b.norm())
        if cos < worst_cos or rel > worst_rel:
            print(f"  {k}: cos {cos:.4f} rel {rel:.3e}")
        worst_cos, worst_rel, nchecked = min(worst_cos, cos), max(worst_rel, rel), nchecked + 1
        assert cos >= 0.99 and rel <= 8e-2, f"{k}: cos {cos:.4f} rel {rel:.3e}"
    print(f"{name}: {nchecked} parameter gradients checked, worst cos {worst_cos:.4f}, worst rel {worst_rel:.3e}")
    assert nchecked > 300
    # running statistics follow nn.BatchNorm2d (momentum 0.03, unbiased variance)
    rm = dict(m.named_buffers())["backbone.ERBlock_2.0.rbr_dense.bn.running_mean"].cpu()
    assert not torch.allclose(rm, sd["backbone.ERBlock_2.0.rbr_dense.bn.running_mean"])


def test_wgrad_kernel_matches_torch():
    import ctypes as C
    from yolov6_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    for (N, H, W, Cin, Cout, k, s) in [(2, 20, 20, 64, 64, 3, 1), (3, 16, 24, 128, 96, 1, 1), (2, 32, 32, 32, 64, 3, 2),
                                       (2, 16, 16, 48, 16, 3, 1), (1, 40, 40, 256, 512, 3, 1), (2, 16, 16, 320, 80, 1, 1)]:
        x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
        Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
        dy = torch.randn(N, Ho, Wo, Cout, generator=g).to(torch.bfloat16)
        xs = x.float().permute(0, 3, 1, 2).double().requires_grad_(False)
        w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
        y = torch.nn.functional.conv2d(xs, w, stride=s, padding=k // 2)
        (y * dy.float().permute(0, 3, 1, 2).double()).sum().backward()
        ref = w.grad.permute(0, 2, 3, 1)
        dw = torch.zeros(Cout, k, k, Cin, dtype=torch.float32, device=dev)
        d = _lib.WgradDesc()
        xd, dyd = x.to(dev), dy.to(dev)
        d.x, d.N, d.H, d.W, d.Cin, d.x_c_total = xd.data_ptr(), N, H, W, Cin, Cin
        d.dy, d.Cout, d.dy_c_total = dyd.data_ptr(), Cout, Cout
        d.kh = d.kw = k
        d.stride, d.pad, d.dw = s, k // 2, dw.data_ptr()
        _lib.check(_lib.lib().yv6_conv_wgrad(_lib.handle(0), C.byref(d), _lib.stream_ptr()))
        err = float((dw.cpu().double() - ref).abs().max() / (ref.abs().max() + 1e-9))
        assert err < 1e-4, f"wgrad {(N, H, W, Cin, Cout, k, s)}: rel err {err:.3e}"
