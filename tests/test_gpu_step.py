"""GPU tests of the step-level machinery: the CUDA-graph training step (step.py) against the autograd path of the same
engine, the fused SGD + EMA kernel (optim.py) against torch.optim.SGD + the reference's ModelEMA arithmetic, gradient
accumulation, and -- with two GPUs -- the gradient all-reduce against torch's DistributedDataParallel
(core/engine.py:456-468, 171-172)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import golden_keys
from oracle import fabricate as fab
from oracle import loss as oloss

pytestmark = pytest.mark.gpu


def make_model(name="yolov6n", seed=0):
    from yolov6_b200.model import build_model
    sd = fab.fabricate_state_dict(golden_keys(name), seed=seed)
    for k in sd:
        if (".cls_preds." in k or ".reg_preds." in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    m = build_model(name, 80, torch.device("cuda:0"))
    m.load_state_dict(sd)
    return m.train()


def make_loss(size, name="yolov6n"):
    from yolov6_b200.loss import ComputeLoss
    kw = dict(use_dfl=False, reg_max=0, iou_type="siou") if name in ("yolov6n", "yolov6s") else dict(use_dfl=True, reg_max=16, iou_type="giou")
    return ComputeLoss(num_classes=80, ori_img_size=size, warmup_epoch=0, **kw)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("name,graph", [("yolov6n", True), ("yolov6n", False), ("yolov6m", True)])
def test_train_step_matches_autograd_path(name, graph):
    from yolov6_b200.step import TrainStep
    B, S = 2, 128
    x = fab.synthetic_images(B, S, S, seed=3).cuda()
    targets = oloss.synthetic_targets(B, seed=4).cuda()
    m1, m2 = make_model(name), make_model(name)
    c1, c2 = make_loss(S, name), make_loss(S, name)
    preds, _ = m1(x)
    loss, items = c1(preds, targets, 1, 0, S, S)
    loss.backward()
    ref = {n: p.grad.clone() for n, p in m1.named_parameters() if p.grad is not None}
    step = TrainStep(m2, c2, B, S, S, in_dtype=torch.float32, max_gt=64, graph=graph)
    step.load(x, targets)
    out = step.run(epoch_num=1).clone()
    torch.cuda.synchronize()
    assert not step.overflowed()
    assert abs(float(out[0]) - float(loss)) <= 1e-4 * abs(float(loss)), (float(out[0]), float(loss))
    np.testing.assert_allclose(out[1:4].cpu().numpy(), items.cpu().numpy(), rtol=1e-4, atol=1e-7)
    fl = step.eng.flat
    worst = 0.0
    for n, g in ref.items():
        if float(g.norm()) < 1e-12:
            continue
        e = rel(fl.grad_view(n), g)
        worst = max(worst, e)
        assert e < 5e-3, f"{n}: {e:.3e}"      # same kernels; fp32 atomics make the accumulation order differ
    first = fl.gflat.clone()
    step.run(epoch_num=1)                        # a second run of the same batch reproduces the gradients (arena cleared)
    assert rel(fl.gflat, first) < 5e-3
    step.run(epoch_num=1, accumulate=True)       # gradient accumulation adds (core/engine.py:374-376 `accumulate`)
    assert rel(fl.gflat, 2 * first) < 5e-3
    # the running statistics saw three more batches than the autograd model
    nb1 = dict(m1.named_buffers())["backbone.ERBlock_2.0.rbr_dense.bn.num_batches_tracked" if name == "yolov6n" else "backbone.ERBlock_2.0.rbr_dense.bn.num_batches_tracked"]
    nb2 = dict(m2.named_buffers())["backbone.ERBlock_2.0.rbr_dense.bn.num_batches_tracked"]
    assert int(nb1) == 1 and int(nb2) >= 3
    print(f"{name} graph={graph}: worst gradient rel err vs autograd path {worst:.2e}")


def reference_param_groups(model):
    """build_optimizer of the reference (solver/build.py:12-19), restated for the test."""
    g_bnw, g_w, g_b = [], [], []
    for v in model.modules():
        if hasattr(v, 'bias') and isinstance(v.bias, nn.Parameter):
            g_b.append(v.bias)
        if isinstance(v, nn.BatchNorm2d):
            g_bnw.append(v.weight)
        elif hasattr(v, 'weight') and isinstance(v.weight, nn.Parameter):
            g_w.append(v.weight)
    return g_bnw, g_w, g_b


def test_fused_sgd_ema_matches_torch():
    from yolov6_b200.optim import FusedSGDEMA
    m1, m2 = make_model("yolov6n"), make_model("yolov6n")
    e1, e2 = m1.train_engine(), m2.train_engine()
    g_bnw, g_w, g_b = reference_param_groups(m1)
    opt1 = torch.optim.SGD(g_bnw, lr=0.02, momentum=0.9, nesterov=True)
    opt1.add_param_group({'params': g_w, 'weight_decay': 5e-4})
    opt1.add_param_group({'params': g_b})
    ema1 = {k: v.clone() for k, v in m1.state_dict().items()}
    opt2 = FusedSGDEMA(m2, lr=0.02, momentum=0.9, weight_decay=5e-4, ema_decay=0.9999)
    P1 = dict(m1.named_parameters())
    gen = torch.Generator(device="cuda").manual_seed(0)
    for it in range(3):
        lrs = [0.02 * (it + 1), 0.01, 0.05 / (it + 1)]
        for grp, lr in zip(opt1.param_groups, lrs):
            grp['lr'] = lr
        opt2.set_lr(lrs)
        g = torch.randn(e2.flat.n_train, device="cuda", generator=gen) * 0.1
        e2.flat.gflat.copy_(g)
        e1.flat.gflat.copy_(g)
        for n in e1.flat.names:
            P1[n].grad = e1.flat.grad_view(n).clone()
        # a running statistic changes between steps like in training
        for mm in (m1, m2):
            dict(mm.named_buffers())["backbone.stem.rbr_dense.bn.running_mean"].add_(0.01 * (it + 1))
        opt1.step()
        d = 0.9999 * (1 - math.exp(-(it + 1) / 2000))
        for k, v in m1.state_dict().items():          # ModelEMA.update (utils/ema.py:28-37)
            if v.dtype.is_floating_point:
                ema1[k].mul_(d).add_((1 - d) * v.detach())
        opt2.step()
    torch.cuda.synchronize()
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    for k in sd1:
        if sd1[k].dtype.is_floating_point:
            assert torch.allclose(sd1[k], sd2[k], rtol=1e-5, atol=1e-7), k
    ema2 = opt2.ema_state_dict()
    for k in ema1:
        if ema1[k].dtype.is_floating_point:
            assert torch.allclose(ema1[k], ema2[k], rtol=1e-5, atol=1e-7), f"ema {k}"
    # BottleRep.alpha / detect.proj are in no optimizer group of the reference and stay untouched
    assert torch.equal(sd2["detect.proj"], torch.linspace(0, 16, 17).cuda())
    # the eval engine is rebuilt from the new weights
    m2.eval()
    x = fab.synthetic_images(1, 64, 64, seed=0).cuda()
    m1.eval()
    with torch.no_grad():
        a, b = m1(x)[0], m2(x)[0]
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from yolov6_b200.model import build_model
    from yolov6_b200.step import TrainStep
    B, S = 2, 128
    sd = fab.fabricate_state_dict(golden_keys("yolov6n"), seed=0)
    for k in sd:
        if (".cls_preds." in k or ".reg_preds." in k) and k.endswith("weight"):
            sd[k] = sd[k] * 0.1
    xs = [fab.synthetic_images(B, S, S, seed=30 + r).to(dev) for r in range(world)]
    ts = [oloss.synthetic_targets(B, seed=40 + r).to(dev) for r in range(world)]

    def fresh():
        m = build_model("yolov6n", 80, dev)
        m.load_state_dict(sd)
        return m.train()
    # (1) what the result must be: the sum over ranks of the single-rank gradients (per-rank BatchNorm statistics)
    want = None
    for r in range(world):
        m = fresh()
        st = TrainStep(m, make_loss(S), B, S, S, in_dtype=torch.float32, graph=False, n_buckets=1)
        st.sync = None
        st.load(xs[r], ts[r])
        st.run(epoch_num=1)
        g = st.eng.flat.gflat.clone()
        want = g if want is None else want + g
    # (2) the reference's recipe over the drop-in model: DDP wrapper, loss * world_size (core/engine.py:171-172, 464-466)
    m = fresh()
    ddp = DDP(m, device_ids=[rank], output_device=rank)
    crit = make_loss(S)
    preds, _ = ddp(xs[rank])
    loss, _ = crit(preds, ts[rank], 1, 0, S, S)
    (loss * world).backward()
    fl = m.train_engine().flat
    got_ddp = torch.cat([torch.nn.functional.pad(dict(m.named_parameters())[n].grad.reshape(-1), (0, (-fl.slots[n][1]) % 4)) for n in fl.names])
    e_ddp = float((got_ddp - want).norm() / want.norm())
    # (3) the engine's own bucketed all-reduce (graph segments + NCCL on the communication stream)
    m = fresh()
    st = TrainStep(m, make_loss(S), B, S, S, in_dtype=torch.float32, graph=True, n_buckets=3)
    st.load(xs[rank], ts[rank])
    st.run(epoch_num=1)
    st.run(epoch_num=1)
    torch.cuda.synchronize()
    got = st.eng.flat.gflat
    e_sync = float((got - want).norm() / want.norm())
    gathered = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(gathered, got)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    q.put((rank, e_ddp, e_sync, same, len(st.eng.bucket_range), st.sync.bytes_per_step))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradient_allreduce_matches_ddp():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, e_ddp, e_sync, same, nb, nbytes in res:
        print(f"rank {rank}: DDP wrapper {e_ddp:.2e}, GradSync {e_sync:.2e}, buckets {nb}, {nbytes / 1e6:.1f} MB")
        assert e_ddp < 5e-3 and e_sync < 5e-3 and same and nb == 3
