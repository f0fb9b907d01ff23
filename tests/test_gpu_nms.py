"""GPU parity of the batched NMS kernels: bit-exact kept rows (boxes, scores, class ids, order) against
the reference's golden outputs and against the oracle on further seeded / edge-case inputs."""
import numpy as np
import pytest
import torch

from conftest import golden_json, golden_npz
from oracle import fabricate as fab
from oracle import nms as onms

pytestmark = pytest.mark.gpu


def run(p, **kw):
    from yolov6_b200.nms import non_max_suppression
    return [o.cpu().numpy() for o in non_max_suppression(p.cuda(), **kw)]


def test_bit_exact_vs_reference_golden():
    g = golden_npz("nms.npz")
    for i, (B, A, nc, seed, kw) in enumerate(golden_json("nms_cases.json")):
        p = fab.synthetic_predictions(B, A, nc, seed)
        out = run(p, **kw)
        counts = np.array([o.shape[0] for o in out])
        assert np.array_equal(counts, g[f"c{i}_counts"]), (i, counts.tolist(), g[f"c{i}_counts"].tolist())
        rows = np.concatenate(out) if counts.sum() else np.zeros((0, 6), np.float32)
        assert np.array_equal(rows, g[f"c{i}_rows"]), f"case {i}: kept rows differ from the reference"


@pytest.mark.parametrize("B,A,nc,seed,kw", [
    (4, 8400, 80, 11, dict(conf_thres=0.25, iou_thres=0.45)),
    (2, 34000, 80, 12, dict(conf_thres=0.5, iou_thres=0.65, max_det=1000)),
    (3, 1000, 80, 13, dict(conf_thres=0.03, iou_thres=0.65, multi_label=True)),
    (2, 777, 3, 14, dict(conf_thres=0.001, iou_thres=0.3, agnostic=True, max_det=100)),
    (2, 500, 80, 15, dict(conf_thres=0.2, iou_thres=0.0)),
    (2, 500, 80, 16, dict(conf_thres=0.2, iou_thres=1.0)),
    (1, 1, 1, 17, dict(conf_thres=0.0, iou_thres=0.5)),
])
def test_bit_exact_vs_oracle(B, A, nc, seed, kw):
    p = fab.synthetic_predictions(B, A, nc, seed)
    out = run(p, **kw)
    ref = onms.non_max_suppression(p.numpy(), **kw)
    for a, b in zip(out, ref):
        assert a.shape == b.shape
        assert np.array_equal(a, b)


def test_source_indices_and_empty_images():
    from yolov6_b200.nms import nms_batched
    p = fab.synthetic_predictions(3, 600, 80, seed=21)
    p[1, :, 5:] = 0.0                                  # image 1: nothing passes
    out, count, src, overflow = nms_batched(p.cuda(), 0.25, 0.45)
    ref, ref_idx = onms.non_max_suppression(p.numpy(), 0.25, 0.45, return_index=True)
    count = count.cpu().numpy()
    assert int(overflow.item()) == 0 and count[1] == 0
    for b in range(3):
        assert count[b] == ref[b].shape[0]
        assert np.array_equal(src[b, :count[b]].cpu().numpy(), ref_idx[b].astype(np.int32))


def test_detect_pipeline_matches_eager_model_plus_nms():
    """DetectPipeline (one CUDA graph: H2D -> network -> decode -> NMS -> D2H) returns exactly what the eager
    calls `model(x)` + `non_max_suppression` return (core/inferer.py:70-82)."""
    from yolov6_b200.model import build_model
    from yolov6_b200.nms import non_max_suppression
    from yolov6_b200.pipeline import DetectPipeline
    from yolov6_b200.synth import randomize_
    dev = torch.device("cuda:0")
    m = randomize_(build_model("yolov6n", 80, dev), seed=3).eval()
    B, S = 3, 192
    kw = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
    pipe = DetectPipeline(m, B, S, S, host_input=True, **kw)
    g = torch.Generator().manual_seed(9)
    for _ in range(2):
        img = (torch.rand(B, 3, S, S, generator=g) * 255).to(torch.uint8)
        dets = pipe(img)
        with torch.no_grad():
            pred = m(img.to(dev))[0]
        ref = non_max_suppression(pred, **kw)
        assert len(dets) == B
        assert sum(len(d) for d in dets) > 0
        for d, r in zip(dets, ref):
            assert torch.equal(d, r.cpu())


def test_detect_ring_overlapped_copies_match_eager():
    """DetectRing: H2D copies on a copy stream overlapping the previous batch's kernels; results must equal the
    eager model + NMS for every batch, in order."""
    from yolov6_b200.model import build_model
    from yolov6_b200.nms import non_max_suppression
    from yolov6_b200.pipeline import DetectRing
    from yolov6_b200.synth import randomize_
    dev = torch.device("cuda:0")
    m = randomize_(build_model("yolov6n", 80, dev), seed=3).eval()
    B, S = 2, 160
    kw = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
    ring = DetectRing(m, B, S, S, **kw)
    g = torch.Generator().manual_seed(4)
    imgs = [(torch.rand(B, 3, S, S, generator=g) * 255).to(torch.uint8) for _ in range(5)]
    outs = []
    ring.submit(imgs[0])
    for i in range(1, len(imgs)):
        ring.submit(imgs[i])
        outs.append(ring.collect())
    outs.append(ring.collect())
    for img, dets in zip(imgs, outs):
        with torch.no_grad():
            ref = non_max_suppression(m(img.to(dev))[0], **kw)
        for d, r in zip(dets, ref):
            assert torch.equal(d, r.cpu())


@pytest.mark.parametrize("host_input", [True, False])
def test_detect_stream_software_pipeline_matches_eager(host_input):
    """DetectStream: the graph of step i runs the network of batch i and, in parallel, the NMS of batch i - 1 over the
    other head-output set.  Every batch's detections (collected one step later, the last one after drain()) must equal
    the eager model + NMS, in order."""
    from yolov6_b200.model import build_model
    from yolov6_b200.nms import non_max_suppression
    from yolov6_b200.pipeline import DetectStream
    from yolov6_b200.synth import randomize_
    dev = torch.device("cuda:0")
    m = randomize_(build_model("yolov6n", 80, dev), seed=3).eval()
    B, S = 2, 160
    kw = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
    ds = DetectStream(m, B, S, S, host_input=host_input, **kw)
    g = torch.Generator().manual_seed(4)
    imgs = [(torch.rand(B, 3, S, S, generator=g) * 255).to(torch.uint8) for _ in range(5)]
    feed = imgs if host_input else [(i.float() / 255).to(dev) for i in imgs]
    outs = []
    for x in feed:
        ds.submit(x)
        r = ds.collect()
        if r is not None:
            outs.append(r)
    assert ds.collect() is None          # the last batch has not been post-processed yet
    ds.drain()
    outs.append(ds.collect())
    assert len(outs) == len(imgs) and ds.collect() is None
    total = 0
    for x, dets in zip(feed, outs):
        with torch.no_grad():
            ref = non_max_suppression(m(x.to(dev))[0], **kw)
        for d, r in zip(dets, ref):
            total += len(r)
            assert torch.equal(d.cpu(), r.cpu())
    assert total > 0


def test_detect_farm_two_independent_lanes_match_eager():
    """DetectFarm: two DetectStreams with their own engines / buffers / CUDA streams fed round-robin; the batches of the two
    lanes overlap on the GPU.  Detections per batch must equal the eager model + NMS."""
    from yolov6_b200.model import build_model
    from yolov6_b200.nms import non_max_suppression
    from yolov6_b200.pipeline import DetectFarm
    from yolov6_b200.synth import randomize_
    dev = torch.device("cuda:0")
    m = randomize_(build_model("yolov6n", 80, dev), seed=5).eval()
    B, S = 2, 160
    kw = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
    farm = DetectFarm(m, B, S, S, lanes=2, host_input=True, **kw)
    g = torch.Generator().manual_seed(11)
    imgs = [(torch.rand(B, 3, S, S, generator=g) * 255).to(torch.uint8) for _ in range(8)]
    outs, nxt = {}, [0, 1]                   # nxt[lane] = index of the oldest uncollected batch of that lane
    for j, x in enumerate(imgs):
        ln = j % 2
        farm.submit(x)                       # batch j goes to lane j % 2
        r = farm.lanes[ln].collect()         # the lane's previous batch, post-processed by the step just submitted
        if r is not None:
            outs[nxt[ln]] = r
            nxt[ln] += 2
    for ln, lane in enumerate(farm.lanes):
        with torch.cuda.stream(farm.streams[ln]):
            lane.drain()
        torch.cuda.synchronize()
        r = lane.collect()
        assert r is not None and lane.collect() is None
        outs[nxt[ln]] = r
    assert sorted(outs) == list(range(len(imgs)))
    total = 0
    for j, x in enumerate(imgs):
        with torch.no_grad():
            ref = non_max_suppression(m(x.to(dev))[0], **kw)
        for d, r in zip(outs[j], ref):
            total += len(r)
            assert torch.equal(d, r.cpu())
    assert total > 0


@pytest.mark.parametrize("name,size,kw", [("yolov6n", 160, dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)),
                                          ("yolov6m", 128, dict(conf_thres=0.001, iou_thres=0.45, max_det=100)),
                                          ("yolov6l6", 128, dict(conf_thres=0.0005, iou_thres=0.6, multi_label=True, agnostic=True, classes=[0, 1, 2, 5, 7]))])
def test_nms_on_head_tensors_equals_nms_on_decoded_predictions(name, size, kw):
    """yv6_nms_batched_head (no [B,A,5+nc] tensor; boxes decoded for candidates only, plain ltrb and DFL heads) returns the
    rows of yv6_head_decode + yv6_nms_batched bit for bit."""
    from yolov6_b200.model import build_model
    from yolov6_b200.nms import nms_batched, nms_batched_head
    from yolov6_b200.synth import randomize_
    dev = torch.device("cuda:0")
    m = randomize_(build_model(name, 80, dev), seed=4).eval()
    x = torch.rand(3, 3, size, size, generator=torch.Generator().manual_seed(2)).to(dev)
    eng = m.engine()
    with torch.no_grad():
        pred = eng.forward(x).clone()
        cls, reg, sizes = eng.forward(x, decode=False)
        # the synthetic checkpoint's score level depends on model and input size: put the threshold where ~4000 of the
        # batch's (anchor, class) scores pass, so that every case has candidates, suppression and survivors
        kw = dict(kw, conf_thres=float(cls.flatten().kthvalue(cls.numel() - 4000).values))
        a = nms_batched(pred, **kw)
        b = nms_batched_head(cls, reg, sizes, m.graph.strides, **kw)
    print(name, "detections per image:", a[1].tolist())
    assert int(a[1].sum()) > 0, "the case must produce detections to mean anything"
    for u, v in zip(a, b):
        assert torch.equal(u, v)


@pytest.mark.parametrize("spread,expect_fallback", [(4.0, True), (200.0, False)])
def test_head_mode_topk_prefix_and_its_fallback(spread, expect_fallback):
    """Head-tensor mode sorts only the best ~2048 candidates first; when suppression is so heavy that they do not yield
    max_det boxes the full sort + a second greedy pass must take over.  Synthetic head tensors: many candidates, boxes
    packed into a few clusters (spread 4 px: almost everything is suppressed) or scattered."""
    import ctypes as C
    from yolov6_b200 import _lib
    from yolov6_b200.nms import nms_batched, nms_batched_head
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    B, nc, sizes, strides = 2, 80, [(40, 40), (20, 20), (10, 10)], [8, 16, 32]
    A = sum(h * w for h, w in sizes)
    cls = (torch.rand(B, A, nc, generator=g) ** 3).to(dev).contiguous()             # ~ 47 k scores above 0.05 per image
    # ltrb distances that pull every box towards one of four cluster centres
    reg = torch.zeros(B, A, 4)
    off = 0
    for (h, w), s in zip(sizes, strides):
        ys, xs = torch.meshgrid(torch.arange(h) + 0.5, torch.arange(w) + 0.5, indexing="ij")
        cx = (torch.randint(0, 4, (h * w,), generator=g).float() * 60 + 60 + torch.randn(h * w, generator=g) * spread) / s
        cy = (torch.randint(0, 4, (h * w,), generator=g).float() * 60 + 60 + torch.randn(h * w, generator=g) * spread) / s
        half = 20.0 / s
        ax, ay = xs.reshape(-1), ys.reshape(-1)
        reg[:, off:off + h * w] = torch.stack([ax - (cx - half), ay - (cy - half), (cx + half) - ax, (cy + half) - ay], 1)
        off += h * w
    reg = reg.to(dev).contiguous()
    pred = torch.empty(B, A, 5 + nc, device=dev)
    lh, lw = (C.c_int32 * 3)(*[h for h, _ in sizes]), (C.c_int32 * 3)(*[w for _, w in sizes])
    ls = (C.c_float * 3)(*[float(s) for s in strides])
    _lib.check(_lib.lib().yv6_head_decode(_lib.handle(0), C.c_void_p(cls.data_ptr()), C.c_void_p(reg.data_ptr()), C.c_void_p(pred.data_ptr()),
                                          B, nc, 4, 3, lh, lw, ls, _lib.stream_ptr()))
    # class-agnostic suppression for the clustered case: with per-class NMS the 80 classes x 16 clusters alone give max_det rows
    kw = dict(conf_thres=0.05, iou_thres=0.5, multi_label=True, max_det=300, agnostic=expect_fallback)
    a = nms_batched(pred, **kw)
    b = nms_batched_head(cls, reg, sizes, strides, **kw)
    counts = a[1].tolist()
    print("spread", spread, "kept per image", counts)
    assert (min(counts) < 300) == expect_fallback, "the case must (not) exhaust the prefix to test what it claims"
    for u, v in zip(a[:3], b[:3]):
        assert torch.equal(u, v)
