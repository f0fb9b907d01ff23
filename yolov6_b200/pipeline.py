"""Steady-state detection pipeline: host images -> detections, replayed as ONE CUDA graph.

The reference's `Inferer.infer` / `Evaler.predict_model` (core/inferer.py:70-82, core/evaler.py:118-134)
run H2D copy, `.float()/255`, the model and NMS as separate eager calls per batch.  `DetectPipeline`
captures the same sequence -- pinned-host uint8 batch -> device (stem kernel scales by 1/255 on the
fly) -> network kernels -> decode -> batched NMS -> detections back to pinned host memory -- into a
CUDA graph once per (batch, size); `__call__` is then a single graph launch, so the ~80 kernel
launches and all Python/ctypes work disappear from the steady state.

`overlap_h2d=True` keeps the host->device copy of the images out of the graph and issues it on a copy
stream instead: with two pipelines used alternately (see `DetectRing`) the copy of batch i+1 runs on the
copy engine while the kernels of batch i run on the SMs, so a serving loop is bound by max(copy, compute)
instead of their sum.
"""
import torch

from .nms import nms_batched_head, workspace_bytes


class DetectPipeline:
    def __init__(self, model, batch, height, width, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                 multi_label=False, max_det=300, host_input=True, overlap_h2d=False, copy_stream=None):
        self.model = model.eval()
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("DetectPipeline needs a CUDA model")
        self.kw = dict(conf_thres=conf_thres, iou_thres=iou_thres, classes=classes, agnostic=agnostic,
                       multi_label=multi_label, max_det=max_det)
        self.host_input = host_input
        self.overlap_h2d = bool(overlap_h2d and host_input)
        if self.overlap_h2d:
            self.copy_stream = copy_stream or torch.cuda.Stream(device=self.dev)
            self.ev_copied = torch.cuda.Event()
            self.ev_done = torch.cuda.Event()
            self._used = False
        dt = torch.uint8 if host_input else torch.float32
        self.x_dev = torch.zeros(batch, 3, height, width, dtype=dt, device=self.dev)
        self.x_host = torch.zeros(batch, 3, height, width, dtype=torch.uint8).pin_memory() if host_input else None
        self.out_host = torch.zeros(batch, max_det, 6).pin_memory()
        self.count_host = torch.zeros(batch + 1, dtype=torch.int32).pin_memory()
        self.eng = model.engine()
        self.eng.pin(batch, height, width, dt)          # the graph points at this shape's buffers
        A = sum((height // int(s)) * (width // int(s)) for s in model.graph.strides)
        nc = model.graph.num_classes
        # scratch owned by this pipeline: a shared cache could be replaced (freed) while the graph still references it
        self.nms_ws = torch.empty(workspace_bytes(batch, A, nc, multi_label), dtype=torch.uint8, device=self.dev)
        self.graph = None
        self._warm()

    def _body(self):
        if self.host_input and not self.overlap_h2d:
            self.x_dev.copy_(self.x_host, non_blocking=True)
        # the [B, A, 5+nc] prediction tensor of Detect.forward (effidehead.py:131-139) is never materialised here: the NMS
        # kernels scan the class scores where the cls_pred convs wrote them and decode boxes for the candidates only
        cls, reg, sizes = self.eng.forward(self.x_dev, decode=False)
        out, count, src, overflow = nms_batched_head(cls, reg, sizes, self.model.graph.strides, workspace=self.nms_ws, **self.kw)
        self.out_dev, self.count_dev = out, count
        if self.host_input:
            self.out_host.copy_(out, non_blocking=True)
            self.count_host[:-1].copy_(count, non_blocking=True)
            self.count_host[-1:].copy_(overflow, non_blocking=True)

    def _warm(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):     # plans, tensor maps, cudaFuncSetAttribute, allocator warm-up
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self._body()

    def launch(self):
        """Enqueue one batch (input already in `x_host` / `x_dev`); asynchronous."""
        if self.overlap_h2d:
            cur = torch.cuda.current_stream(self.dev)
            if self._used:
                self.copy_stream.wait_event(self.ev_done)      # the previous batch of THIS pipeline is done with x_dev
            with torch.cuda.stream(self.copy_stream):
                self.x_dev.copy_(self.x_host, non_blocking=True)
                self.ev_copied.record(self.copy_stream)
            cur.wait_event(self.ev_copied)
            self.graph.replay()
            self.ev_done.record(cur)
            self._used = True
            return
        self.graph.replay()

    def __call__(self, images=None):
        """images: uint8 [B,3,H,W] host tensor (host_input) or fp32 device tensor.  Returns the list of
        per-image [k,6] detections (host tensors when host_input) like `non_max_suppression`."""
        if images is not None:
            (self.x_host if self.host_input else self.x_dev).copy_(images)
        self.launch()
        torch.cuda.current_stream(self.dev).synchronize()
        if self.host_input:
            counts = self.count_host.tolist()
            if counts[-1]:
                raise RuntimeError("non_max_suppression: more than 65536 near-identical candidate scores in one image; raise conf_thres")
            return [self.out_host[i, :counts[i]].clone() for i in range(self.out_host.shape[0])]
        counts = self.count_dev.tolist()
        return [self.out_dev[i, :counts[i]] for i in range(self.out_dev.shape[0])]


class DetectRing:
    """Two `DetectPipeline`s used alternately with one shared copy stream: `submit(images)` enqueues a batch and
    returns immediately, `collect()` returns the detections of the oldest batch in flight.  The H2D copy of a
    batch overlaps the kernels of the batch before it (the reference's Inferer is strictly serial,
    core/inferer.py:70-82)."""

    def __init__(self, model, batch, height, width, depth=2, **kw):
        dev = next(model.parameters()).device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.pipes = [DetectPipeline(model, batch, height, width, host_input=True, overlap_h2d=True,
                                     copy_stream=self.copy_stream, **kw) for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.head = self.tail = 0          # submitted / collected counters

    def submit(self, images):
        assert self.head - self.tail < len(self.pipes), "collect() a batch before submitting another"
        p = self.pipes[self.head % len(self.pipes)]
        p.x_host.copy_(images)
        p.launch()
        self.done[self.head % len(self.pipes)].record(torch.cuda.current_stream(p.dev))
        self.head += 1

    def collect(self):
        assert self.tail < self.head, "nothing in flight"
        i = self.tail % len(self.pipes)
        self.done[i].synchronize()
        p = self.pipes[i]
        counts = p.count_host.tolist()
        self.tail += 1
        if counts[-1]:
            raise RuntimeError("non_max_suppression: more than 65536 near-identical candidate scores in one image; raise conf_thres")
        return [p.out_host[j, :counts[j]].clone() for j in range(p.out_host.shape[0])]
