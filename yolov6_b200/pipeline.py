"""Steady-state detection pipeline: host images -> detections, replayed as ONE CUDA graph.

The reference's `Inferer.infer` / `Evaler.predict_model` (core/inferer.py:70-82, core/evaler.py:118-134)
run H2D copy, `.float()/255`, the model and NMS as separate eager calls per batch.  `DetectPipeline`
captures the same sequence -- pinned-host uint8 batch -> device (stem kernel scales by 1/255 on the
fly) -> network kernels -> decode -> batched NMS -> detections back to pinned host memory -- into a
CUDA graph once per (batch, size); `__call__` is then a single graph launch, so the ~80 kernel
launches and all Python/ctypes work disappear from the steady state.
"""
import torch

from .nms import nms_batched


class DetectPipeline:
    def __init__(self, model, batch, height, width, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                 multi_label=False, max_det=300, host_input=True):
        self.model = model.eval()
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("DetectPipeline needs a CUDA model")
        self.kw = dict(conf_thres=conf_thres, iou_thres=iou_thres, classes=classes, agnostic=agnostic,
                       multi_label=multi_label, max_det=max_det)
        self.host_input = host_input
        dt = torch.uint8 if host_input else torch.float32
        self.x_dev = torch.zeros(batch, 3, height, width, dtype=dt, device=self.dev)
        self.x_host = torch.zeros(batch, 3, height, width, dtype=torch.uint8).pin_memory() if host_input else None
        self.out_host = torch.zeros(batch, max_det, 6).pin_memory()
        self.count_host = torch.zeros(batch + 1, dtype=torch.int32).pin_memory()
        self.eng = model.engine()
        self.graph = None
        self._warm()

    def _body(self):
        if self.host_input:
            self.x_dev.copy_(self.x_host, non_blocking=True)
        pred = self.eng.forward(self.x_dev)
        out, count, src, overflow = nms_batched(pred, **self.kw)
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
        self.graph.replay()

    def __call__(self, images=None):
        """images: uint8 [B,3,H,W] host tensor (host_input) or fp32 device tensor.  Returns the list of
        per-image [k,6] detections (host tensors when host_input) like `non_max_suppression`."""
        if images is not None:
            (self.x_host if self.host_input else self.x_dev).copy_(images)
        self.graph.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        if self.host_input:
            counts = self.count_host.tolist()
            if counts[-1]:
                raise RuntimeError("non_max_suppression: candidate overflow (> 65536 per image); raise conf_thres")
            return [self.out_host[i, :counts[i]].clone() for i in range(self.out_host.shape[0])]
        counts = self.count_dev.tolist()
        return [self.out_dev[i, :counts[i]] for i in range(self.out_dev.shape[0])]
