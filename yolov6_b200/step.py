"""One training step as a replayed CUDA graph.

`Trainer.train_in_steps` (core/engine.py:142-176) is, per batch: H2D copy + `.float()/255` (prepro_data, :407-410),
forward under autocast, ComputeLoss (host-side target padding, :184-192 of loss.py), `loss * world_size`, backward
through autograd with DDP's bucketed all-reduce, GradScaler, SGD step, EMA update -- several thousand kernel launches
and a handful of host synchronisations.  `TrainStep` runs the same arithmetic as

    copy images (uint8 or fp32) and padded targets into static buffers
    -> graph segment 0: clear accumulators, repack weights, forward, assignment + loss + its gradients, backward of
       the ops of gradient bucket 0, unpack bucket 0
    -> [all-reduce bucket 0 on the communication stream]  ||  graph segment 1: backward of bucket 1 ...
    -> wait for the all-reduces -> fused SGD + EMA kernel

with static shapes: batch, image size, at most `max_targets` target rows per batch and `max_gt` boxes per image
(rows beyond that are dropped and reported by `overflowed()`).  Gradients are the SUM over ranks of the per-rank
gradients, which is what the reference's `loss * world_size` + DDP averaging produce (core/engine.py:171-172,464-466);
bf16 storage has fp32's exponent range, so no GradScaler is needed.
"""
import torch
import torch.distributed as dist

from . import _lib
from .dist import GradSync


class TrainStep:
    def __init__(self, model, compute_loss, batch, height, width, in_dtype=torch.float32, max_targets=None, max_gt=64,
                 optimizer=None, n_buckets=None, graph=True, group=None, compute_loss_ab=None):
        self.fuse_ab = bool(getattr(model, "fuse_ab", False))
        if self.fuse_ab and compute_loss_ab is None:
            raise ValueError("a fuse_ab model needs compute_loss_ab (yolov6_b200.loss_fuseab.ComputeLoss): core/engine.py:161-166 adds "
                             "the anchor-aided loss to the anchor-free one")
        self.model = model.train()
        self.loss = compute_loss
        self.loss_ab = compute_loss_ab
        self.dev = next(model.parameters()).device
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if n_buckets is None:
            n_buckets = 3 if self.world > 1 else 1
        self.eng = model.train_engine(n_buckets=n_buckets)
        self.B, self.H, self.W = batch, height, width
        self.max_gt = int(max_gt)
        self.max_targets = int(max_targets or batch * max_gt)
        self.targets = torch.full((self.max_targets, 6), -1.0, dtype=torch.float32, device=self.dev)
        self.eng._plan(batch, height, width, in_dtype)
        self.images = self.eng.x_static
        self.opt = optimizer
        self.sync = GradSync(self.eng.flat.gflat, self.eng.bucket_range, group) if self.world > 1 else None
        self.use_graph = bool(graph)
        self.graphs = {}           # assigner branch (True = ATSS warm-up) -> list of graph segments
        self.state = None
        self.accumulate = False
        self.sizes = self.eng.sizes

    # ------------------------------------------------------------------ the step body, in bucket-sized segments
    def _segments(self):
        cuts = self.eng.bucket_call_index()
        return [0] + cuts[:-1], cuts[:-1] + [len(self.eng.bwd_calls)]

    def _segment(self, j, epoch_num):
        eng = self.eng
        firsts, lasts = self._segments()
        if j == 0:
            sp = _lib.stream_ptr()
            eng.begin_step(sp)
            eng.flat.iflat.add_(1)
            eng.run_forward(sp)
            self.state = self.loss.forward_backward(eng.cls, eng.reg, self.sizes, self.targets, epoch_num, self.H, self.W,
                                                    max_gt=self.max_gt, grad_scores=eng.grad_cls, grad_distri=eng.grad_reg)
            if self.fuse_ab:      # total_loss += total_loss_ab; loss_items += loss_items_ab (core/engine.py:164-166)
                st_ab = self.loss_ab.forward_backward(eng.cls_ab, eng.reg_ab, self.sizes, self.targets, epoch_num, self.H, self.W,
                                                      max_gt=self.max_gt, grad_scores=eng.grad_cls_ab, grad_distri=eng.grad_reg_ab)
                self.state["out_af"], self.state["out_ab"] = self.state["out"], st_ab["out"]
                total = self.state["out"].clone()
                total[:4] += st_ab["out"][:4]
                self.state["out"] = total
                self.state["keep_ab"] = st_ab
        eng.backward(None, None, accumulate=self.accumulate, first=firsts[j], last=lasts[j])

    def _capture(self, epoch_num):
        nseg = len(self._segments()[0])
        fl = self.eng.flat
        snap = (fl.pflat.clone(), fl.gflat.clone(), fl.iflat.clone())   # warm-up runs must not train: running statistics,
        side = torch.cuda.Stream(device=self.dev)                       # batch counters and accumulated gradients are restored
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(2):                      # plans, tensor maps, cudaFuncSetAttribute, allocator warm-up
                for j in range(nseg):
                    self._segment(j, epoch_num)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        fl.pflat.copy_(snap[0])
        fl.gflat.copy_(snap[1])
        fl.iflat.copy_(snap[2])
        del snap
        torch.cuda.synchronize(self.dev)
        graphs = []
        pool = None
        for j in range(nseg):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, pool=pool):
                self._segment(j, epoch_num)
            pool = gph.pool()
            graphs.append(gph)
        return graphs

    # ------------------------------------------------------------------ public
    def load(self, images, targets):
        """images: [B,3,H,W] (host or device; pinned host memory makes the copy asynchronous); targets: [n,6] fp32
        (img, cls, cx, cy, w, h normalised), n <= max_targets."""
        self.images.copy_(images, non_blocking=True)
        n = int(targets.shape[0])
        if n > self.max_targets:
            raise RuntimeError(f"{n} target rows exceed max_targets={self.max_targets}")
        self.targets[:n].copy_(targets, non_blocking=True)
        if n < self.max_targets:
            self.targets[n:].fill_(-1.0)             # rows with image index -1 are ignored by yv6_targets_pad

    def run(self, epoch_num=0, accumulate=False):
        """Forward + loss + backward (+ gradient all-reduce).  Returns the device tensor float64[8] =
        (loss, iou, dfl, cls, target_scores_sum, num_pos) of this rank -- no host synchronisation."""
        atss = epoch_num < self.loss.warmup_epoch
        nseg = len(self._segments()[0])
        if self.use_graph:
            if accumulate != self.accumulate or atss not in self.graphs:
                self.accumulate = accumulate
                self.graphs = {atss: self._capture(epoch_num)}
            graphs = self.graphs[atss]
        for j in range(nseg):
            if self.use_graph:
                graphs[j].replay()
            else:
                self.accumulate = accumulate
                self._segment(j, epoch_num)
            if self.sync is not None:
                self.sync.bucket_ready(j)
        if self.sync is not None:
            self.sync.finish()
        return self.state["out"]

    def step(self, images=None, targets=None, epoch_num=0):
        """load -> run -> optimizer (when one was given)."""
        if images is not None:
            self.load(images, targets)
        out = self.run(epoch_num)
        if self.opt is not None:
            self.opt.step()
        return out

    def overflowed(self):
        """True when the last batch had an image with more than max_gt boxes (host read)."""
        return bool((self.state["gt_count"] > self.max_gt).any().item())
