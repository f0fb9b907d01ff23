"""Data-parallel plumbing: the hot path shards by image -- every rank owns `global_batch / world`
images.  Inference needs no collective; training has exactly one: the gradient all-reduce of DDP
(core/engine.py:456-468), with the loss pre-multiplied by world_size (core/engine.py:171-172) and DDP averaging,
i.e. every rank ends up with the SUM over ranks of the per-rank gradients (per-rank BatchNorm statistics and
per-rank loss normalisers, as in the reference -- no SyncBN).  `GradSync` does that on the flat gradient buffer of
the training engine (flat.py): the buffer is laid out in backward-completion order and cut into a few contiguous
buckets; as soon as the backward pass has unpacked a bucket, its NCCL all-reduce is enqueued on a communication
stream and overlaps the rest of the backward (NVSwitch: bucket size is chosen for launch latency, not link count).
One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def env():
    """(rank, local_rank, world_size) from the torchrun environment (tools/train.py:141-147 reads the same)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_range(n_items, rank, world):
    """Contiguous image shard [lo, hi) of rank `rank`; sizes differ by at most one, union is exact."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    """Device-side max of a scalar over all ranks (multi-GPU timings are reported as the slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_detections(out, count, dst=0):
    """Gather fixed-size detections [b, max_det, 6] + counts [b] from every rank to `dst` (image order =
    rank order).  Returns (out_all, count_all) on dst, (None, None) elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out, count
    world, rank = dist.get_world_size(), dist.get_rank()
    outs = [torch.empty_like(out) for _ in range(world)] if rank == dst else None
    cnts = [torch.empty_like(count) for _ in range(world)] if rank == dst else None
    dist.gather(out, outs, dst=dst)
    dist.gather(count, cnts, dst=dst)
    if rank != dst:
        return None, None
    return torch.cat(outs), torch.cat(cnts)


def bucket_ranges(total, n_buckets, align=4):
    """Contiguous element ranges that cover [0, total) exactly, boundaries aligned to `align` elements."""
    n_buckets = max(1, int(n_buckets))
    cuts = [0]
    for k in range(1, n_buckets):
        cuts.append(min(total, (total * k // n_buckets + align - 1) // align * align))
    cuts.append(total)
    return [(cuts[i], cuts[i + 1]) for i in range(n_buckets)]


class GradSync:
    """Sum of the flat gradient buffer over all ranks, bucket by bucket.

    flat_grad: 1-D tensor (the training engine's `flat.gflat`, or any flat buffer in the gloo tests);
    ranges: contiguous [lo, hi) element ranges in the order in which the backward pass completes them.
    CUDA: `bucket_ready(k)` is called on the compute stream right after bucket k was written; the all-reduce runs on
    `comm_stream` behind an event, `finish()` makes the compute stream wait for all of them.  CPU (gloo): synchronous.
    """

    def __init__(self, flat_grad, ranges, group=None, average=False):
        self.flat, self.ranges, self.group = flat_grad, [(int(a), int(b)) for a, b in ranges], group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.average = average
        self.cuda = flat_grad.is_cuda
        self.works = []
        self.bytes_per_step = sum(b - a for a, b in self.ranges) * flat_grad.element_size()
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(device=flat_grad.device)
            self.events = [torch.cuda.Event() for _ in self.ranges]
            self.t0, self.t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def bucket_ready(self, k):
        lo, hi = self.ranges[k]
        if self.world == 1 or hi <= lo:
            return
        view = self.flat[lo:hi]
        if not self.cuda:
            dist.all_reduce(view, group=self.group)
            if self.average:
                view.div_(self.world)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        self.events[k].record(cur)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(self.events[k])
            if not self.works:
                self.t0.record(self.comm_stream)
            self.works.append(dist.all_reduce(view, group=self.group, async_op=True))

    def all_ready(self):
        for k in range(len(self.ranges)):
            self.bucket_ready(k)

    def finish(self):
        """Compute stream waits for every outstanding bucket (call before the optimizer reads the gradients)."""
        if not self.cuda or not self.works:
            return
        for w in self.works:
            w.wait()                       # makes the CURRENT stream wait for that collective
        self.works = []
        self.t1.record(torch.cuda.current_stream(self.flat.device))
        if self.average:
            self.flat.div_(self.world)

    def last_ms(self):
        """Device time from the start of the first bucket's all-reduce to the point where the compute stream has them all
        (call after a synchronize)."""
        return self.t0.elapsed_time(self.t1) if self.cuda and self.world > 1 else 0.0
