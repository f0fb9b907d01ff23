"""Data-parallel plumbing: the hot path shards by image -- every rank owns
`global_batch / world` images, inference needs no collective, and the only cross-rank traffic of a
benchmark / evaluation run is the reduction of a timing scalar (max over ranks) and, optionally, the
gather of fixed-size detections to rank 0.  One process per GPU, `torch.distributed` (NCCL on GPUs,
gloo in the CPU tests)."""
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
